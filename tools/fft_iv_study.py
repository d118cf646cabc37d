#!/usr/bin/env python
"""CPU study: is SURVEY metric (iv) attainable by an independent fp32 FFT of the kernel's structure?  The oracle's fp32
frames, tables, power / mel / log -- only the FFT differs: torch (the oracle), scipy pocketfft float32, and a numpy float32
emulation of the kernel's algorithm (256-point complex FFT of the packed sequence as 16 x 16, radix-4 x 4 butterflies,
fp64-rounded twiddles, real-FFT split).  Prints (ii) / (iv) like tools/parity_matrix.py."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.fft
import torch
from oracle import frontend_oracle as O

torch.set_num_threads(8)


def dft16(a):
    a = a.astype(np.complex64)
    w4 = np.array([[np.exp(-2j * np.pi * p * c / 4) for c in range(4)] for p in range(4)]).round()
    v = a.reshape(a.shape[:-1] + (4, 4))
    y0 = np.stack([sum((v[..., p, :] * np.complex64(w4[p, c])) for p in range(4)) for c in range(4)], -2)
    tw = np.array([[np.exp(-2j * np.pi * q * c / 16) for q in range(4)] for c in range(4)]).astype(np.complex64)
    y0 = (y0 * tw).astype(np.complex64)
    out = np.zeros(a.shape, np.complex64)
    for d in range(4):
        out[..., np.arange(4) + 4 * d] = sum((y0[..., :, q] * np.complex64(w4[q, d])) for q in range(4))
    return out


def kernel_fft(fr):
    buf = fr.numpy().astype(np.float32)
    F = buf.shape[0]
    z = (buf[:, 0::2] + 1j * buf[:, 1::2]).astype(np.complex64)
    A = dft16(np.swapaxes(z.reshape(F, 16, 16), 1, 2))
    W = np.exp(-2j * np.pi * np.outer(np.arange(16), np.arange(16)) / 256).astype(np.complex64)
    B = (A * W).astype(np.complex64)
    C = dft16(np.swapaxes(B, 1, 2))
    Z = np.zeros((F, 256), np.complex64)
    for K1 in range(16):
        Z[:, K1 + 16 * np.arange(16)] = C[:, K1, :]
    k = np.arange(257)
    Zx = np.concatenate([Z, Z[:, :1]], 1)
    Zk, Zc = Zx[:, k], np.conj(Zx[:, 256 - k])
    W5 = np.exp(-2j * np.pi * k / 512).astype(np.complex64)
    E = ((Zk + Zc) * np.complex64(0.5)).astype(np.complex64)
    Od = ((Zk - Zc) * np.complex64(-0.5j)).astype(np.complex64)
    return torch.from_numpy((E + (W5 * Od).astype(np.complex64)).astype(np.complex64))


def fbank_with_fft(wav, fft):
    frames = O.kaldi_windowed_frames(wav[0].to(torch.float32), 400, 160, 512, 1.0)
    spectrum = fft(frames).abs().pow(2.0)
    mel = torch.mm(spectrum, O.kaldi_mel_banks(80, 512, 16000.0).T)
    return torch.max(mel, torch.tensor(O.EPS32)).log()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    for kind, gen, N in (("cfg2-shaped noise", O.synth_noise, 128000), ("speech-like", O.synth_speechlike, 128000)):
        for name, fft in (("scipy pocketfft f32", lambda fr: torch.from_numpy(scipy.fft.rfft(fr.numpy()))), ("kernel-structure f32 (numpy)", kernel_fft)):
            ii_fail = iv_utts = iv_dims = 0
            worst_ii = worst_iv = 0.0
            for s in range(n):
                x = gen(N, 100 + s)
                ref, tru, got = O.kaldi_fbank(x), O.truth64_fbank(x), fbank_with_fft(x, fft)
                ii = float((got - ref).abs().max() / ref.abs().max())
                eg = (got.double() - tru).abs().max(0).values
                er = (ref.double() - tru).abs().max(0).values
                ratio = eg / er.clamp_min(1e-30)
                nd = int((ratio > 1.5).sum())
                ii_fail += ii > 1e-4; worst_ii = max(worst_ii, ii)
                iv_utts += nd > 0; iv_dims += nd; worst_iv = max(worst_iv, float(ratio.max()))
            print("%-18s %-30s utts %3d | (ii) fail %3d worst %.3g | (iv) fail utts %3d, dims %4d of %5d, worst ratio %.2f" % (
                kind, name, n, ii_fail, worst_ii, iv_utts, iv_dims, n * 80, worst_iv), flush=True)


if __name__ == "__main__":
    main()
