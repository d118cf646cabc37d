#!/usr/bin/env python
"""How far do two CPU fp32 implementations of the reference's own arithmetic disagree?

Same fp32 frames (the oracle's framing, bit-identical to torchaudio), same fp32 tables, same power / mel / log ops --
only the FFT library differs: torch.fft.rfft (MKL on x86: the reference's path) vs numpy's / scipy's pocketfft in
float32.  SURVEY.md 8(c) metrics (ii) and (iv) are evaluated between them exactly as they are evaluated for the GPU
kernel, on cfg2-like white-noise and speech-like utterances.  CPU only; prints one line per input kind.
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.fft
import torch
from oracle import frontend_oracle as O

torch.set_num_threads(8)


def fbank_with_fft(wav, fft):
    frames = O.kaldi_windowed_frames(wav[0].to(torch.float32), 400, 160, 512, 1.0)
    X = fft(frames)
    spectrum = X.abs().pow(2.0)
    banks = O.kaldi_mel_banks(80, 512, 16000.0)
    mel = torch.mm(spectrum, banks.T)
    return torch.max(mel, torch.tensor(O.EPS32)).log()


def fft_numpy(fr):
    x = np.fft.rfft(fr.numpy())
    assert x.dtype == np.complex64, x.dtype
    return torch.from_numpy(x)


def fft_scipy(fr):
    x = scipy.fft.rfft(fr.numpy())
    assert x.dtype == np.complex64, x.dtype
    return torch.from_numpy(x)


def metrics(a, b, truth):
    """a = candidate, b = the oracle (torch/MKL)."""
    d = (a - b).abs()
    ii = float(d.max() / b.abs().max())                                # (ii) norm-relative, all bins
    ea = (a.double() - truth).abs().max(0).values                      # per mel bin, max over frames
    eb = (b.double() - truth).abs().max(0).values
    ratio = ea / eb.clamp_min(1e-30)
    return ii, float(ratio.max()), int((ratio > 1.5).sum()), float(d[:, :3].max()), float(d[:, 3:].max()), \
        float((b.double() - truth).abs().max() / b.abs().max())


def main():
    n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    res = {}
    for kind, gen in (("noise", O.synth_noise), ("speech", O.synth_speechlike)):
        for name, fft in (("numpy_pocketfft_f32", fft_numpy), ("scipy_pocketfft_f32", fft_scipy)):
            fails_ii, fails_iv, worst_ii, worst_iv, bins_iv, worst_ref_truth = 0, 0, 0.0, 0.0, 0, 0.0
            for s in range(n_utts):
                x = gen(128000, 1000 + s)
                ref = O.kaldi_fbank(x)
                tru = O.truth64_fbank(x)
                got = fbank_with_fft(x, fft)
                ii, iv, nb, lo, hi, rt = metrics(got, ref, tru)
                fails_ii += ii > 1e-4
                fails_iv += iv > 1.5
                bins_iv += nb
                worst_ii = max(worst_ii, ii)
                worst_iv = max(worst_iv, iv)
                worst_ref_truth = max(worst_ref_truth, rt)
            res[(kind, name)] = dict(utts=n_utts, fail_ii=int(fails_ii), worst_ii=worst_ii, fail_iv=int(fails_iv),
                                     worst_iv=worst_iv, bins_over_1p5=bins_iv, oracle_vs_truth_normrel=worst_ref_truth)
            print(kind, name, json.dumps(res[(kind, name)]), flush=True)


if __name__ == "__main__":
    main()
