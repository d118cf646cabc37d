#!/usr/bin/env python
"""CPU study behind DESIGN.md section 2: where does the kernel's FFT lose against a native real-input fp32 FFT?
Emulates the kernel's 512-point real FFT (256-point complex FFT of the packed sequence as 16 x 16 with radix-4 x 4
butterflies, twiddles rounded once from fp64, real-FFT split) in numpy float32, with any subset of its stages switched to
float64, and compares the per-bin error of X[k] (against a float64 FFT, in units of eps * rms|X|) with scipy's pocketfft
in float32.  White-noise frames with the kernel's framing (x[n] - x[n-1], Povey window)."""
import sys
import numpy as np
import scipy.fft

rng = np.random.default_rng(0)
F = 20000
x = rng.standard_normal((F, 401)).astype(np.float32)
y = (x[:, 1:] - x[:, :-1]).astype(np.float32)
win = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(400) / 399)) ** 0.85
y = (y * win.astype(np.float32)).astype(np.float32)
buf = np.zeros((F, 512), np.float32)
buf[:, :400] = y
truth = scipy.fft.rfft(buf.astype(np.float64), axis=1)
pocket = scipy.fft.rfft(buf, axis=1)
assert pocket.dtype == np.complex64
scale = np.sqrt((np.abs(truth) ** 2).mean())
eps = np.finfo(np.float32).eps


def report(name, X):
    err = np.abs(X.astype(np.complex128) - truth) / (eps * scale)
    lo, mid, hi = err[:, 1:33], err[:, 33:129], err[:, 129:256]
    print("%-58s rms %.2f | bins 1-32 %.2f, 33-128 %.2f, 129-255 %.2f | p99.9 %.2f  max %.2f" % (
        name, np.sqrt((err ** 2).mean()), np.sqrt((lo ** 2).mean()), np.sqrt((mid ** 2).mean()), np.sqrt((hi ** 2).mean()),
        np.quantile(err, 0.999), err.max()))


def dft16(a, dt):
    """a: (..., 16) complex along the last axis, n = 4 p + q; radix-4 over p, twiddle W16^(q c), radix-4 over q."""
    ct = np.complex128 if dt == np.float64 else np.complex64
    a = a.astype(ct)
    w4 = np.array([[np.exp(-2j * np.pi * p * c / 4) for c in range(4)] for p in range(4)])   # exact +-1, +-i
    v = a.reshape(a.shape[:-1] + (4, 4))                    # [..., p, q]
    # first radix-4 over p (adds / subs only: multiply by exact units)
    y0 = np.stack([sum((v[..., p, :] * ct(w4[p, c])) for p in range(4)) for c in range(4)], -2)   # [..., c, q]
    tw = np.array([[np.exp(-2j * np.pi * q * c / 16) for q in range(4)] for c in range(4)]).astype(ct)
    y0 = (y0 * tw).astype(ct)
    out = np.zeros(a.shape, ct)
    for d in range(4):
        out[..., np.arange(4) + 4 * d] = sum((y0[..., :, q] * ct(w4[q, d])) for q in range(4))
    return out


def kernel_fft(buf, s1, tw, s2, split):
    """stage dtypes: each of s1 (first DFT16), tw (W256 twiddle), s2 (second DFT16), split in {np.float32, np.float64}"""
    c = lambda dt: np.complex128 if dt == np.float64 else np.complex64
    z = (buf[:, 0::2] + 1j * buf[:, 1::2]).astype(np.complex64)          # z[n], n = t + 16 j
    a = z.reshape(F, 16, 16)                                               # [f, j, t]
    A = dft16(np.swapaxes(a, 1, 2), s1)                                    # [f, t, K1]
    W = np.exp(-2j * np.pi * np.outer(np.arange(16), np.arange(16)) / 256)   # [t, K1]
    Wt = W.astype(np.complex64).astype(c(tw)) if tw == np.float32 else W
    B = (A.astype(c(tw)) * Wt).astype(c(tw))
    C = dft16(np.swapaxes(B, 1, 2), s2)                                    # [f, K1, K2] -> Z[K1 + 16 K2]
    Z = np.zeros((F, 256), c(s2))
    for K1 in range(16):
        Z[:, K1 + 16 * np.arange(16)] = C[:, K1, :]
    Z = Z.astype(c(split))
    k = np.arange(257)
    Zk = np.concatenate([Z, Z[:, :1]], 1)[:, k]
    Zc = np.conj(np.concatenate([Z, Z[:, :1]], 1)[:, 256 - k])
    W5 = np.exp(-2j * np.pi * k / 512)
    W5 = W5.astype(np.complex64).astype(c(split)) if split == np.float32 else W5
    E = ((Zk + Zc) * c(split)(0.5)).astype(c(split))
    O = ((Zk - Zc) * c(split)(-0.5j)).astype(c(split))
    return (E + (W5 * O).astype(c(split))).astype(c(split))


f32, f64 = np.float32, np.float64
report("scipy pocketfft rfft, float32", pocket)
report("kernel structure, all float32", kernel_fft(buf, f32, f32, f32, f32))
report("  first DFT16 in float64", kernel_fft(buf, f64, f32, f32, f32))
report("  W256 twiddle multiply in float64", kernel_fft(buf, f32, f64, f32, f32))
report("  second DFT16 in float64", kernel_fft(buf, f32, f32, f64, f32))
report("  real-FFT split in float64", kernel_fft(buf, f32, f32, f32, f64))
report("  everything but the split in float64", kernel_fft(buf, f64, f64, f64, f32))
report("kernel structure, all float64 (sanity)", kernel_fft(buf, f64, f64, f64, f64))
