"""Host-side constant tables, built once per FrontEnd with the same fp32 torch CPU arithmetic the
reference runs on every call, so the device kernels read bit-identical tables.

ta: = path inside torchaudio (the library the reference's ``_kaidi_wav2mel`` calls,
ref: lid/audio_processor.py:41-69).
"""
from __future__ import annotations

import math

import torch


def povey_window(frame_len: int) -> torch.Tensor:
    """Kaldi's default window: hann(periodic=False) ** 0.85.   ta: compliance/kaldi.py:98-100"""
    return torch.hann_window(frame_len, periodic=False, dtype=torch.float32).pow(0.85)


def kaldi_window(window_type: str, frame_len: int, blackman_coeff: float = 0.42) -> torch.Tensor:
    """The five windows of ``_feature_window_function`` with its fp32 torch expressions.   ta: compliance/kaldi.py:86-113"""
    if window_type == "povey":
        return povey_window(frame_len)
    if window_type == "hanning":
        return torch.hann_window(frame_len, periodic=False, dtype=torch.float32)
    if window_type == "hamming":
        return torch.hamming_window(frame_len, periodic=False, alpha=0.54, beta=0.46, dtype=torch.float32)
    if window_type == "rectangular":
        return torch.ones(frame_len, dtype=torch.float32)
    if window_type == "blackman":
        a = 2 * math.pi / (frame_len - 1)
        n = torch.arange(frame_len, dtype=torch.float32)
        return (blackman_coeff - 0.5 * torch.cos(a * n) + (0.5 - blackman_coeff) * torch.cos(2 * a * n)).to(torch.float32)
    raise ValueError("Invalid window type " + window_type)


def _mel(freq):
    return 1127.0 * math.log(1.0 + freq / 700.0)


def mel_banks(n_mels: int, fft_len: int, sample_rate: float, low_freq: float = 20.0,
              high_freq: float = 0.0) -> torch.Tensor:
    """Dense (n_mels, fft_len/2 + 1) triangular filters on the Kaldi mel scale, vtln_warp = 1, last
    (Nyquist) column zero.                                        ta: compliance/kaldi.py:436-511,621-630

    The expression order below is the one torchaudio evaluates (python floats for the scalar edges,
    fp32 tensors for everything broadcast), which is what fixes the fp32 rounding of every weight."""
    if n_mels <= 3:
        raise AssertionError("Must have at least 3 mel bins")
    nyquist = 0.5 * sample_rate
    if high_freq <= 0.0:
        high_freq += nyquist
    if not (0.0 <= low_freq < nyquist and 0.0 < high_freq <= nyquist and low_freq < high_freq):
        raise AssertionError("Bad values in options: low-freq {} and high-freq {} vs. nyquist {}".format(
            low_freq, high_freq, nyquist))
    bin_width = sample_rate / fft_len
    lo, hi = _mel(low_freq), _mel(high_freq)
    step = (hi - lo) / (n_mels + 1)
    idx = torch.arange(n_mels).unsqueeze(1)
    left_edge = lo + idx * step
    peak = lo + (idx + 1.0) * step
    right_edge = lo + (idx + 2.0) * step
    fft_mel = (1127.0 * (1.0 + (bin_width * torch.arange(fft_len / 2)) / 700.0).log()).unsqueeze(0)
    rising = (fft_mel - left_edge) / (peak - left_edge)
    falling = (right_edge - fft_mel) / (right_edge - peak)
    tri = torch.max(torch.zeros(1), torch.min(rising, falling))
    return torch.nn.functional.pad(tri, (0, 1), mode="constant", value=0).contiguous()


def dct_matrix(n_ceps: int, n_mels: int) -> torch.Tensor:
    """(n_mels, n_ceps) orthonormal DCT-II whose first column is sqrt(1/n_mels).
    ta: compliance/kaldi.py:648-658, functional/functional.py:636-667"""
    pos = torch.arange(float(n_mels))
    order = torch.arange(float(n_mels)).unsqueeze(1)
    basis = torch.cos(math.pi / float(n_mels) * (pos + 0.5) * order)
    basis[0] *= 1.0 / math.sqrt(2.0)
    basis *= math.sqrt(2.0 / float(n_mels))
    basis = basis.t().clone()
    basis[:, 0] = math.sqrt(1 / float(n_mels))
    return basis[:, :n_ceps].contiguous()


def lifter(n_ceps: int, cepstral_lifter: float) -> torch.Tensor:
    """1 + 0.5 L sin(pi i / L); ones when L == 0.               ta: compliance/kaldi.py:661-666,793-796"""
    if cepstral_lifter == 0.0:
        return torch.ones(n_ceps, dtype=torch.float32)
    i = torch.arange(n_ceps)
    return (1.0 + 0.5 * cepstral_lifter * torch.sin(math.pi * i / cepstral_lifter)).to(torch.float32)


def hann_window(win_length: int) -> torch.Tensor:
    """Periodic Hann window, the default of torchaudio's Spectrogram (ta: transforms/_transforms.py Spectrogram)."""
    return torch.hann_window(win_length)


def htk_mel_banks(n_mels: int, fft_len: int, sample_rate: int = 16000, f_min: float = 0.0,
                  f_max: float = None) -> torch.Tensor:
    """(n_mels, fft_len/2 + 1) triangular filters on the HTK scale 2595*log10(1 + f/700), norm=None: the transpose of
    ``melscale_fbanks`` as ``MelScale`` builds it (ta: functional/functional.py:492-588).  The reference never forwards
    its ``sr`` argument to MelSpectrogram (ref: lid/audio_processor.py:91-101), so the scale is always 16 kHz."""
    n_freqs = fft_len // 2 + 1
    if f_max is None:
        f_max = float(sample_rate // 2)
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    falling = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    rising = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(falling, rising))     # (n_freqs, n_mels)
    return fb.t().contiguous()


def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """Hann-windowed sinc FIR bank of ``torchaudio.transforms.Resample(orig_freq, new_freq)`` with its defaults, built
    with the same torch expressions in float64 and cast to float32 (ta: functional/functional.py
    _get_sinc_resample_kernel).  Returns ``(kernel (new/g, 2*width + orig/g) float32, width)``, g = gcd."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(orig, new)
    base_freq *= rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=None)[:, None, None] / new + idx
    t *= base_freq
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    scale = base_freq / orig
    kernels = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t)
    kernels *= window * scale
    return kernels.to(dtype=torch.float32)[:, 0, :].contiguous(), width


def windowed_dft_bank(n_fft: int):
    """The wav2vec-exp FBank's transform as a GEMM bank: rows 2 b / 2 b + 1 are hann(n_fft)[k] * cos / -sin(2 pi b k / n_fft),
    b = 0 .. n_fft / 2 (ref: wav2vec-exp/s3prl_model.py:192-196: F.spectrogram with torch.hann_window(n_fft), periodic).
    Built in float64 and rounded once.  The row count is padded with zero rows to what the tensor-core kernel tiles:
    a multiple of 32 with a divisor that is a multiple of 32 in [128, 256] (or the count itself when <= 256)."""
    n_bins = n_fft // 2 + 1
    k = torch.arange(n_fft, dtype=torch.float64)
    w = torch.hann_window(n_fft, dtype=torch.float64)
    b = torch.arange(n_bins, dtype=torch.float64).unsqueeze(1)
    ang = 2.0 * math.pi * torch.remainder(b * k, float(n_fft)) / n_fft
    rows = torch.stack([w * torch.cos(ang), -w * torch.sin(ang)], 1).reshape(2 * n_bins, n_fft)
    need = 2 * n_bins
    nw = (need + 31) // 32 * 32
    while nw > 256 and not any(nw % c == 0 for c in range(128, 257, 32)):
        nw += 32
    bank = torch.zeros((nw, n_fft), dtype=torch.float32)
    bank[:need] = rows.to(torch.float32)
    return bank.contiguous(), nw
