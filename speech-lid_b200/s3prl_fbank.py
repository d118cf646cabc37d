"""On-device replacement of the ``FBank`` module of the reference's wav2vec experiment
(ref: wav2vec-exp/s3prl_model.py:174-204; used by ``MutuGLU`` with ``n_fft = 320 * 2``, :44, :133):

    F.spectrogram(x, pad=0, hann_window(n_fft), n_fft, hop=n_fft // 2, win=n_fft, power=2, center=False)
    -> F.melscale_fbanks(n_fft // 2 + 1, 0, 8000, 16000, fbank_size, norm=None, "htk")
    -> F.amplitude_to_DB(10, amin=1e-10, db_multiplier=0, top_db=None) -> (spec - mean) / (std + 1e-9)

with ONE mean / unbiased std over the whole spectrogram of an utterance.  ``n_fft = 640 = 5 * 2^7`` is no size of the
fused 512-point FFT kernel, so the transform runs as a GEMM -- frames x windowed DFT basis -- on the tensor-core windowed
GEMM of the resampler (tcgen05, 3 x TF32: csrc/lidfe_resample_tc.cuh), followed by the power / mel / dB / statistics
kernel and the scalar normalisation (csrc/lidfe_stft_fbank.cuh)."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import torch

from . import _lib, tables


class S3prlFBank:
    def __init__(self, fbank_size: int = 80, n_fft: int = 640, device=None):
        self.lib = _lib.load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("speech_lid_b200.S3prlFBank needs a CUDA device (sm_100a); there is no CPU path")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.n_mels, self.n_fft, self.hop = int(fbank_size), int(n_fft), int(n_fft) // 2
        if self.n_fft < 4 or self.n_fft % 2 or self.n_fft > 2048:
            raise ValueError("n_fft must be even, 4 <= n_fft <= 2048")
        self.n_bins = self.n_fft // 2 + 1
        bank, self.nw = tables.windowed_dft_bank(self.n_fft)               # [nw][n_fft] fp32, rows >= 2 n_bins are zero
        fb = tables.htk_mel_banks(self.n_mels, self.n_fft, 16000, 0.0, 8000.0)      # [n_mels][n_bins]
        nz = fb != 0
        lo = torch.where(nz.any(1), nz.float().argmax(1), torch.zeros(self.n_mels, dtype=torch.long))
        hi = torch.where(nz.any(1), self.n_bins - 1 - nz.flip(1).float().argmax(1), -torch.ones(self.n_mels, dtype=torch.long))
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.lidfe_wgemm_create(C.byref(self.handle), self.hop, self.nw, bank.data_ptr(), self.n_fft, 0))
            self.melT = fb.contiguous().to(self.device)
            self.mel_lo = lo.to(torch.int32).to(self.device)
            self.mel_hi = hi.to(torch.int32).to(self.device)

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and self.handle.value:
                self.lib.lidfe_resampler_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass

    def num_frames(self, n_samples: int) -> int:
        """torch.stft(center=False): 1 + (N - n_fft) // hop"""
        return 1 + (int(n_samples) - self.n_fft) // self.hop if n_samples >= self.n_fft else 0

    def forward_list(self, wavs: Sequence[torch.Tensor], normalize: bool = True) -> List[torch.Tensor]:
        """list of (T_i,) waveforms -> list of (frames_i, fbank_size) feature matrices on the device, i.e. what
        ``MutuGLU.forward`` builds with ``self.fbank(x).transpose(0, 1)`` (ref: wav2vec-exp/s3prl_model.py:148)."""
        wavs = [w.reshape(-1) for w in wavs]
        n_in = [int(w.numel()) for w in wavs]
        frames = [self.num_frames(n) for n in n_in]
        if min(frames) <= 0:
            raise RuntimeError("S3prlFBank: an utterance is shorter than n_fft (torch.stft raises as well)")
        B = len(wavs)
        in_off, g_off, rows, pos, gpos, rpos = [], [], [], 0, 0, 0
        for n, f in zip(n_in, frames):
            in_off.append(pos)
            pos += n
            g_off.append(gpos)
            gpos += f * self.nw
            rows.append(rpos)
            rpos += f
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device).cuda_stream
            packed = torch.cat([w.to(self.device, torch.float32) for w in wavs])
            g = torch.empty(gpos, dtype=torch.float32, device=self.device)
            out = torch.empty((rpos, self.n_mels), dtype=torch.float32, device=self.device)
            tab = torch.tensor([in_off, n_in, g_off, [f * self.nw for f in frames], frames, rows], dtype=torch.int64,
                               device=self.device)
            stats = torch.empty(2 * B, dtype=torch.float64, device=self.device)
            _lib.check(self.lib.lidfe_resample(self.handle, B, packed.data_ptr(), tab[0].data_ptr(), tab[1].data_ptr(),
                                               g.data_ptr(), tab[2].data_ptr(), tab[3].data_ptr(), max(frames) * self.nw, st))
            _lib.check(self.lib.lidfe_stft_mel_db(g.data_ptr(), tab[2].data_ptr(), tab[4].data_ptr(), B, max(frames), self.nw,
                                                  self.melT.data_ptr(), self.mel_lo.data_ptr(), self.mel_hi.data_ptr(),
                                                  self.n_bins, self.n_mels, 1e-10, out.data_ptr(), tab[5].data_ptr(),
                                                  stats.data_ptr(), int(bool(normalize)), st))
        return [out[r:r + f] for r, f in zip(rows, frames)]

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """``FBank.forward(x)``: (T,) -> (1, fbank_size, frames) like the reference's module (on ``x``'s device)."""
        y = self.forward_list([x])[0].transpose(0, 1).unsqueeze(0)
        return y if x.is_cuda else y.cpu()
