"""On-device replacement of the reference's ``DataProcessor`` (ref: lid/ConformerLangModel.py:131-178): polyphase
sinc resampling of 22.05 / 44.1 kHz input to the models' 16 kHz, i.e. ``torchaudio.transforms.Resample(orig, 16000)``
with its defaults, as one CUDA kernel over a packed batch of utterances."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import torch

from . import _lib, tables


class Resampler:
    def __init__(self, orig_freq: int, new_freq: int = 16000, device=None):
        self.lib = _lib.load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("speech_lid_b200.Resampler needs a CUDA device (sm_100a); there is no CPU path")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.orig_freq, self.new_freq = int(orig_freq), int(new_freq)
        self.identity = self.orig_freq == self.new_freq
        self.handle = C.c_void_p()
        if not self.identity:
            kernel, width = tables.sinc_resample_kernel(self.orig_freq, self.new_freq)
            self.kernel, self.width = kernel, width
            with torch.cuda.device(self.device):
                _lib.check(self.lib.lidfe_resampler_create(C.byref(self.handle), self.orig_freq, self.new_freq,
                                                            kernel.data_ptr(), kernel.shape[1], width))

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and self.handle.value:
                self.lib.lidfe_resampler_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass

    def out_len(self, n_in: int) -> int:
        """ceil(new * n / orig) (ta: functional/functional.py _apply_sinc_resample_kernel)."""
        return int(n_in) if self.identity else int(self.lib.lidfe_resample_out_len(self.handle, int(n_in)))

    def resample_list(self, wavs: Sequence[torch.Tensor], out_lens: Sequence[int] = None) -> List[torch.Tensor]:
        """Every 1-D waveform resampled on its own (equal to ``Resample`` applied to it alone, and -- because the FIR
        only ever sees zeros past an utterance's end -- to the rows of ``Resample`` applied to the zero-padded
        batch).  ``out_lens`` may ask for fewer samples than ``out_len(n)``."""
        wavs = [w.reshape(-1) for w in wavs]
        if self.identity:
            return [w.to(self.device) for w in wavs]
        n_in = [int(w.numel()) for w in wavs]
        n_out = [self.out_len(n) for n in n_in]
        if out_lens is not None:
            n_out = [min(int(a), b) for a, b in zip(out_lens, n_out)]
        in_off, pos = [], 0
        for n in n_in:
            in_off.append(pos)
            pos += n
        out_off, opos = [], 0
        for n in n_out:
            out_off.append(opos)
            opos += (n + 3) // 4 * 4                                # 16-byte aligned starts: ready for the front-end's TMA
        with torch.cuda.device(self.device):
            packed = torch.cat([w.to(self.device, torch.float32) for w in wavs]) if wavs else torch.empty(0, device=self.device)
            out = torch.zeros(max(opos, 1), dtype=torch.float32, device=self.device)
            tab = torch.tensor([in_off, n_in, out_off, n_out], dtype=torch.int64, device=self.device)
            _lib.check(self.lib.lidfe_resample(self.handle, len(wavs), packed.data_ptr(), tab[0].data_ptr(), tab[1].data_ptr(),
                                               out.data_ptr(), tab[2].data_ptr(), tab[3].data_ptr(), max(n_out),
                                               torch.cuda.current_stream(self.device).cuda_stream))
        return [out[o:o + n] for o, n in zip(out_off, n_out)]

    def data_processor(self, x: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        """``DataProcessor.forward(x, sample_rate=orig_freq)`` (ref: lid/ConformerLangModel.py:146-169): the reference
        pads the batch to its longest utterance, resamples, and crops utterance i to
        ``int(len_i / longest * resampled_padded_length)`` samples -- same lengths here."""
        if self.identity:
            return [w.to(self.device) for w in x]
        longest = max(int(w.shape[-1]) for w in x)
        padded_out = self.out_len(longest)
        lens = [int(int(w.shape[-1]) / longest * padded_out) for w in x]
        return self.resample_list(x, out_lens=lens)
