"""SpecAugment mask parameters, drawn on the host exactly like the reference draws them.

ref: lid/audio_processor.py:225-227 -- ``for i in range(mask_times): TimeMasking(int(T*t_mask));
FrequencyMasking(f_mask)``; ta: functional/functional.py:885-958 ``mask_along_axis``: two
``torch.rand(1)`` per mask on the CPU default generator (``value`` then ``min_value``), none when the
mask parameter is < 1, bounds ``[long(min_value), long(min_value) + long(value))``, fill value 0.0.

Drawing the integers here (same generator, same order) and applying them in the kernel epilogue makes
the device result bit-exact to the reference given the same RNG state.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch


def draw_masks(frames: Sequence[int], n_mels: int = 80, t_mask: float = 0.05, f_mask: float = 27,
               mask_times: int = 0, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """int32 [B, mask_times, 4] = (t0, t1, f0, f1) per utterance, utterances in batch order.

    ``generator=None`` consumes the global CPU generator, i.e. the very stream the reference would consume
    when it processes the same utterances in the same order after the same ``torch.manual_seed``."""
    out = torch.zeros((len(frames), max(mask_times, 0), 4), dtype=torch.int32)
    for b, T in enumerate(frames):
        for i in range(mask_times):
            for col, (axis_len, param) in enumerate(((int(T), int(T * t_mask)), (int(n_mels), f_mask))):
                if param < 1:          # ta: functional/functional.py:930-931 -> returns before any draw
                    continue
                value = torch.rand(1, generator=generator) * param
                min_value = torch.rand(1, generator=generator) * (axis_len - value)
                start = int(min_value.long())
                end = start + int(value.long())
                if end - start >= param:   # ta: functional/functional.py:948-949
                    raise ValueError("Number of columns to be masked should be less than mask_param")
                out[b, i, 2 * col] = start
                out[b, i, 2 * col + 1] = end
    return out
