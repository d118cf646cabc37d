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
    B, M = len(frames), max(int(mask_times), 0)
    out = torch.zeros((B, M, 4), dtype=torch.int32)
    if B == 0 or M == 0:
        return out
    # One draw of all uniforms instead of two torch.rand(1) per mask: the CPU generator hands out float32 uniforms
    # serially, so rand(K) is the same K numbers K calls of rand(1) return (tests/test_host_logic.py pins this against
    # the serial loop).  A 256-utterance batch costs 0.2 ms instead of 30 ms of per-call overhead.
    T = torch.tensor([int(t) for t in frames], dtype=torch.int64)
    t_param = (T.double() * t_mask).long()          # int(T * t_mask): Python float arithmetic = float64, truncation
    f_param = torch.full((B,), int(f_mask) if float(f_mask) == int(f_mask) else -1, dtype=torch.int64)
    if (f_param < 0).any():
        return _draw_masks_serial(frames, n_mels, t_mask, f_mask, mask_times, generator)
    # per (utterance, mask, axis): drawn only when the parameter is >= 1       ta: functional/functional.py:930-931
    param = torch.stack([t_param, f_param], dim=1)[:, None, :].expand(B, M, 2)            # [B, M, 2]
    axis_len = torch.stack([T, torch.full_like(T, int(n_mels))], dim=1)[:, None, :].expand(B, M, 2)
    live = param >= 1
    n_live = int(live.sum())
    if n_live == 0:
        return out
    u = torch.rand(2 * n_live, generator=generator)                                         # value, min_value, ...
    p32 = param[live].to(torch.float32)
    value = u[0::2] * p32
    min_value = u[1::2] * (axis_len[live].to(torch.float32) - value)
    start = min_value.long()
    end = start + value.long()
    if bool(((end - start) >= param[live]).any()):   # ta: functional/functional.py:948-949
        raise ValueError("Number of columns to be masked should be less than mask_param")
    se = torch.zeros((B, M, 2, 2), dtype=torch.int32)
    se[..., 0][live] = start.to(torch.int32)
    se[..., 1][live] = end.to(torch.int32)
    return se.reshape(B, M, 4).contiguous()


def _draw_masks_serial(frames: Sequence[int], n_mels: int = 80, t_mask: float = 0.05, f_mask: float = 27,
                       mask_times: int = 0, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """The literal loop (two ``torch.rand(1)`` per mask, like ``mask_along_axis``): the definition ``draw_masks`` is
    tested against, and the path for a non-integer ``f_mask``."""
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
