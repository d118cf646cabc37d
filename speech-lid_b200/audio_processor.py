"""Drop-in for ``lid/audio_processor.py`` of kouyt5/speech-lid: same function names, argument names,
defaults and error behaviour, computed by the sm_100a kernels behind ``liblidfe.so``.

    ref: lid/audio_processor.py:8-38    wav2mel
    ref: lid/audio_processor.py:108-115 normalize_wav
    ref: lid/audio_processor.py:118-122 read_audio
    ref: lid/audio_processor.py:125-167 wav_augment (dither + 0.97 pre-emphasis; sox/reverb raise)
    ref: lid/audio_processor.py:198-228 spectrogram_augment (time/frequency masking; t_stretch raises)

Tensors may live on the host (as in the reference, which runs in DataLoader workers) or on the GPU; the
result comes back on the input's device.  Host tensors cost a PCIe round trip per call -- the batched
``FrontEnd.featurize`` is the entry meant for training loops.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from .frontend import FrontEnd
from .specaug import draw_masks

_frontends: Dict[Tuple, FrontEnd] = {}

# The single-utterance wrappers below are latency-bound (a PCIe round trip and two launches per call), so both wav2mel
# branches run in the precise arithmetic mode (``FrontEnd(precise=True)``: float64 internally, one rounding): it costs
# nothing measurable here and the result is never further from the fp64 truth than the reference's own.  Set to False
# for the fast fp32 kernels (what the batched ``FrontEnd.featurize`` uses by default).
PRECISE_DROPIN = True


def _frontend(n_mels: int, sr: int, win_length: float, hop_length: float, device, kind: str = "kaldi",
              pad: int = 0) -> FrontEnd:
    precise = bool(PRECISE_DROPIN)
    key = (int(n_mels), int(sr), float(win_length), float(hop_length), str(device), kind, int(pad), precise)
    fe = _frontends.get(key)
    if fe is None:
        fe = FrontEnd(n_mels=n_mels, sr=sr, win_length=win_length, hop_length=hop_length, preemph=1.0,
                      device=device, kind=kind, pad=pad, precise=precise)
        _frontends[key] = fe
    return fe


def _device_of(x: torch.Tensor):
    return x.device if x.is_cuda else torch.device("cuda:%d" % torch.cuda.current_device())


def wav2mel(x, use_kaildi: bool = False, win_length: float = 0.025, hop_length: float = 0.01,
            n_mels: int = 80, n_fft: int = 512, pad: int = 0, sr: int = 16000):
    """x (1, T) -> (1, n_mels, T').  ``use_kaildi=True`` is the Kaldi-fbank branch (``n_fft`` and ``pad`` are ignored by
    it, as in the reference, ref: lid/audio_processor.py:27-29); the default is MelSpectrogram + AmplitudeToDB(top_db=80)
    (ref: lid/audio_processor.py:72-105)."""
    if x.dim() != 2:
        raise ValueError("wav2mel expects a (channel, time) tensor")
    if use_kaildi:
        fe = _frontend(n_mels, sr, win_length, hop_length, _device_of(x))
    else:
        # default branch: MelSpectrogram(n_fft, win, hop, pad, center, reflect, power 2) + AmplitudeToDB(top_db=80)
        if n_fft != 512:
            raise NotImplementedError("speech_lid_b200: only n_fft=512 is built (the value every config uses)")
        if x.shape[0] != 1:
            raise ValueError("the MelSpectrogram branch is built for mono (1, T) input")
        fe = _frontend(n_mels, sr, win_length, hop_length, _device_of(x), kind="melspec_db", pad=pad)
        if fe.num_frames(x.shape[-1]) <= 0:      # torch.stft: reflect padding needs more samples than it mirrors
            raise RuntimeError("Argument #4: Padding size should be less than the corresponding input dimension")
    plan = fe.cached_plan([x.shape[-1]], padded=False)        # kaldi branch: AssertionError when T < 400
    packed = fe.pack([x], plan)
    feats = fe.featurize_packed(packed, plan)                  # (T', n_mels)
    out = feats.transpose(0, 1).unsqueeze(0)                   # ref: lid/audio_processor.py:63-65
    return out if x.is_cuda else out.cpu()


def spectrogram_augment(spec, sr: int = 16000, n_mels: int = 80, hop_length: float = 0.01,
                        t_mask: float = 0.05, f_mask: float = 27, mask_times: int = 0,
                        t_stretch: bool = False):
    """spec (1, n_mels, T) -> masked copy.  Mask bounds are drawn from the global CPU generator in the
    reference's order; the fill value is 0.0."""
    if t_stretch:
        raise NotImplementedError("t_stretch (phase-vocoder TimeStretch, ref: lid/audio_processor.py:220-224) "
                                  "is off the hot path and not provided")
    if spec.dim() < 2:
        raise ValueError("Spectrogram must have at least two dimensions (time and frequency) "
                         "({} given).".format(spec.dim()))
    if mask_times <= 0:
        return spec
    T, F = int(spec.shape[-1]), int(spec.shape[-2])
    masks = draw_masks([T], F, t_mask, f_mask, mask_times)
    fe = _frontend(F if 4 <= F <= 80 else 80, 16000, 0.025, 0.01, _device_of(spec))
    if F != fe.n_out or spec.dim() != 3 or spec.shape[0] != 1:
        raise ValueError("spectrogram_augment expects (1, n_mels<=80, T)")
    rows = spec.to(fe.device)[0].transpose(0, 1).contiguous()      # (T, F) rows as the kernels lay them out
    plan = fe.cached_plan([400 + 160 * (T - 1)], padded=False)     # a plan with exactly T frames
    fe.mask_apply(rows, plan, masks)
    out = rows.transpose(0, 1).unsqueeze(0)
    return out if spec.is_cuda else out.cpu()


def normalize_wav(wav: torch.Tensor):
    """(wav - mean) / (std + 1e-6) over time, unbiased std.   wav: (1, T)"""
    fe = _frontend(80, 16000, 0.025, 0.01, _device_of(wav))
    return _wave(fe, wav, normalize=True)


def _wave(fe: FrontEnd, wav: torch.Tensor, normalize=False, dither=0.0, noise=None, preemph=0.0):
    """(C, T) waveform through the wave-stage kernel; every channel is an utterance of its own, as the reference's
    ``torch.std_mean(wav, dim=-1)`` / element-wise stages treat the rows (ref: lid/audio_processor.py:108-134)."""
    if wav.dim() != 2 or wav.shape[0] < 1:
        raise ValueError("expected a (C, T) waveform")
    C, n = int(wav.shape[0]), int(wav.shape[-1])
    if normalize and C > 1 and C != n:
        # the reference subtracts a (C,) mean from a (C, T) tensor: only mono broadcasts (ref: lid/audio_processor.py:112-113)
        raise RuntimeError("The size of tensor a (%d) must match the size of tensor b (%d) at non-singleton dimension 1" % (n, C))
    if n < fe.frame_len:
        # (the plan machinery is built around frames; the reference's element-wise stages would accept it)
        raise AssertionError("waveform shorter than one frame")
    plan = fe.cached_plan([n] * C, padded=False)
    packed = fe.pack([wav[c].to(torch.float32) for c in range(C)], plan)
    nz = None
    if noise is not None:
        nz = fe.pack([noise[c].to(torch.float32) for c in range(C)], plan)
    out = fe.wave_stages(packed, plan, normalize=normalize, dither=dither, noise=nz, preemph=preemph)
    out = torch.stack([out[o:o + n] for o in plan.offsets], 0)
    return out if wav.is_cuda else out.cpu()


def read_audio(audio_path: str, normalize: bool = True):
    import torchaudio
    wav, sr = torchaudio.load(audio_path)
    if normalize:
        wav = normalize_wav(wav)
    return wav, sr


def wav_augment(wav, sr, speed_shift: bool = False, pitch_shift: bool = False, reverb: bool = False):
    """Dither ``wav += 1e-5 * U[0,1)`` (in place, like the reference) then 0.97 pre-emphasis keeping sample 0.
    The uniform noise is drawn on the host from the global CPU generator, as ``torch.rand_like`` does for the
    reference's CPU tensors, so results match it for the same RNG state."""
    if speed_shift or pitch_shift:
        raise NotImplementedError("sox speed/pitch effects (ref: lid/audio_processor.py:135-154) are off the hot path")
    if reverb:
        raise NotImplementedError("WavAugment reverb (ref: lid/audio_processor.py:155-163) is off the hot path")
    noise = torch.rand(wav.shape, dtype=wav.dtype)
    fe = _frontend(80, 16000, 0.025, 0.01, _device_of(wav))
    dithered = _wave(fe, wav, dither=1e-5, noise=noise)
    wav.copy_(dithered)                       # the reference mutates its argument (``wav += ...``, :129)
    return _wave(fe, wav, preemph=0.97), sr
