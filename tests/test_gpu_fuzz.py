"""Seeded random sweep of the CUDA path against the oracle: ragged batches, padded and packed layouts, masks, every
CMVN mode, float and int16 input.  Complements the hand-picked cases of test_gpu_parity.py."""
import random

import pytest
import torch

from oracle import frontend_oracle as O
from test_gpu_parity import _check_fbank

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lid():
    import speech_lid_b200 as m
    return m


def _lengths(rng, n):
    out = []
    for _ in range(n):
        kind = rng.random()
        if kind < 0.2:
            out.append(rng.randint(400, 2000))                 # 1 .. 11 frames
        elif kind < 0.9:
            out.append(rng.randint(2000, 64000))
        else:
            out.append(rng.randint(64000, 200000))
    return out


@pytest.mark.parametrize("seed", range(6))
def test_random_batches(lid, seed):
    rng = random.Random(1234 + seed)
    fe = lid.FrontEnd(n_mels=80)
    lens = _lengths(rng, rng.randint(1, 12))
    gen = O.synth_speechlike if seed % 2 else O.synth_noise
    wavs = [gen(n, 5000 + 17 * seed + i) for i, n in enumerate(lens)]
    frames = [O.kaldi_num_frames(n) for n in lens]
    want = [O.kaldi_fbank(w) for w in wavs]
    padded = bool(seed & 2)
    n_masks = rng.choice([0, 1, 2, 3])
    torch.manual_seed(seed)
    masks = lid.draw_masks(frames, 80, 0.05, 27, n_masks) if n_masks else None

    def rows(out, i):                                              # utterance i's valid rows in either layout
        if padded:
            return out[i, :frames[i]]
        start = sum(frames[:i])
        return out[start:start + frames[i]]

    # plain features
    raw, percents = fe.featurize(wavs, padded=padded)
    raw = raw.cpu()
    if padded:
        assert raw.shape == (len(lens), max(frames), 80)
        assert torch.allclose(percents.cpu(), torch.tensor([f / max(frames) for f in frames]), atol=1e-6)
        for i, f in enumerate(frames):
            assert torch.all(raw[i, f:] == 0)
    else:
        assert raw.shape == (sum(frames), 80)
    for i in range(len(lens)):
        _check_fbank(rows(raw, i), want[i], "seed %d utt %d" % (seed, i), all_bins=bool(seed % 2))

    # masks in the epilogue: exactly the plain features with the table's boxes zeroed
    if masks is not None:
        got, _ = fe.featurize(wavs, masks=masks, padded=padded)
        got = got.cpu()
        for i in range(len(lens)):
            b = [tuple(int(v) for v in masks[i, q]) for q in range(masks.shape[1])]
            ref = O.apply_mask_bounds(rows(raw, i).T.unsqueeze(0), b)[0].T
            assert torch.equal(rows(got, i), ref), "seed %d utt %d masks" % (seed, i)

    # per-utterance CMVN (+ masks) on the device's own raw features
    got, _ = fe.featurize(wavs, masks=masks, cmvn="utt", padded=padded)
    got = got.cpu()
    for i in range(len(lens)):
        if frames[i] < 2:
            assert torch.isnan(rows(got, i)).all() or masks is not None      # std of one frame
            continue
        ref = O.cmvn_per_utt(rows(raw, i))
        if masks is not None:
            b = [tuple(int(v) for v in masks[i, q]) for q in range(masks.shape[1])]
            ref = O.apply_mask_bounds(ref.T.unsqueeze(0), b)[0].T
        assert torch.allclose(rows(got, i), ref, rtol=1e-5, atol=2e-5), (seed, i, (rows(got, i) - ref).abs().max())   # fp32 mean, 1/std

    # global CMVN in two passes: accumulate, finalise on the host side, apply
    if padded:
        plan = fe.make_plan(lens, padded=True)
        packed = fe.pack([w.squeeze(0).cuda() for w in wavs], plan)
        stats = torch.zeros(2 * 80 + 1, dtype=torch.float64, device="cuda")
        feats = fe.featurize_packed(packed, plan, cmvn="global_accum", stats_out=stats)
        assert int(stats[-1].item()) == sum(frames)
        fe.cmvn_apply(feats, plan, stats, masks=masks.cuda() if masks is not None else None)
        allrows = torch.cat([rows(raw, i) for i in range(len(lens))]).double()
        mean, std = allrows.mean(0), allrows.std(0, unbiased=True)
        for i in range(len(lens)):
            ref = ((rows(raw, i).double() - mean) / (std + 1e-9)).float()
            if masks is not None:
                b = [tuple(int(v) for v in masks[i, q]) for q in range(masks.shape[1])]
                ref = O.apply_mask_bounds(ref.T.unsqueeze(0), b)[0].T
            assert torch.allclose(feats[i, :frames[i]].cpu(), ref, rtol=1e-5, atol=2e-5), (seed, i)


@pytest.mark.parametrize("seed", range(2))
def test_random_batches_int16(lid, seed):
    rng = random.Random(99 + seed)
    fe = lid.FrontEnd(n_mels=80, in_dtype=torch.int16, in_scale=1.0 / 32768.0)
    lens = _lengths(rng, rng.randint(2, 8))
    g = torch.Generator().manual_seed(700 + seed)
    pcm = [(torch.randn(n, generator=g) * 4000).clamp(-32768, 32767).to(torch.int16) for n in lens]
    feats, _ = fe.featurize(pcm)
    feats = feats.cpu()
    for i, p in enumerate(pcm):
        want = O.kaldi_fbank((p.to(torch.float32) * (1.0 / 32768.0)).unsqueeze(0))
        _check_fbank(feats[i, :want.shape[0]], want, "int16 seed %d utt %d" % (seed, i))
