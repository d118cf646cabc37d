#!/usr/bin/env python
"""Quick GPU sanity run used during kernel development: a few small batches through every mode, each step printed
(flush) so that a hang can be located.  Run under `timeout`."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import speech_lid_b200 as lid
from oracle import frontend_oracle as O


def say(*a):
    print(*a, flush=True)


def main():
    dev = torch.device("cuda:0")
    fe = lid.FrontEnd(n_mels=80)
    say("created")
    lens = [16000, 4000, 24000, 8560]
    wavs = [O.synth_noise(n, 40 + i) for i, n in enumerate(lens)]
    feats, pct = fe.featurize(wavs)
    torch.cuda.synchronize()
    say("none ok", float(feats.abs().max()))
    want = [O.kaldi_fbank(w) for w in wavs]
    for i, w in enumerate(want):
        say(" utt", i, "err", float((feats[i, :w.shape[0]].cpu() - w).abs().max()), "pad zero", bool((feats[i, w.shape[0]:] == 0).all()))
    torch.manual_seed(0)
    frames = [w.shape[0] for w in want]
    masks = lid.draw_masks(frames, 80, 0.05, 27, 2)
    y, _ = fe.featurize(wavs, masks=masks, cmvn="utt")
    torch.cuda.synchronize()
    say("utt ok")
    for i in range(len(wavs)):
        b = [tuple(int(v) for v in masks[i, q]) for q in range(masks.shape[1])]
        ref = O.apply_mask_bounds(O.cmvn_per_utt(feats[i, :frames[i]].cpu()).T.unsqueeze(0), b)[0].T
        say(" utt", i, "cmvn err", float((y[i, :frames[i]].cpu() - ref).abs().max()))
    # twice more: the workspace must be back at rest
    y2, _ = fe.featurize(wavs, masks=masks, cmvn="utt")
    torch.cuda.synchronize()
    say("utt again equal", bool(torch.equal(y, y2)))
    # bigger batch
    B, N = 64, 128000
    g = torch.Generator(device=dev).manual_seed(1)
    w = torch.randn(B, N, device=dev, generator=g)
    plan = fe.make_plan([N] * B, padded=True)
    out = torch.empty(B, plan.t_max, 80, device=dev)
    for mode in ("none", "utt"):
        t0 = time.time()
        fe.featurize_packed(w.reshape(-1), plan, out=out, cmvn=mode)
        torch.cuda.synchronize()
        say("big", mode, "ok %.3f s" % (time.time() - t0), float(out.abs().max()))
    if len(sys.argv) > 1:
        return
    fem = lid.FrontEnd(kind="melspec_db", pad=16)
    z, _ = fem.featurize(wavs)
    torch.cuda.synchronize()
    for i, wv in enumerate(wavs):
        ref = O.melspec_db(wv, pad=16)[0].T
        say(" melspec utt", i, "err", float((z[i, :ref.shape[0]].cpu() - ref).abs().max()))
    fec = lid.FrontEnd(n_mels=80, n_ceps=40)
    c, _ = fec.featurize(wavs)
    torch.cuda.synchronize()
    for i, wv in enumerate(wavs):
        ref = O.kaldi_mfcc(wv)
        say(" mfcc utt", i, "err", float((c[i, :ref.shape[0]].cpu() - ref).abs().max()))
    c2, _ = fec.featurize(wavs, cmvn="utt")
    torch.cuda.synchronize()
    say("mfcc utt ok", float(c2.abs().max()))


if __name__ == "__main__":
    main()
