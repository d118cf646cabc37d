#!/usr/bin/env python
"""Where does the GPU fbank differ most from the oracle on the ragged SpecAugment test batch? (GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import speech_lid_b200 as lid
from oracle import frontend_oracle as O

fe = lid.FrontEnd(n_mels=80)
lens = [128000, 48000, 3300, 16000]
wavs = [O.synth_noise(n, 700 + i) for i, n in enumerate(lens)]
frames = [O.kaldi_num_frames(n) for n in lens]
got, _ = fe.featurize(wavs)
got = got.cpu()
solo = [fe.featurize([w])[0][0].cpu() for w in wavs]
for i, w in enumerate(wavs):
    ref = O.kaldi_fbank(w)
    tru = O.truth64_fbank(w)
    g = got[i, :frames[i]]
    d = (g - ref).abs()
    idx = torch.nonzero(d == d.max())[0].tolist()
    print("utt", i, "frames", frames[i], "max|gpu-ref| %.3e at frame %d bin %d" % (d.max(), idx[0], idx[1]),
          "gpu %.6f ref %.6f truth %.6f" % (g[idx[0], idx[1]], ref[idx[0], idx[1]], tru[idx[0], idx[1]]),
          "batch==solo", torch.equal(g, solo[i]),
          "bins>=3 max %.3e" % d[:, 3:].max(), "scale %.3f" % ref.abs().max())
    top = torch.topk(d[:, 3:].flatten(), 5)
    for v, k in zip(top.values.tolist(), top.indices.tolist()):
        f, b = divmod(k, 77)
        print("    frame %d bin %d err %.3e gpu %.5f ref %.5f truth %.5f" % (f, b + 3, v, g[f, b + 3], ref[f, b + 3], tru[f, b + 3]))
