#!/usr/bin/env python
"""Development: does cutting the cfg2 batch into G utterance groups (fbank(g) -> apply(g) back to back, so that a group's
rows are still in L2 when its normalisation pass reads them) beat one fbank + one apply launch over the whole batch?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import speech_lid_b200 as lid

dev = torch.device("cuda:0")
B, n = 256, 128000
fe = lid.FrontEnd(n_mels=80)
g = torch.Generator(device=dev).manual_seed(1)
NBUF = 3
ins = [torch.randn(B * n, device=dev, generator=g) for _ in range(NBUF)]
for G in (1, 2, 4, 8):
    b = B // G
    plans = [fe.make_plan([n] * b, padded=True) for _ in range(G)]
    outs = [torch.empty(B, plans[0].t_max, 80, device=dev) for _ in range(NBUF)]
    torch.manual_seed(0)
    masks = lid.draw_masks(plans[0].frames * G, 80, 0.05, 27, 2).to(dev)

    def step(i):
        x, o = ins[i % NBUF], outs[i % NBUF]
        for k in range(G):
            fe.featurize_packed(x[k * b * n:(k + 1) * b * n], plans[k], out=o[k * b:(k + 1) * b], masks=masks[k * b:(k + 1) * b], cmvn="utt")

    for i in range(5):
        step(i)
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for i in range(100):
        step(i)
    e.record()
    torch.cuda.synchronize()
    print("groups %d: %.1f us per step" % (G, a.elapsed_time(e) * 10), flush=True)
