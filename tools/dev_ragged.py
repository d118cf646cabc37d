#!/usr/bin/env python
"""Development: where a ragged DeviceCollate step (256 utterances of 1-20 s, a new length signature every step) spends
its time -- plan, mask draw, host packing + H2D, kernels, D2H read.  Wall clock with a synchronise after every stage."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import speech_lid_b200 as lid

dev = torch.device("cuda:0")
fe = lid.FrontEnd(n_mels=80)
g = torch.Generator().manual_seed(77)
batches = []
for _ in range(4):
    lens = torch.randint(16000, 320001, (256,), generator=g).tolist()
    batches.append([torch.randn(n, generator=g) for n in lens])


def sync():
    torch.cuda.synchronize(dev)


acc = {}
for it in range(12):
    wavs = batches[it % 4]
    lens = [int(w.shape[-1]) for w in wavs]
    t = [time.perf_counter()]
    plan = fe.make_plan(lens, padded=True); sync(); t.append(time.perf_counter())
    masks = lid.draw_masks(plan.frames, 80, 0.05, 27, 2); t.append(time.perf_counter())
    packed = fe.pack(wavs, plan); sync(); t.append(time.perf_counter())
    out = fe.featurize_packed(packed, plan, masks=masks.to(dev), cmvn="utt"); sync(); t.append(time.perf_counter())
    float(out[0, 0, 0]); t.append(time.perf_counter())
    plan.close()
    if it >= 4:
        for k, a, b in zip(("plan", "masks", "pack+h2d", "kernels", "d2h"), t[:-1], t[1:]):
            acc[k] = acc.get(k, 0.0) + (b - a) * 1e3 / 8
audio = sum(sum(int(w.shape[-1]) for w in b) for b in batches) / 4 / 16000
print("audio-s per step %.0f, MB per step %.0f, pack threads %d" % (audio, audio * 16000 * 4 / 1e6, fe.pack_threads))
print("  ".join("%s %.2f ms" % kv for kv in acc.items()), " total %.2f ms -> %.0f audio-s/s" % (sum(acc.values()), audio / sum(acc.values()) * 1e3))
for thr in (1, 2, 4, 8, 16, 32):
    fe.pack_threads = thr
    plan = fe.make_plan([int(w.shape[-1]) for w in batches[0]], padded=True)
    for _ in range(2):
        fe.pack(batches[0], plan); sync()
    t0 = time.perf_counter()
    for _ in range(4):
        fe.pack(batches[0], plan); sync()
    print("pack threads %2d: %.2f ms" % (thr, (time.perf_counter() - t0) / 4 * 1e3))
