#!/usr/bin/env python
"""Development: host-side cost of a ragged DeviceCollate step with PINNED items (bench.py's e2e_ragged leg): host time
spent in each stage without synchronising in between, then the wait for the device."""
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
    batches.append([torch.randn(n, generator=g).pin_memory() for n in lens])
acc = {}
for it in range(16):
    wavs = batches[it % 4]
    lens = [int(w.shape[-1]) for w in wavs]
    t = [time.perf_counter()]
    frames = [fe.num_frames(n) for n in lens]; t.append(time.perf_counter())
    masks = lid.draw_masks(frames, 80, 0.05, 27, 2); t.append(time.perf_counter())
    plan = fe.make_plan(lens, padded=True); t.append(time.perf_counter())
    packed = fe.pack(wavs, plan); t.append(time.perf_counter())
    out = fe.featurize_packed(packed, plan, masks=masks, cmvn="utt"); t.append(time.perf_counter())
    float(out[0, 0, 0]); t.append(time.perf_counter())
    plan.close(); t.append(time.perf_counter())
    if it >= 4:
        for k, a, b in zip(("num_frames", "masks", "plan", "pack(issue)", "featurize(issue)", "wait", "close"), t[:-1], t[1:]):
            acc[k] = acc.get(k, 0.0) + (b - a) * 1e3 / 12
audio = sum(sum(int(w.shape[-1]) for w in b) for b in batches) / 4 / 16000
print("audio-s per step %.0f, MB per step %.0f" % (audio, audio * 16000 * 4 / 1e6))
print("  ".join("%s %.2f ms" % kv for kv in acc.items()), " total %.2f ms -> %.0f audio-s/s" % (sum(acc.values()), audio / sum(acc.values()) * 1e3))

# the product path: DeviceCollate (ships first, plans and draws the masks under the DMA)
collate = lid.DeviceCollate(fe, {"a": 0}, train=True, t_mask=0.05, f_mask=27, mask_times=2, cmvn="utt")
items = [[(w, torch.zeros(3, dtype=torch.long), "p", "a") for w in b] for b in batches]
for it in range(4):
    float(collate(items[it % 4])[0][0, 0, 0])
torch.cuda.synchronize()
t0 = time.perf_counter()
for it in range(12):
    float(collate(items[it % 4])[0][0, 0, 0])
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 12
print("DeviceCollate step %.2f ms -> %.0f audio-s/s" % (dt * 1e3, audio / dt))

# device-side timeline of one step: events around the copies and the kernels
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
accd = [0.0, 0.0, 0.0]
for it in range(8):
    wavs = batches[it % 4]
    lens = [int(w.shape[-1]) for w in wavs]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev[0].record()
    offsets, pos = [], 0
    for n in lens:
        offsets.append(pos); pos += (n + 3) // 4 * 4
    from speech_lid_b200.frontend import _Layout
    packed = fe.pack(wavs, _Layout(lengths=lens, offsets=offsets, total_samples=pos))
    ev[1].record()
    plan = fe.make_plan(lens, padded=True, offsets=offsets)
    masks = lid.draw_masks(plan.frames, 80, 0.05, 27, 2)
    ev[2].record()
    out = fe.featurize_packed(packed, plan, masks=masks, cmvn="utt")
    ev[3].record()
    torch.cuda.synchronize()
    host = (time.perf_counter() - t0) * 1e3
    if it >= 2:
        for k in range(3):
            accd[k] += ev[k].elapsed_time(ev[k + 1]) / 6
    plan.close()
print("device timeline: memset+copies %.2f ms, plan upload %.2f ms, mask upload + kernels %.2f ms; host wall of the last step %.2f ms" % (accd[0], accd[1], accd[2], host))
