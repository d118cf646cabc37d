#!/usr/bin/env python
"""Development: how much of cmvn_apply_kernel's time is HBM reads?  Times the apply pass on a 256 x 798 x 80 batch whose rows
were (a) just written (as much of the 65 MB as the L2 keeps) and (b) pushed out of L2 by a 400 MB memset in between."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import speech_lid_b200 as lid

fe = lid.FrontEnd(n_mels=80)
plan = fe.make_plan([128000] * 256, padded=True)
src = torch.randn(256, plan.t_max, 80, device="cuda")
feats = torch.empty_like(src)
stats = torch.zeros(161, dtype=torch.float64, device="cuda")
stats[:80] = 0.1 * 204288; stats[80:160] = 1.5 * 204288; stats[160] = 204288
junk = torch.empty(100_000_000, device="cuda")
for label, flush in (("rows just written", False), ("rows flushed from L2", True)):
    ts = []
    for _ in range(12):
        feats.copy_(src)
        if flush:
            junk.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fe.cmvn_apply(feats, plan, stats)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:])
    print("%-22s apply pass %.1f us (median of %d)" % (label, ts[len(ts) // 2], len(ts)))
