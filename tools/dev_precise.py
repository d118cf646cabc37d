#!/usr/bin/env python
"""Time the precise (float64) mode against the fast fp32 kernels on cfg2 (256 x 8 s), CUDA events, device-resident input.
Run on the GPU box: python tools/dev_precise.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import speech_lid_b200 as lid


def time_mode(fe, packed, plan, cmvn, masks, iters=20):
    out = fe.featurize_packed(packed, plan, cmvn=cmvn, masks=masks)
    for _ in range(3):
        fe.featurize_packed(packed, plan, out=out, cmvn=cmvn, masks=masks)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fe.featurize_packed(packed, plan, out=out, cmvn=cmvn, masks=masks)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def main():
    B, n = 256, 128000
    g = torch.Generator(device="cuda").manual_seed(1)
    res = {}
    for name, kw in (("fast", {}), ("precise", dict(precise=True))):
        fe = lid.FrontEnd(n_mels=80, **kw)
        plan = fe.make_plan([n] * B, padded=True)
        packed = torch.randn(plan.total_samples, device="cuda", generator=g)
        frames = [int(f) for f in plan.frames] if hasattr(plan, "frames") else [798] * B
        torch.manual_seed(0)
        masks = lid.draw_masks(frames, 80, 0.05, 27, 2)
        res[name] = dict(none_us=time_mode(fe, packed, plan, "none", None), utt_masks_us=time_mode(fe, packed, plan, "utt", masks))
        print(name, json.dumps(res[name]), flush=True)
    fem = lid.FrontEnd(n_mels=80, n_ceps=40, precise=True)
    plan = fem.make_plan([64000] * 512, padded=True)
    packed = torch.randn(plan.total_samples, device="cuda", generator=g)
    res["precise_mfcc_cfg3"] = dict(none_us=time_mode(fem, packed, plan, "none", None))
    print("precise_mfcc_cfg3", json.dumps(res["precise_mfcc_cfg3"]), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/dev_precise.json", "w"), indent=1)


if __name__ == "__main__":
    main()
