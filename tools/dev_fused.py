#!/usr/bin/env python
"""Development: per-utterance CMVN with the second stage inside the warp kernel (LIDFE_WFUSED=1, utterance groups
LIDFE_WGROUPS) against the two-launch path, on cfg2-like and ragged batches, repeated launches on one plan (launch
parity).  The statistics are summed by atomics in another order, so equality is to ~1 ulp, not bitwise.  Also times both.
Run under `timeout`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import speech_lid_b200 as lid

dev = torch.device("cuda:0")


def make(fused, groups):
    os.environ["LIDFE_WFUSED"] = str(fused)
    os.environ["LIDFE_WGROUPS"] = str(groups)
    return lid.FrontEnd(n_mels=80)


def case(name, lens, padded, reps=4):
    g = torch.Generator().manual_seed(len(lens))
    wavs = [torch.randn(n, generator=g) for n in lens]
    ref = None
    for fused, groups in ((0, 1), (1, 1), (1, 4), (1, 8)):
        fe = make(fused, groups)
        plan = fe.make_plan(lens, padded=padded)
        packed = fe.pack(wavs, plan)
        torch.manual_seed(3)
        masks = lid.draw_masks(plan.frames, 80, 0.05, 27, 2).to(dev)
        outs = []
        for r in range(reps):
            out = fe.featurize_packed(packed, plan, masks=masks if r % 2 == 0 else None, cmvn="utt")
            torch.cuda.synchronize()
            outs.append(out.clone())
        if ref is None:
            ref = outs
            print("%-16s two-launch reference: %s, max|x| %.3f" % (name, tuple(outs[0].shape), float(torch.nan_to_num(outs[0]).abs().max())), flush=True)
            continue
        # (an utterance of one frame has no standard deviation: NaN rows, as torch.std of one sample)
        worst = max(float(torch.nan_to_num(a - b, nan=0.0).abs().max()) for a, b in zip(outs, ref))
        same = min(float(((a == b) | (torch.isnan(a) & torch.isnan(b))).float().mean()) for a, b in zip(outs, ref))
        nan = any(not torch.equal(torch.isnan(a), torch.isnan(b)) for a, b in zip(outs, ref))
        print("%-16s fused groups=%d: max|diff| %.3g, bit-equal share %.5f %s" % (name, groups, worst, same, "NaN MISMATCH" if nan else ""), flush=True)
        assert worst < 1e-4 and not nan, (name, groups, worst)


def timing():
    B, N = 256, 128000
    g = torch.Generator(device=dev).manual_seed(1)
    ins = [torch.randn(B * N, device=dev, generator=g) for _ in range(3)]
    for fused, groups in ((0, 1), (1, 1), (1, 2), (1, 3), (1, 4), (1, 6), (1, 8)):
        fe = make(fused, groups)
        plan = fe.make_plan([N] * B, padded=True)
        outs = [torch.empty(B, plan.t_max, 80, device=dev) for _ in range(3)]
        torch.manual_seed(1234)
        masks = lid.draw_masks(plan.frames, 80, 0.05, 27, 2).to(dev)
        for mode in ("utt", "none"):
            kw = dict(cmvn=mode, masks=masks)
            for i in range(5):
                fe.featurize_packed(ins[i % 3], plan, out=outs[i % 3], **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 30
            for i in range(n):
                fe.featurize_packed(ins[i % 3], plan, out=outs[i % 3], **kw)
            e1.record()
            torch.cuda.synchronize()
            print("fused=%d groups=%d mode=%-5s %.1f us/step" % (fused, groups, mode, e0.elapsed_time(e1) / n * 1e3), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "check"):
        case("small_ragged", [16000, 4000, 24000, 8560, 400, 559, 560, 64000, 1040, 720], True)
        case("small_packed", [16000, 4000, 24000, 8560, 400, 559, 560, 64000, 1040, 720], False)
        case("cfg2_like", [128000] * 256, True, reps=3)
        gl = torch.Generator().manual_seed(9)
        case("ragged_256", torch.randint(16000, 320001, (256,), generator=gl).tolist(), True, reps=3)
        case("one_frame_utts", [400] * 40 + [128000] * 8 + [401] * 3, False)
    if what in ("all", "time"):
        timing()
