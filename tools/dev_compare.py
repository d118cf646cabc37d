#!/usr/bin/env python
"""Development: the warp-autonomous kernel against fbank_kernel (LIDFE_WARP_KERNEL=0) on ragged and full-size batches,
every cmvn mode.  Same arithmetic per frame, so raw features must be bit-equal; statistics are summed in another order."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

def run():
    import speech_lid_b200 as lid
    dev = torch.device("cuda:0")
    res = {}
    g = torch.Generator().manual_seed(5)
    for name, lens, padded in (("ragged_padded", [16000, 4000, 24000, 8560, 400, 559, 560, 64000, 1040, 720], True),
                               ("ragged_packed", [16000, 4000, 24000, 8560, 400, 559, 560, 64000, 1040, 720], False),
                               ("cfg2_like", [128000] * 64, True)):
        for kw in (dict(), dict(in_dtype=torch.int16, in_scale=1.0 / 32768), dict(n_mels=40), dict(preemph=0.97), dict(n_ceps=40)):
            fe = lid.FrontEnd(**kw)
            if kw.get("in_dtype") == torch.int16:
                wavs = [(torch.randn(n, generator=g) * 3000).to(torch.int16) for n in lens]
            else:
                wavs = [torch.randn(n, generator=g) for n in lens]
            plan = fe.make_plan(lens, padded=padded)
            packed = fe.pack(wavs, plan)
            torch.manual_seed(3)
            masks = lid.draw_masks(plan.frames, fe.n_out, 0.05, 27 if fe.n_out == 80 else 13, 2).to(dev)
            for mode in ("none", "utt", "global_accum", "global_apply"):
                for use_masks in (False, True):
                    k = dict(cmvn=mode)
                    if use_masks and mode != "global_accum":
                        k["masks"] = masks
                    stats = torch.zeros(2 * fe.n_out + 1, dtype=torch.float64, device=dev)
                    if mode == "global_accum":
                        k["stats_out"] = stats
                    if mode == "global_apply":
                        st = torch.zeros(2 * fe.n_out + 1, dtype=torch.float64, device=dev)
                        fe.featurize_packed(packed, plan, cmvn="global_accum", stats_out=st)
                        k["stats_in"] = st
                    print("case", name, kw, mode, use_masks, file=sys.stderr, flush=True)
                    out = fe.featurize_packed(packed, plan, **k)
                    out2 = fe.featurize_packed(packed, plan, **k)        # twice: the workspace must be back at rest
                    torch.cuda.synchronize()
                    key = "%s|%s|%s|%d" % (name, sorted(kw.items(), key=str), mode, use_masks)
                    res[key] = (out.cpu(), stats.cpu() / 2 if mode == "global_accum" else None, bool(torch.equal(out, out2) or mode == "utt"))
    return res

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] != "-v":
        torch.save(run(), sys.argv[1])
        sys.exit(0)
    outs = []
    for flag in ("1", "0"):
        f = "/tmp/cmp_%s.pt" % flag
        subprocess.check_call([sys.executable, __file__, f], env=dict(os.environ, LIDFE_WARP_KERNEL=flag))
        outs.append(torch.load(f))
    a, b = outs
    bad = 0
    for k in a:
        x, sx, rep = a[k]
        y, sy, _ = b[k]
        nan = bool(torch.isnan(x).any()) != bool(torch.isnan(y).any())
        x, y = torch.nan_to_num(x, nan=12345.0), torch.nan_to_num(y, nan=12345.0)
        same = torch.equal(x, y)
        d = float((x - y).abs().max())
        ds = float(((sx - sy).abs() / sy.abs().clamp_min(1)).max()) if sx is not None else 0.0
        ok = (same or ("utt" in k and d < 2e-6) or ("global_apply" in k and d < 2e-6)) and ds < 2e-7 and rep and not nan
        bad += not ok
        if not ok or "-v" in sys.argv:
            print("%-90s equal=%s maxdiff=%.3g stats_rel=%.3g repeat=%s" % (k, same, d, ds, rep))
    print("compare: %d cases, %d bad" % (len(a), bad))
