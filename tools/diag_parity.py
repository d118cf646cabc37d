#!/usr/bin/env python
"""Print error statistics of the GPU fbank against the fp32 oracle and the fp64 truth (GPU box)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import speech_lid_b200 as lid
from oracle import frontend_oracle as O

fe = lid.FrontEnd(n_mels=80)
res = {}
for kind, gen in (("noise", O.synth_noise), ("speech", O.synth_speechlike)):
    for seed in range(3):
        x = gen(128000, 10 + seed)
        got = fe.featurize([x])[0][0].cpu()
        ref = O.kaldi_fbank(x)
        tru = O.truth64_fbank(x)
        d = (got - ref).abs()
        e_gpu = (got.double() - tru).abs().max(0).values
        e_ref = (ref.double() - tru).abs().max(0).values
        r = dict(norm_rel=float(d.max() / ref.abs().max()), max_abs=float(d.max()),
                 max_abs_bins0_2=float(d[:, :3].max()), max_abs_bins3_9=float(d[:, 3:10].max()),
                 max_abs_bins10p=float(d[:, 10:].max()),
                 frac_rtol1e4=float(torch.isclose(got, ref, rtol=1e-4, atol=0).float().mean()),
                 gpu_vs_truth=[float(e_gpu[:3].max()), float(e_gpu[3:10].max()), float(e_gpu[10:].max())],
                 ref_vs_truth=[float(e_ref[:3].max()), float(e_ref[3:10].max()), float(e_ref[10:].max())],
                 worst_ratio=float((e_gpu / (e_ref + 1e-12)).max()))
        res["%s_%d" % (kind, seed)] = r
        print(kind, seed, json.dumps(r))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/diag_parity.json", "w"), indent=1)
# RMS view of the same comparison
for kind, gen in (("noise", O.synth_noise),):
    for seed in range(3):
        x = gen(128000, 10 + seed)
        got = fe.featurize([x])[0][0].cpu().double()
        ref = O.kaldi_fbank(x).double(); tru = O.truth64_fbank(x)
        for a, b in ((0, 3), (3, 10), (10, 80)):
            print("rms", kind, seed, (a, b), "gpu %.3g ref %.3g gpu-vs-ref %.3g" % (
                (got - tru)[:, a:b].pow(2).mean().sqrt(), (ref - tru)[:, a:b].pow(2).mean().sqrt(),
                (got - ref)[:, a:b].pow(2).mean().sqrt()))
