#!/usr/bin/env python
"""A few small launches of every kernel and mode, for compute-sanitizer (memcheck / racecheck) on the GPU box:
   compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import speech_lid_b200 as lid

torch.manual_seed(0)
lens = [16000, 4000, 24000, 8560, 400, 700]
wavs = [torch.randn(n) for n in lens]
wavs = [(w - w.mean()) / (w.std() + 1e-6) for w in wavs]
frames = [1 + (n - 400) // 160 for n in lens]
masks = lid.draw_masks(frames, 80, 0.05, 27, 2)

fe = lid.FrontEnd(n_mels=80)
fe.featurize(wavs)
fe.featurize(wavs, masks=masks)
fe.featurize(wavs, masks=masks, cmvn="utt")
fe.featurize(wavs, padded=False)
plan = fe.make_plan(lens, padded=True)
packed = fe.pack([w.cuda() for w in wavs], plan)
stats = torch.zeros(161, dtype=torch.float64, device="cuda")
feats = fe.featurize_packed(packed, plan, cmvn="global_accum", stats_out=stats)
fe.cmvn_apply(feats, plan, stats, masks=masks.cuda())
fe.featurize_packed(packed, plan, cmvn="global_apply", stats_in=stats, masks=masks.cuda())
fe.wave_stages(packed, plan, normalize=True)

mf = lid.FrontEnd(n_mels=80, n_ceps=40)
mf.featurize(wavs)                                   # two-kernel MFCC path
mf.featurize(wavs, masks=masks, cmvn="utt")          # in-kernel DCT + statistics
lid.FrontEnd(n_mels=23, n_ceps=13, preemph=0.97).featurize(wavs)
lid.FrontEnd(n_mels=40, preemph=0.97).featurize(wavs)

ms = lid.FrontEnd(kind="melspec_db", pad=16)
ms.featurize(wavs)
ms.featurize(wavs, masks=masks)

i16 = lid.FrontEnd(n_mels=80, in_dtype=torch.int16, in_scale=1.0 / 32768.0)
i16.featurize([(w * 3000).clamp(-32768, 32767).to(torch.int16) for w in wavs])

pr = lid.FrontEnd(n_mels=80, precise=True)           # float64 kernel (lidfe_fbank_precise.cuh)
pr.featurize(wavs)
pr.featurize(wavs, masks=masks, cmvn="utt")
pr.featurize(wavs, padded=False)
pstats = torch.zeros(161, dtype=torch.float64, device="cuda")
pplan = pr.make_plan(lens, padded=False)
pr.featurize_packed(pr.pack(wavs, pplan), pplan, cmvn="global_accum", stats_out=pstats)
lid.FrontEnd(n_mels=80, n_ceps=40, precise=True).featurize(wavs)
lid.FrontEnd(n_mels=23, n_ceps=13, preemph=0.97, precise=True).featurize(wavs)
lid.FrontEnd(kind="melspec_db", pad=16, precise=True).featurize(wavs, masks=masks)
torch.cuda.synchronize()
print("sanitize_case ok, launches", lid.load_library().lidfe_launch_count())
