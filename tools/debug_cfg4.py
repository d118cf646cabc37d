import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import speech_lid_b200 as lid
from oracle import frontend_oracle as O
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(3)
U = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
lengths = torch.randint(16000, 320001, (U,), generator=g).tolist()
fe = lid.FrontEnd(n_mels=80, device=dev)
plan = fe.make_plan(lengths, padded=False)
gd = torch.Generator(device=dev).manual_seed(3000)
packed = torch.randn(plan.total_samples, device=dev, generator=gd)
host = {}
for j in range(6):
    w = O.synth_noise(lengths[j], 7000 + j); host[j] = w
    packed[plan.offsets[j]:plan.offsets[j] + lengths[j]] = w[0].to(dev)
torch.manual_seed(99)
masks = lid.draw_masks(plan.frames, 80, 0.05, 27, 2).to(dev)
raw = fe.featurize_packed(packed, plan)                       # mode none
row = 0
for j in range(6):
    T = plan.frames[j]
    ref = O.kaldi_fbank(host[j]); got = raw[row:row+T].cpu()
    d = (got - ref).abs()
    print("utt", j, "T", T, "row", row, "raw max abs", float(d.max()), "bad rows", (d.max(1).values > 1e-2).nonzero().flatten()[:10].tolist())
    row += T
stats = torch.zeros(161, dtype=torch.float64, device=dev)
raw3 = fe.featurize_packed(packed, plan, cmvn="global_accum", stats_out=stats)
print("mode3 equals mode0:", torch.equal(raw3, raw), "count", stats[160].item(), sum(plan.frames))
ref_stats = torch.cat([raw.double().sum(0), (raw.double()**2).sum(0)])
print("stats rel err", float(((stats[:160] - ref_stats).abs() / ref_stats.abs()).max()))
y = fe.cmvn_apply(raw3.clone(), plan, stats, masks=masks)
mean, std = lid.finalize_stats(stats.cpu())
row = 0
for j in range(6):
    T = plan.frames[j]
    ref = O.cmvn_apply(raw[row:row+T].cpu(), mean, std)
    b = [tuple(int(v) for v in masks[j, q]) for q in range(2)]
    ref = O.apply_mask_bounds(ref.T.unsqueeze(0), b)[0].T
    got = y[row:row+T].cpu()
    mism = ((got == 0) != (ref == 0))
    print("utt", j, b, "mask mismatches", int(mism.sum()), mism.nonzero()[:6].tolist(), "max abs", float((got-ref).abs().max()))
    row += T
