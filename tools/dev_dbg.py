import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import speech_lid_b200 as lid
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N = int(sys.argv[2]) if len(sys.argv) > 2 else 128000
g = torch.Generator().manual_seed(5)
fe = lid.FrontEnd()
wavs = [torch.randn(N, generator=g) for _ in range(B)]
plan = fe.make_plan([N] * B, padded=True)
packed = fe.pack(wavs, plan)
stats = torch.zeros(161, dtype=torch.float64, device=dev)
out = fe.featurize_packed(packed, plan, cmvn="global_accum", stats_out=stats)
torch.cuda.synchronize()
o = out.double().reshape(-1, 80)
want = torch.cat([o.sum(0), (o * o).sum(0), torch.tensor([float(o.shape[0])], device=dev, dtype=torch.float64)])
rel = ((stats - want).abs() / want.abs().clamp_min(1))
print("wspans/spans", fe.lib.lidfe_plan_num_spans(plan.handle), "max rel", float(rel.max()), "count", float(stats[160]), "want", float(want[160]))
print("sum ratio dims 0..15", (stats[:16] / want[:16]).cpu().numpy().round(4))
print("sq ratio dims 0..15", (stats[80:96] / want[80:96]).cpu().numpy().round(4))
print("sum ratio dims 64..79", (stats[64:80] / want[64:80]).cpu().numpy().round(4))
