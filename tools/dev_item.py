import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes, numpy as np
import speech_lid_b200 as lid
dev = torch.device("cuda:0")
B, N = int(sys.argv[1]), 128000
fe = lid.FrontEnd(n_mels=80)
plan = fe.make_plan([N] * B, padded=True)
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(B * N, device=dev, generator=g)
out = torch.empty(B, plan.t_max, 80, device=dev)
for i in range(4):
    fe.featurize_packed(x, plan, out=out, cmvn="utt")
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (8 * 592))()
fe.lib.lidfe_dbg_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
fe.lib.lidfe_dbg_read(buf, 8 * 592)
a = np.array(buf[:]).reshape(592, 8)
a = a[a[:, 2] > 0]
t0 = a[:, 0].min()
print("B=%d ctas=%d spans-done max %.1f us, end max %.1f us, items(warp0) sum %d" % (B, len(a), (a[:, 1].max() - t0) / 1e3, (a[:, 2].max() - t0) / 1e3, a[:, 4].sum()))
