import os, sys
sys.path.insert(0, '/root/repo')
import torch, speech_lid_b200 as lid
fb = lid.FrontEnd(n_mels=80)
plan = fb.make_plan([128000]*256, padded=True)
w = torch.randn(256*128000, device='cuda')
out = torch.empty(256, 798, 80, device='cuda')
masks = lid.draw_masks([798]*256, 80, 0.05, 27, 2)
for i in range(6):
    fb.featurize_packed(w, plan, out=out, masks=masks)
torch.cuda.synchronize()
