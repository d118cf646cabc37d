#!/usr/bin/env python
"""cfg5, GPU half: featurise a seeded batch with the CUDA front-end and save waveforms + features so
that tools/cfg5_consumer_check.py can feed them to the reference's Conformer consumer in the build container
(/root/reference does not exist on the GPU box).  Output: gpurun_out/cfg5_feats.npz (a 16-utterance subset of cfg5's
128 x 8 s so that the CPU forward of the reference model stays in the minutes)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import speech_lid_b200 as lid

B, N, SEED = 16, 128000, 4


def main():
    g = torch.Generator().manual_seed(SEED)
    wavs = []
    for _ in range(B):
        x = torch.randn(1, N, generator=g)
        mean, std = x.mean(), x.std()
        wavs.append(((x - mean) / (std + 1e-6)).squeeze(0))          # normalize_wav (ref: lid/audio_processor.py:108-115)
    fe = lid.FrontEnd(n_mels=80)
    feats, percents = fe.featurize([w.cuda() for w in wavs], cmvn="none")
    torch.cuda.synchronize()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "cfg5_feats.npz"), wavs=torch.stack(wavs).numpy(),
                        feats=feats.cpu().numpy(), percents=percents.cpu().numpy())
    print("cfg5 features", tuple(feats.shape), "launches", lid.load_library().lidfe_launch_count())


if __name__ == "__main__":
    main()
