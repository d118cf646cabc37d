#!/usr/bin/env python
"""BASELINE config 4: variable-length (1-20 s) utterances packed by an offset table, sharded by utterance over the
GPUs of one box, global CMVN through ONE all-reduce of [sum_d, sumsq_d, count] (2*80+1 doubles).

    python tools/cfg4_global_cmvn.py [--utts 10000]                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/cfg4_global_cmvn.py --utts 10000

Every rank derives the same length list (seed 3) and the same LPT partition, generates its shard on the device
(seed 3000+rank), and runs  featurize(global_accum) -> all_reduce -> cmvn_apply.  Checks: (a) a CPU-generated subset
of each shard against the oracle; (b) the all-reduced statistics equal the sum of the per-rank ones; (c) the
normalised features have global mean 0 / unbiased std 1 per dim.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=10000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--check", type=int, default=8, help="utterances per rank checked against the CPU oracle")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import speech_lid_b200 as lid
    from oracle import frontend_oracle as O

    g = torch.Generator().manual_seed(3)
    lengths = torch.randint(16000, 320001, (args.utts,), generator=g).tolist()
    shards = lid.lpt_partition(lengths, world)
    mine = [lengths[i] for i in shards[rank]]
    fe = lid.FrontEnd(n_mels=80, device=dev)
    plan = fe.make_plan(mine, padded=False)
    gd = torch.Generator(device=dev).manual_seed(3000 + rank)
    packed = torch.randn(plan.total_samples, device=dev, generator=gd)
    # a few utterances come from the host so they can be checked against the oracle
    ncheck = min(args.check, len(mine))
    host = {}
    for j in range(ncheck):
        w = O.synth_noise(mine[j], 7000 + 100 * rank + j)
        host[j] = w
        packed[plan.offsets[j]:plan.offsets[j] + mine[j]] = w[0].to(dev)
    torch.manual_seed(99 + rank)
    masks = lid.draw_masks(plan.frames, 80, 0.05, 27, 2).to(dev)

    def step():
        stats = torch.zeros(161, dtype=torch.float64, device=dev)
        raw = fe.featurize_packed(packed, plan, cmvn="global_accum", stats_out=stats)
        local_stats = stats.clone()
        lid.allreduce_stats(stats)
        fe.cmvn_apply(raw, plan, stats, masks=masks)
        return raw, stats, local_stats

    feats, stats, local_stats = step()
    torch.cuda.synchronize(dev)

    # (a) parity of the checked utterances (undo nothing: compare against oracle chain with the same global stats)
    mean, std = lid.finalize_stats(stats.cpu())
    worst = 0.0
    row = 0
    for j in range(len(mine)):
        T = plan.frames[j]
        if j in host:
            ref = O.cmvn_apply(O.kaldi_fbank(host[j]), mean, std)
            b = [tuple(int(v) for v in masks[j, q]) for q in range(masks.shape[1])]
            ref = O.apply_mask_bounds(ref.T.unsqueeze(0), b)[0].T
            got = feats[row:row + T].cpu()
            # every masked position is exactly 0 (an unmasked value may also be 0: x == mean to fp32 precision)
            masked = torch.zeros_like(ref, dtype=torch.bool)
            for t0, t1, f0, f1 in b:
                masked[t0:t1, :] = True
                masked[:, f0:f1] = True
            assert torch.all(got[masked] == 0), "masked positions not zero (rank %d utt %d)" % (rank, j)
            worst = max(worst, float((got - ref).abs().max()))
        row += T
    assert worst < 5e-3, worst      # 1/std amplifies the fbank round-off of the three low bins
    # (b) all-reduce == sum of per-rank vectors
    if world > 1:
        gathered = [torch.zeros_like(local_stats) for _ in range(world)]
        dist.all_gather(gathered, local_stats)
        assert torch.allclose(torch.stack(gathered).sum(0), stats, rtol=1e-12, atol=1e-6)
    total_frames = int(stats[160].item())
    # (c) global moments of the (unmasked) normalised features
    raw2 = fe.featurize_packed(packed, plan, cmvn="global_apply", stats_in=stats)
    m = torch.cat([raw2.double().sum(0), (raw2.double() ** 2).sum(0)])
    if world > 1:
        dist.all_reduce(m)
    gm = m[:80] / total_frames
    gv = (m[80:] - total_frames * gm ** 2) / (total_frames - 1)
    assert gm.abs().max() < 1e-4 and (gv.sqrt() - 1).abs().max() < 1e-4, (gm.abs().max(), (gv.sqrt() - 1).abs().max())

    # timing: K passes of the whole path incl. the all-reduce, CUDA events, max over ranks
    for _ in range(2):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    audio_s = sum(lengths) / 16000.0
    if rank == 0:
        print(json.dumps({"config": "cfg4: %d utterances 1-20 s, packed, utterance-sharded x%d, global CMVN all-reduce" % (args.utts, world),
                          "n_gpus": world, "audio_s": round(audio_s, 1), "total_frames": total_frames,
                          "ms_per_pass": round(float(ms.item()), 3),
                          "audio_s_per_s": round(audio_s / (float(ms.item()) * 1e-3), 1),
                          "worst_abs_err_checked_utts": worst, "scaling": "strong"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
