"""world_size-2 gloo test of the only exchange step on the path: utterance sharding + the global-CMVN statistics
all-reduce (2*n_out+1 doubles).  Features come from the CPU oracle here; on the GPU box the same host code runs
over NCCL with device tensors (bench.py / tests -m gpu)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import speech_lid_b200 as lid
from oracle import frontend_oracle as O

LENGTHS = [16000, 5000, 9000, 12000, 700, 24000, 3000, 8000]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shards = lid.lpt_partition(LENGTHS, world)                 # computed identically on every rank
        mine = [O.kaldi_fbank(O.synth_noise(LENGTHS[i], 50 + i)) for i in shards[rank]]
        stats = O.cmvn_stats(mine) if mine else torch.zeros(161, dtype=torch.float64)
        lid.allreduce_stats(stats)                                 # the one collective on the path
        mean, std = lid.finalize_stats(stats)
        normed = {i: O.cmvn_apply(f, mean, std) for i, f in zip(shards[rank], mine)}
        torch.save(dict(stats=stats, normed=normed, shard=shards[rank]), os.path.join(out_dir, "rank%d.pt" % rank))
    finally:
        dist.destroy_process_group()


def test_global_cmvn_allreduce_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    feats = [O.kaldi_fbank(O.synth_noise(n, 50 + i)) for i, n in enumerate(LENGTHS)]
    want_stats = O.cmvn_stats(feats)
    mean, std = O.cmvn_finalize(want_stats)
    seen = []
    for r in range(world):
        d = torch.load(os.path.join(str(tmp_path), "rank%d.pt" % r))
        assert torch.allclose(d["stats"], want_stats, rtol=1e-12, atol=1e-9)      # every rank holds the global sums
        assert d["stats"][160].item() == sum(f.shape[0] for f in feats)
        for i, y in d["normed"].items():
            assert torch.allclose(y, O.cmvn_apply(feats[i], mean, std), rtol=1e-6, atol=1e-6)
        seen += d["shard"]
    assert sorted(seen) == list(range(len(LENGTHS)))                             # every utterance on exactly one rank
    fm, fs = lid.finalize_stats(want_stats)
    assert torch.allclose(fm, mean) and torch.allclose(fs, std)


def test_allreduce_is_identity_without_process_group():
    s = torch.arange(161, dtype=torch.float64)
    assert torch.equal(lid.allreduce_stats(s.clone()), s)
