#!/usr/bin/env python
"""bench.py -- audio-seconds/second of the speech-lid front-end hot path on B200.

Workload (BASELINE.json configs[1]): 80-dim Kaldi fbank + SpecAugment (2 time + 2 frequency masks) +
per-utterance CMVN on a batch of 256 x 8-s 16 kHz utterances per GPU (synthetic, seeded).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One JSON line on rank 0.  `value` = whole-job audio-s/s with inputs resident in HBM (weak scaling: every rank
processes its own 256-utterance batch; utterances are independent, no data-path collective).  `e2e` = the same
metric through the public host API with pinned HOST buffers (H2D of the waveforms and D2H of the features inside
the timed region).  `roofline` describes the dominant kernel (the fused fbank kernel) from CUDA-event timings of
every launch inside the timed region.  `cpu_baseline` = the oracle port of the reference path on the box's host
cores (rank 0, N=1 only).  --impl reference times that CPU path as the reference arm.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_UTTS = 256
N_SAMPLES = 128000
SR = 16000
N_MELS = 80
T_MASK, F_MASK, MASK_TIMES = 0.05, 27, 2
AUDIO_S_PER_BATCH = B_UTTS * N_SAMPLES / SR          # 2048
ALG_BYTES_PER_FRAME = 160 * 4 + N_MELS * 4             # 640 B in + 320 B out per 10-ms frame
ALG_FLOP_PER_FRAME = 15.0e3
METRIC = "audio-sec/sec 80-dim fbank+SpecAugment+CMVN"
WORKLOAD = "cfg2: 80-dim kaldi fbank + SpecAugment(t_mask=0.05,f_mask=27,x2) + per-utterance CMVN, 256 x 8-s utterances per GPU"


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, one worker process per host core
# ------------------------------------------------------------------------------------------------
def _cpu_worker_init():
    import torch
    torch.set_num_threads(1)


def _cpu_worker(args):
    """Reference path for a share of the batch, exactly as MergedDataset.__getitem__ runs it per utterance
    (ref: lid/raw_datasets.py:270-305): wav2mel(use_kaildi=True) -> spectrogram_augment -> [our CMVN]."""
    import torch
    from oracle import frontend_oracle as O
    seed, n_utts, n_samples, reps = args
    g = torch.Generator().manual_seed(seed)
    wavs = [O.normalize_wav(torch.randn(1, n_samples, generator=g)) for _ in range(n_utts)]
    gen = torch.Generator().manual_seed(seed + 1)
    times = []
    acc = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        for w in wavs:
            spec = O.wav2mel_kaldi(w)                                            # (1, 80, T)
            feat = O.cmvn_per_utt(spec[0].T)                                     # (T, 80)
            spec = O.spectrogram_augment(feat.T.unsqueeze(0), T_MASK, F_MASK, MASK_TIMES, generator=gen)
            acc += float(spec[0, 0, 0])
        times.append(time.perf_counter() - t0)
    return times, acc


def run_cpu_arm(steps, warmup, n_utts_total=B_UTTS, n_samples=N_SAMPLES):
    """Returns (audio_s_per_s, seconds_per_step, cores).  Every step is the full 256 x 8-s batch split over all
    host cores; per-step time = the slowest worker's time for that step (workers run concurrently)."""
    import multiprocessing as mp
    cores = max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    workers = min(cores, n_utts_total)
    share = [n_utts_total // workers + (1 if i < n_utts_total % workers else 0) for i in range(workers)]
    ctx = mp.get_context("spawn")
    t_wall0 = time.perf_counter()
    with ctx.Pool(workers, initializer=_cpu_worker_init) as pool:
        res = pool.map(_cpu_worker, [(1000 + i, share[i], n_samples, steps + warmup) for i in range(workers)])
    wall = time.perf_counter() - t_wall0
    per_step = [max(r[0][s] for r in res) for s in range(warmup, warmup + steps)]
    sec = sum(per_step) / len(per_step)
    audio_s = n_utts_total * n_samples / SR
    return audio_s / sec, sec, workers, wall


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU while the timed region runs (pynvml; nvidia-smi fallback)."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._h = None

    def _decode(self, mask):
        nv = self._nv
        names = {"hw_slowdown": "nvmlClocksThrottleReasonHwSlowdown",
                 "hw_thermal_slowdown": "nvmlClocksThrottleReasonHwThermalSlowdown",
                 "sw_thermal_slowdown": "nvmlClocksThrottleReasonSwThermalSlowdown",
                 "sw_power_cap": "nvmlClocksThrottleReasonSwPowerCap",
                 "hw_power_brake": "nvmlClocksThrottleReasonHwPowerBrakeSlowdown"}
        for k, attr in names.items():
            bit = getattr(nv, attr, None)
            if bit is not None and (mask & bit):
                self.reasons.add(k)

    def sample_once(self):
        if self._h is None:
            return
        try:
            self.samples.append(self._nv.nvmlDeviceGetClockInfo(self._h, self._nv.NVML_CLOCK_SM))
            self._decode(self._nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
        except Exception:
            pass

    def _loop(self):
        while not self._stop.is_set():
            self.sample_once()
            self._stop.wait(0.002)

    def start(self):
        if self._h is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        if self._h is None:      # fallback: one nvidia-smi query right after the region
            try:
                import subprocess
                out = subprocess.check_output(
                    ["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                     "--format=csv,noheader,nounits"], timeout=20).decode().strip().split(",")
                self.samples.append(int(out[0]))
                self.max_mhz = int(out[1])
            except Exception:
                pass
        return {"sm_mhz": (statistics.median(self.samples) if self.samples else None),
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def fp32_peak_probe(torch, device):
    """FFMA micro-benchmark through torch (plumbing only; not on the hot path): a long chain of fused multiply-adds
    is not expressible in eager torch, so the FP32 ceiling is taken from SM count x 128 lanes x 2 x clock."""
    props = torch.cuda.get_device_properties(device)
    return props.multi_processor_count


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=dev)

    import speech_lid_b200 as lid
    fe = lid.FrontEnd(n_mels=N_MELS, device=dev)
    lib = lid.load_library()
    plan = fe.make_plan([N_SAMPLES] * B_UTTS, padded=True)
    frames_per_step = plan.total_frames

    # resident inputs: NBUF rotating sets so that no step finds its waveforms or its output lines in L2
    NBUF = 3
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    ins, outs = [], []
    for _ in range(NBUF):
        w = torch.randn(B_UTTS, N_SAMPLES, device=dev, generator=g)
        w = (w - w.mean(1, keepdim=True)) / (w.std(1, keepdim=True) + 1e-6)      # normalize_wav'ed, as the reference feeds it
        ins.append(w.reshape(-1).contiguous())
        outs.append(torch.empty(B_UTTS, plan.t_max, N_MELS, device=dev))
    torch.manual_seed(1234)
    masks = lid.draw_masks(plan.frames, N_MELS, T_MASK, F_MASK, MASK_TIMES).to(dev)

    def step(i):
        fe.featurize_packed(ins[i % NBUF], plan, out=outs[i % NBUF], masks=masks, cmvn="utt")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(i)
    barrier()

    sampler = ClockSampler(torch.cuda._get_nvml_device_index(dev) if hasattr(torch.cuda, "_get_nvml_device_index") else local_rank)
    prof_stride = 4 if args.steps >= 16 else 1     # sample the kernel timing: an event record between kernels costs ~us
    fe.profile_begin(args.steps, stride=prof_stride)
    launches0 = lib.lidfe_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    ev0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    ev1.record()
    sampler.sample_once()
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    launches = lib.lidfe_launch_count() - launches0
    kernel_ms = fe.profile_end()

    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * AUDIO_S_PER_BATCH / (ms_step * 1e-3)

    # ---- e2e: public host API, pinned host buffers, H2D + kernels + D2H per step --------------------------
    host_in = torch.empty(B_UTTS * N_SAMPLES, dtype=torch.float32).pin_memory()
    host_in.copy_(ins[0].cpu())
    host_out = torch.empty(B_UTTS, plan.t_max, N_MELS, dtype=torch.float32).pin_memory()
    host_masks = masks.cpu().pin_memory()

    e2e_chunks = int(os.environ.get("LIDFE_E2E_CHUNKS", "8"))

    def e2e_step():
        fe.featurize_host(host_in, plan, host_out, masks=host_masks, cmvn="utt", chunks=e2e_chunks)

    e2e_steps = max(3, min(args.steps, 20))
    E2E_BLOCKS = 3

    def timed_blocks(step_fn):
        """Seconds per step: median over E2E_BLOCKS blocks of e2e_steps steps (the PCIe link is shared with whatever else
        runs on the host; one disturbed block would otherwise decide the number)."""
        for _ in range(2):
            step_fn()
        secs = []
        for _ in range(E2E_BLOCKS):
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                step_fn()
            torch.cuda.synchronize(dev)
            secs.append((time.perf_counter() - t0) / e2e_steps)
        return sorted(secs)[len(secs) // 2]

    e2e_sec = timed_blocks(e2e_step)
    t = torch.tensor([e2e_sec], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * AUDIO_S_PER_BATCH / float(t.item())
    h2d = host_in.numel() * 4 + host_masks.numel() * 4
    d2h = host_out.numel() * 4

    # informational: the same step fed with raw int16 PCM (2 B/sample over PCIe; 1/32768 scaling + normalize_wav on the
    # device, row f2).  NOT the headline e2e: the reference's wav2mel boundary takes float32 waveforms.
    host_pcm = (ins[0].cpu() * 3000.0).clamp(-32768, 32767).to(torch.int16).pin_memory()

    def pcm_step():
        fe.featurize_host(host_pcm, plan, host_out, masks=host_masks, cmvn="utt", chunks=e2e_chunks)

    pcm_sec = timed_blocks(pcm_step)
    t = torch.tensor([pcm_sec], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pcm_value = world * AUDIO_S_PER_BATCH / float(t.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak_src = "fallback"
    hbm_peak = 6650.0
    if os.path.exists(peaks_path):
        try:
            hbm_peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "measured"
        except Exception:
            pass
    k_ms = statistics.mean(kernel_ms) if kernel_ms else float("nan")
    alg_bytes = frames_per_step * ALG_BYTES_PER_FRAME
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    clk = (clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965) * 1e6
    fp32_peak_at_clock = sms * 128 * 2 * clk / 1e12
    fp32_peak_boost = sms * 128 * 2 * (clocks.get("sm_max_mhz") or 1965) * 1e6 / 1e12
    fp32_achieved = frames_per_step * ALG_FLOP_PER_FRAME / (k_ms * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("fbank_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "fbank_kernel<float,false>", "achieved": round(achieved, 1),
                "peak": hbm_peak, "peak_source": peak_src, "unit": "GB/s", "frac": round(achieved / hbm_peak, 4),
                "traffic": traffic, "kernel_ms": round(k_ms, 5), "kernel_share_of_step": round(k_ms / ms_step, 3),
                "kernel_timing": "CUDA event pair around every %d-th launch of the timed region (%d samples)" % (prof_stride, len(kernel_ms)),
                "alg_bytes_per_launch": alg_bytes,
                "fp32": {"achieved_tflops": round(fp32_achieved, 2), "peak_tflops_at_sampled_clock": round(fp32_peak_at_clock, 1),
                         "peak_tflops_at_max_clock": round(fp32_peak_boost, 1),
                         "frac_at_max_clock": round(fp32_achieved / fp32_peak_boost, 4),
                         "note": "algorithmic 15.0 kflop/frame (SURVEY.md 8d); peak = SMs x 128 lanes x 2 x clock"}}

    cpu = None
    if world == 1 and not args.no_cpu:
        v, sec, cores, wall = run_cpu_arm(steps=1, warmup=1)
        cpu = {"value": round(v, 1), "unit": "audio-s/s", "cores": cores, "kind": "port",
               "sample": "1 timed pass (after 1 warm-up) of the full 256 x 8-s batch, oracle port of wav2mel(use_kaildi=True)"
                         "+spectrogram_augment+CMVN, one process per core, torch threads=1; %.1f s wall incl. start-up" % wall}

    line = {"metric": METRIC, "value": round(value, 1), "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 5), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "utterances_per_gpu": B_UTTS, "samples_per_utterance": N_SAMPLES,
                       "frames_per_gpu_step": frames_per_step, "parallelism": "utterance-sharded x%d" % world,
                       "l2": "3 rotating input/output sets (196 MB per step > 126 MB L2)"},
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 1), "unit": "audio-s/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "blocks": E2E_BLOCKS,
                    "agg": "median block", "pipeline_chunks": e2e_chunks},
            "e2e_int16_pcm": {"value": round(pcm_value, 1), "unit": "audio-s/s", "h2d_bytes_per_step": host_pcm.numel() * 2,
                              "d2h_bytes_per_step": d2h,
                              "note": "informational: host ships int16 PCM, scaling + normalize_wav on the device"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps = max(1, args.steps)
    warmup = max(0, args.warmup)
    v, sec, cores, wall = run_cpu_arm(steps=steps, warmup=warmup)
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 1), "unit": "audio-s/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": round(sec * 1e3, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "utterances_per_gpu": B_UTTS, "samples_per_utterance": N_SAMPLES},
            "cpu_baseline": {"value": round(v, 1), "unit": "audio-s/s", "cores": cores, "kind": "port",
                             "sample": "every step = the full 256 x 8-s batch split over %d single-thread worker "
                                       "processes (oracle port: same torch CPU ops as the reference's torchaudio path)" % cores},
            "e2e": {"value": round(v, 1), "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
