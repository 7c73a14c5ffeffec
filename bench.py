#!/usr/bin/env python
"""Benchmark of the VBQ rate-distortion quantization step (BASELINE.json metric: coordinates quantized / second).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU)

A "step" is one pass of the hot path over one batch: the Kodak-shaped bls2017 latent batch of BASELINE.json
configs[1] (24 images x 32x48 x 192 channels = 7,077,888 coordinates, learned factorized prior at reference init,
max_bits_per_coord=10, single lambda=0.5), written out as sorted quantile index (int32) + code length (float32)
per coordinate, plus the per-lambda rate/distortion totals.  With N>1 every rank processes its own batch of that
shape (weak scaling, no data-path collective) and the totals are all-reduced over NCCL inside the timed region
(the totals of four consecutive steps per collective).

`value` is device-resident throughput (CUDA events, max over ranks); `e2e` goes through the reference-facing
ChannelwisePriorCDFQuantizer.compress_batch_channel_latents-level call with pinned HOST buffers (H2D and D2H
inside the timed region); `--impl reference` times the CPU oracle port of the reference's TF-eager quantizer."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMAGES, H, W, C, N_BITS = 24, 32, 48, 192, 10
ROWS = IMAGES * H * W
COORDS = ROWS * C
LAMB = 0.5
BYTES_PER_COORD = 16          # read mu, sigma; write quantile index + code length
L2_BYTES = 126 * 2 ** 20
METRIC = "VBQ coordinates quantized per second"
UNIT = "coords/s"
WORKLOAD = "bls2017 Kodak-shaped latents: 24x32x48x192, learned prior (reference init), N=10, lambda=0.5"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled every ~2 ms WHILE the timed region runs (NVML; the fields are the ones
    of the recipe's `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*` line)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                     "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}

            def loop():
                while not self._stop.is_set():
                    try:
                        self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                        for k, bit in names.items():
                            if r & bit:
                                self.reasons.add(k)
                    except Exception:
                        pass
                    time.sleep(0.002)

            self._thread = threading.Thread(target=loop, daemon=True)
            self._thread.start()
        except Exception:
            self._thread = None
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1)

    def summary(self):
        if not self.samples:
            return None
        sm = sorted(self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------
# synthetic workload
# --------------------------------------------------------------------------------------------------------
def prior_parameters(seed=2, init_scale=10.0):
    """Reference initialisation of BMSHJ2018Prior (learned_prior.py:30-58) in NumPy: constant softplus^-1
    matrices, biases U(-.5,.5), zero factors.  Used by both arms so they quantize against the same prior."""
    rng = np.random.default_rng(seed)
    fdims = (1, 3, 3, 3, 1)
    scale = init_scale ** (1 / 4)
    mats, bs, fs = [], [], []
    for i in range(4):
        init = np.log(np.expm1(1 / scale / fdims[i + 1]))
        mats.append(np.logaddexp(0.0, np.full((C, fdims[i + 1], fdims[i]), init, dtype=np.float32)).astype(np.float32))
        bs.append(rng.uniform(-.5, .5, size=(C, fdims[i + 1], 1)).astype(np.float32))
        if i < 3:
            fs.append(np.zeros((C, fdims[i + 1], 1), dtype=np.float32))
    return mats, bs, fs


def make_prior_and_quantizer(device):
    import vbq_b200
    prior = vbq_b200.BMSHJ2018Prior(C, dims=(3, 3, 3), init_scale=10., device=device)
    prior.set_transformed_parameters(*prior_parameters())
    q = vbq_b200.ChannelwisePriorCDFQuantizer(C, N_BITS, device=device)
    q.build_code_points(prior)
    return prior, q


def make_batch_cpu(table, seed, n_images):
    """CPU-only generator for the reference arm: mu by linear interpolation of the tabulated quantile function at
    u ~ U(0.001, 0.999) (same distribution as `make_batch`), logvar ~ N(-3, 1.5^2)."""
    rng = np.random.default_rng(seed)
    rows = n_images * H * W
    srt = np.sort(table, axis=1)
    xi_sorted = (np.arange(srt.shape[1]) + 1.0) / (srt.shape[1] + 1.0)
    u = rng.uniform(0.001, 0.999, (rows, C))
    mu = np.stack([np.interp(u[:, c], xi_sorted, srt[c]) for c in range(C)], axis=1).astype(np.float32)
    logvar = rng.normal(-3.0, 1.5, (rows, C)).astype(np.float32)
    return mu, (np.exp(logvar) ** np.float32(0.5)).astype(np.float32)


def make_batch(prior, seed, device):
    """mu = F_c^-1(U(0.001,0.999)), logvar ~ N(-3, 1.5^2); returns (mu, sigma) (ROWS, C) f32."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    u = torch.rand((ROWS, C), generator=g, device=device, dtype=torch.float64) * 0.998 + 0.001
    mu = prior.inverse_cdf(u).contiguous()
    logvar = torch.randn((ROWS, C), generator=g, device=device, dtype=torch.float32) * 1.5 - 3.0
    sigma = torch.exp(logvar) ** 0.5
    return mu, sigma.contiguous()


# --------------------------------------------------------------------------------------------------------
# CPU arm: oracle port of the reference's TF-eager quantizer (quantizer.py:156-188 + utils.py:363-423)
# --------------------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(table):
    from oracle import vbq_oracle as O
    oq = O.QuantizerNP(C, N_BITS)
    oq.set_code_points(table, build_grids=True)
    _W["q"] = oq


def _cpu_one_image(args):
    mu, sigma = args
    Z, B = _W["q"].compress_batch_channel_latents(mu, sigma, [LAMB], fast_intervals=False)
    return float(B[LAMB].sum())


def cpu_throughput(table, mu, sigma, n_images, workers, repeats=1):
    """coords/s of the oracle port over ``n_images`` Kodak-shaped images, one image per call like the reference's
    evaluation loop (utils.py:535-542), ``workers`` processes."""
    import multiprocessing as mp
    per = H * W
    jobs = [(mu[i * per:(i + 1) * per], sigma[i * per:(i + 1) * per]) for i in range(n_images)]
    if workers <= 1:
        _cpu_init(table)
        t0 = time.perf_counter()
        for _ in range(repeats):
            for j in jobs:
                _cpu_one_image(j)
        dt = (time.perf_counter() - t0) / repeats
    else:
        with mp.get_context("fork").Pool(workers, initializer=_cpu_init, initargs=(table,)) as pool:
            pool.map(_cpu_one_image, jobs[:workers])           # warm the workers
            t0 = time.perf_counter()
            for _ in range(repeats):
                pool.map(_cpu_one_image, jobs, chunksize=1)
            dt = (time.perf_counter() - t0) / repeats
    return n_images * per * C / dt, dt


def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU path (oracle port; TensorFlow 1.15 is not installable here) on the
    host cores, rank 0 only."""
    if rank != 0:
        return
    from oracle import vbq_oracle as O
    pr = O.LearnedPriorNP(*prior_parameters())
    xi = O.xi_heap(N_BITS)
    table = pr.inverse_cdf_f64(np.repeat(xi[:, None], C, axis=1), iters=64).T
    mu, sigma = make_batch_cpu(table, 1000, IMAGES)
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, IMAGES))
    import multiprocessing as mp
    per = H * W
    budget_s = 100.0                       # the whole --steps K run stays within a few minutes whatever K is
    with mp.get_context("fork").Pool(workers, initializer=_cpu_init, initargs=(table,)) as pool:
        def run(rows_per_job, n_jobs):
            jobs = [(mu[(i % IMAGES) * per:(i % IMAGES) * per + rows_per_job],
                     sigma[(i % IMAGES) * per:(i % IMAGES) * per + rows_per_job]) for i in range(n_jobs)]
            t0 = time.perf_counter()
            pool.map(_cpu_one_image, jobs, chunksize=1)
            return time.perf_counter() - t0
        run(per, workers)                  # warm the workers (search grids, caches)
        t_img = run(per, workers)          # one image per worker
        for _ in range(max(0, args.warmup - 2)):
            run(per, workers)
        # a step is a bounded sample of the workload: whole images (one per call, like utils.py:535-542) while they fit
        # the per-step share of the budget, otherwise the first rows of one image per worker
        share = budget_s / max(args.steps, 1)
        if share >= t_img:
            n_jobs = int(min(IMAGES, workers * max(1, int(share / t_img))))
            rows_per_job = per
        else:
            n_jobs = workers
            rows_per_job = int(max(64, per * share / t_img))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run(rows_per_job, n_jobs)
        total = time.perf_counter() - t0
    coords_per_step = n_jobs * rows_per_job * C
    value = coords_per_step * args.steps / total
    sample = "%d jobs of %d rows x %d channels per step (%s), one call per job, %d worker processes" % (
        n_jobs, rows_per_job, C, "whole Kodak-shaped images" if rows_per_job == per else "leading rows of an image",
        workers)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(index):
    """Pin this rank's host threads to the CPUs that are local to its GPU (NVML affinity) before any pinned host
    buffer is allocated, so that the e2e leg's PCIe traffic does not cross sockets.  Best effort."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
        nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(phys))
    except Exception:
        pass



def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import vbq_b200
    from vbq_b200 import ops, sharding

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    bind_to_gpu_numa_node(local_rank)
    prior, q = make_prior_and_quantizer(dev)
    pen, length = q._length_tables([LAMB])
    args.flags = ops.search_flags([LAMB], args.flags)   # what the facade passes for this lambda
    if world > 1 and not args.no_reserve:
        args.flags |= ops.FLAG_RESERVE_SM               # leave one SM to the overlapped NCCL all-reduce

    # rotating buffer sets so that no step finds its inputs in the 126 MB L2
    set_bytes = COORDS * BYTES_PER_COORD
    n_sets = max(3, -(-3 * L2_BYTES // set_bytes))
    sets = []
    for s in range(n_sets):
        mu, sigma = make_batch(prior, 1000 + 17 * rank + s, dev)
        sets.append(dict(mu=mu, sigma=sigma,
                         qidx=torch.empty((1, ROWS, C), dtype=torch.int32, device=dev),
                         bits=torch.empty((1, ROWS, C), dtype=torch.float32, device=dev)))
    # one validated plan per (totals bucket, buffer set): a step is one prebound vbq_quantize call (one kernel launch
    # incl. totals).  N > 1: the totals of n_sets consecutive steps form one bucket that is all-reduced with ONE NCCL call
    # (asynchronous, on NCCL's own stream, overlapping the kernels of the next bucket); two buckets alternate, and a
    # bucket is reused only after its previous all-reduce has finished (all inside the timed region).  Bucketing keeps
    # most kernel boundaries free of stream operations, so that consecutive launches overlap (programmatic dependent
    # launch) as they do on one GPU.
    n_buckets = 2 if world > 1 else 1
    buckets = [torch.zeros((n_sets, 1, 4), dtype=torch.float64, device=dev) for _ in range(n_buckets)]
    plans = [[ops.QuantizePlan(b["mu"], b["sigma"], q.all_code_points, q._packed, pen, length, None, N_BITS,
                               qidx=b["qidx"], bits=b["bits"], totals=buckets[k][s], flags=args.flags,
                               graph=args.graph) for s, b in enumerate(sets)] for k in range(n_buckets)]
    pending = [None] * n_buckets
    last = {"i": -1}

    def step(i):
        last["i"] = i
        k, s_ = (i // n_sets) % n_buckets, i % n_sets
        if s_ == 0 and pending[k] is not None:
            pending[k].wait()
            pending[k] = None
        t = plans[k][s_].run()
        if world > 1 and s_ == n_sets - 1:
            pending[k] = dist.all_reduce(buckets[k], op=dist.ReduceOp.SUM, async_op=True)
        return t

    def drain():
        if world > 1:
            for k in range(n_buckets):
                if pending[k] is not None:
                    pending[k].wait()
                    pending[k] = None
            # a partially filled last bucket (the run did not end on a bucket boundary) is reduced here
            i = last["i"]
            if i >= 0 and i % n_sets != n_sets - 1:
                dist.all_reduce(buckets[(i // n_sets) % n_buckets], op=dist.ReduceOp.SUM)
                last["i"] = -1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    drain()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev[0].record()
        for i in range(args.steps):
            step(args.warmup + i)
        drain()
        ev[1].record()
        barrier()
    total_ms = ev[0].elapsed_time(ev[1])
    # average launch duration of the kernel over the timed region (launch gaps included): the steps are back to back
    # on one stream and each step is exactly one launch of the quantize kernel
    per_launch_ms = [total_ms / args.steps]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = COORDS * world * args.steps / (total_ms * 1e-3)

    # end-to-end through the C-ABI host entry point (vbq_quantize_host, what the reference-facing
    # compress_batch_channel_latents runs for host arrays): HOST pinned buffers in, HOST pinned buffers out, the
    # upload / kernel / download of row chunks overlapped on three streams, all inside the timed region.
    h_mu = [b["mu"].cpu().pin_memory() for b in (sets * 2)[:2]]
    h_sigma = [b["sigma"].cpu().pin_memory() for b in (sets * 2)[:2]]
    h_q = torch.empty((1, ROWS, C), dtype=torch.int32).pin_memory()
    h_b = torch.empty((1, ROWS, C), dtype=torch.float32).pin_memory()
    h_tot = torch.empty((1, 4), dtype=torch.float64).pin_memory()
    pipe = ops.HostPipeline(C, N_BITS, 1, args.chunk_rows, ops.OUT_QIDX | ops.OUT_BITS | ops.OUT_TOTALS, device=dev)

    def e2e_step(i):
        pipe.run(h_mu[i % 2], h_sigma[i % 2], q.all_code_points, q._packed, pen, length, None,
                 qidx=h_q, bits=h_b, totals=h_tot, flags=args.flags)
        if world > 1:
            t_ = h_tot.to(dev)
            sharding.all_reduce_totals(t_)
            h_tot.copy_(t_)
        return float(h_tot[0, 1])

    e2e_steps = max(3, min(args.steps, 20))
    for i in range(2):
        e2e_step(i)
    # the pipeline must reproduce the device-resident results exactly
    step(0)
    drain()
    torch.cuda.synchronize()
    pipe.run(h_mu[0], h_sigma[0], q.all_code_points, q._packed, pen, length, None, qidx=h_q, bits=h_b, totals=h_tot,
             flags=args.flags)
    assert torch.equal(h_q[0], sets[0]["qidx"][0].cpu()) and torch.equal(h_b[0], sets[0]["bits"][0].cpu())
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
    e2e_s = time.perf_counter() - t0
    pipe.close()
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = COORDS * world * e2e_steps / float(t.item())

    if rank != 0:
        return
    peaks, peak_kind = measured_peaks()
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r1_traffic.json")   # dram__bytes_read+write of one ncu --set full capture
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f)["traffic_bytes_per_launch"]
    kern_ms = float(np.median(per_launch_ms))
    achieved = COORDS * BYTES_PER_COORD / (kern_ms * 1e-3) / 1e9
    cpu = None
    if world == 1 and not args.no_cpu:
        b = sets[0]
        n_img = IMAGES
        v, dt = cpu_throughput(q.all_code_points.cpu().numpy(), b["mu"][:n_img * H * W].cpu().numpy(),
                               b["sigma"][:n_img * H * W].cpu().numpy(), n_img, 1)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "%d of the 24 Kodak-shaped images (%d coordinates), one image per call, single process, %.1f s"
                         % (n_img, n_img * H * W * C, dt)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "coords_per_step_per_gpu": COORDS, "max_bits_per_coord": N_BITS,
                   "lambdas": [LAMB], "outputs": "sorted quantile index int32 + code length f32 + totals",
                   "l2": "%d rotating input/output sets (%d MB) > 126 MB L2" % (n_sets, n_sets * set_bytes >> 20),
                   "flags": args.flags, "parallelism": "dp%d, totals (n_lambda,4) f64 all-reduced over NCCL in buckets of %d steps" % (world, n_sets)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_kind": peak_kind,
                     "kernel": "vbq_bisect_kernel", "kernel_ms": kern_ms,
                     "algorithmic_bytes_per_launch": COORDS * BYTES_PER_COORD},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * COORDS * 4,
                "d2h_bytes_per_step": 2 * COORDS * 4 + 32, "steps": e2e_steps,
                "api": "vbq_quantize_host (pinned host in/out, %d-row chunks, 3 streams)" % args.chunk_rows},
        "gpu_launches": args.steps,
        "clocks": clocks.summary(),
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--flags", type=int, default=0, help="VBQ_FLAG_* bits passed to vbq_quantize")
    ap.add_argument("--chunk-rows", type=int, default=9216, help="rows per chunk of the host pipeline (e2e leg)")
    ap.add_argument("--no-reserve", action="store_true", help="multi-GPU: do not leave an SM to the NCCL kernel")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--graph", action="store_true", help="replay one CUDA graph per step instead of the prebound eager "
                    "call (the eager launches overlap through programmatic dependent launch and measure faster)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
