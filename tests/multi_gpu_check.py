#!/usr/bin/env python
"""Multi-GPU correctness check (run under torchrun on a box with >= 2 GPUs; NCCL):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py

Every rank quantizes its contiguous shard of the images; the all-reduced per-lambda totals must equal the totals of
the unsharded batch, the gathered outputs must equal the unsharded outputs bit for bit, and the entropy models fitted
from all-reduced histograms must equal the ones fitted on one GPU.  (The same logic is covered on CPU with gloo and the
oracle by tests/test_sharding.py; this script checks the CUDA + NCCL path.)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vbq_b200                                  # noqa: E402
from vbq_b200 import ops, sharding               # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    C, N, images, rows_per_image = 48, 10, 13, 96      # 13 images: ragged split
    lambs = [float(l) for l in 2 ** np.linspace(-6, 5, 6)]
    prior = vbq_b200.BMSHJ2018Prior(C, device=dev, seed=3)      # same seed on every rank: replicated prior
    q = vbq_b200.ChannelwisePriorCDFQuantizer(C, N, device=dev)
    q.build_code_points(prior)
    g = torch.Generator(device=dev)
    g.manual_seed(11)
    u = torch.rand((images, rows_per_image, C), generator=g, device=dev, dtype=torch.float64) * 0.998 + 0.001
    mu = prior.inverse_cdf(u.reshape(-1, C)).reshape(images, rows_per_image, C)
    logvar = torch.randn((images, rows_per_image, C), generator=g, device=dev) * 1.5 - 3.0

    outs = ops.OUT_ZHAT | ops.OUT_QIDX | ops.OUT_TOTALS
    a, b = sharding.shard_bounds(images, world, rank)
    sq = sharding.ShardedQuantizer(q)
    loc = q.quantize(mu[a:b].reshape(-1, C), logvar[a:b].reshape(-1, C), lambs, logvar=True, outputs=outs,
                     flags=ops.FLAG_RESERVE_SM)
    tot = sharding.all_reduce_totals(loc["totals"].clone())
    tot2 = sq.rd_sweep(mu[a:b].reshape(-1, C), logvar[a:b].reshape(-1, C), lambs, logvar=True)
    full = q.quantize(mu.reshape(-1, C), logvar.reshape(-1, C), lambs, logvar=True, outputs=outs)
    assert torch.allclose(tot, full["totals"], rtol=1e-12), (tot, full["totals"])
    assert torch.allclose(tot2, full["totals"], rtol=1e-12)
    assert torch.equal(loc["qidx"], full["qidx"][:, a * rows_per_image:b * rows_per_image])
    assert torch.equal(loc["zhat"], full["zhat"][:, a * rows_per_image:b * rows_per_image])

    # entropy models: sharded fit (histograms all-reduced over NCCL) == single-GPU fit
    sq.build_entropy_models_from_latents(mu[a:b], logvar[a:b], lambs, add_n_smoothing=1)
    rcl_s, em_s = q.raw_code_length_entropy_models, q.entropy_models
    q.build_entropy_models_from_latents(mu, logvar, lambs, add_n_smoothing=1)
    for l in lambs:
        assert np.array_equal(rcl_s[l], q.raw_code_length_entropy_models[l])
        assert np.array_equal(em_s[l], q.entropy_models[l])
    # peer totals: the search kernel's last CTA delivers the sums to every rank's inbox (no NCCL launch); the collected
    # sum must equal the NCCL all-reduce and, for every rank, be the same bits (rank-ordered addition)
    peer = sharding.PeerTotals(n_lambda_max=len(lambs))
    q.raw_code_length_entropy_models, q.entropy_models = None, None
    q._cache = {}
    m_loc, s_loc = mu[a:b].reshape(-1, C).contiguous(), torch.exp(0.5 * logvar[a:b]).reshape(-1, C).contiguous()
    for lam_list in ([0.5], lambs):
        L = len(lam_list)
        pen, length = q._length_tables(lam_list)
        loc_tot = torch.zeros((L, 4), dtype=torch.float64, device=dev)
        qidx = torch.empty((L, m_loc.shape[0], C), dtype=torch.int32, device=dev)
        bits = torch.empty((L, m_loc.shape[0], C), dtype=torch.float32, device=dev)
        plan = ops.QuantizePlan(m_loc, s_loc, q.all_code_points, q._packed, pen, length, None, N, qidx=qidx, bits=bits,
                                totals=loc_tot, flags=ops.search_flags(lam_list), peer=peer)
        got = [torch.zeros((L, 4), dtype=torch.float64, device=dev) for _ in range(12)]
        seqs = [peer.next_seq() for _ in range(12)]
        for i in range(12):            # more calls than inbox slots; call i delivers call i-1 and collects call i-3 ...
            if i % 2:
                plan.run_peer(seqs[i - 1], loc_tot, seqs[i - 3] if i >= 3 else 0, got[i - 3] if i >= 3 else None)
            else:                      # ... or the stand-alone kernels do
                if i >= 1:
                    peer.push(seqs[i - 1], loc_tot)
                if i >= 3:
                    peer.collect(seqs[i - 3], L, got[i - 3])
                plan.run()
        peer.push(seqs[11], loc_tot)
        for i in range(9, 12):
            peer.collect(seqs[i], L, got[i])
        torch.cuda.synchronize()
        want = sharding.all_reduce_totals(loc_tot.clone())
        for t in got:
            assert torch.allclose(t, want, rtol=1e-13), (t, want)
        every = [torch.zeros_like(got[0]) for _ in range(world)]
        dist.all_gather(every, got[0])
        assert all(torch.equal(e, every[0]) for e in every), "peer totals differ between ranks"
    peer.close()

    dist.barrier()
    if rank == 0:
        print("multi-GPU check ok: world=%d, totals (NCCL and peer inboxes), outputs and entropy models agree with the unsharded run" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
