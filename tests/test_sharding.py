"""CPU tests of the data-parallel layer: shard planning and the all-reduce of per-lambda totals / histogram counts
over a world_size-2 gloo group (the NCCL path on the GPU box runs the same code)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vbq_oracle as O
from vbq_b200 import sharding


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 24, 4096, 1000003):
        for w in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)
    x = torch.arange(10).reshape(5, 2)
    assert torch.equal(sharding.shard_leading_axis(x, 2, 1), x[3:])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q_out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        C, N, images, rows_per_image = 4, 6, 5, 30
        pr = O.LearnedPriorNP.init(C, rng=np.random.default_rng(1), factor_std=0.3)
        oq = O.QuantizerNP(C, N)
        oq.build_code_points(pr.inverse_cdf_f64)
        mu = rng.normal(0, 8, (images, rows_per_image, C)).astype(np.float32)
        sg = np.exp(rng.normal(-1.5, 0.7, mu.shape)).astype(np.float32)
        lambs = [0.05, 0.5, 4.0]
        a, b = sharding.shard_bounds(images, world, rank)          # images are the sharded units
        Zh, nb = oq.compress_batch_channel_latents(mu[a:b].reshape(-1, C), sg[a:b].reshape(-1, C), lambs)
        tot = torch.zeros((len(lambs), 4), dtype=torch.float64)
        cnt = torch.zeros((len(lambs), C, N + 1), dtype=torch.int64)
        for i, l in enumerate(lambs):
            bits, distn = O.rd_totals(mu[a:b].reshape(-1, C), sg[a:b].reshape(-1, C), Zh[l], nb[l])
            tot[i, 0] = tot[i, 1] = bits
            tot[i, 3] = distn
            for c in range(C):
                cnt[i, c] = torch.from_numpy(np.bincount(nb[l][:, c], minlength=N + 1))
        sharding.all_reduce_totals(tot)
        sharding.all_reduce_counts(cnt)
        if rank == 0:
            Zf, nf = oq.compress_batch_channel_latents(mu.reshape(-1, C), sg.reshape(-1, C), lambs)
            ok = True
            for i, l in enumerate(lambs):
                bits, distn = O.rd_totals(mu.reshape(-1, C), sg.reshape(-1, C), Zf[l], nf[l])
                ok &= abs(float(tot[i, 1]) - bits) < 1e-9 and abs(float(tot[i, 3]) - distn) <= 1e-9 * distn
                full = np.stack([np.bincount(nf[l][:, c], minlength=N + 1) for c in range(C)])
                ok &= np.array_equal(cnt[i].numpy(), full)
            q_out.put(bool(ok))
    finally:
        dist.destroy_process_group()


def test_totals_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q_out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q_out.get(timeout=5) is True


def test_allreduce_is_identity_without_process_group():
    t = torch.ones((2, 4), dtype=torch.float64)
    assert sharding.all_reduce_totals(t) is t
    with pytest.raises(TypeError):
        sharding.all_reduce_totals(torch.ones(2, 4))
    with pytest.raises(TypeError):
        sharding.all_reduce_counts(torch.ones(2, 4))
