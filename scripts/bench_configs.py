#!/usr/bin/env python
"""Secondary measurements on one GPU (not the headline bench): the lambda sweep (BASELINE configs[2] shape per
GPU), word-embedding rows (configs[3]) and deep tables N=16 (configs[4]), each at a single-GPU slice of the
configuration.  Prints one JSON object per line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vbq_b200                      # noqa: E402
from vbq_b200 import ops             # noqa: E402
import bench                         # noqa: E402


def timeit(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(steps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / steps * 1e-3


def sweep_case(n_lambda, flags, outputs, images=24):
    dev = torch.device("cuda", 0)
    prior, q = bench.make_prior_and_quantizer(dev)
    mu, sigma = bench.make_batch(prior, 7, dev)
    lambs = [float(l) for l in 2 ** np.linspace(-8, 7, n_lambda)]
    pen, length = q._length_tables(lambs)
    rows, C = mu.shape
    totals = torch.zeros((n_lambda, 4), dtype=torch.float64, device=dev)
    ws = ops.quantize_workspace(n_lambda, dev)
    kw = {}
    if outputs:
        kw = dict(qidx=torch.empty((n_lambda, rows, C), dtype=torch.int32, device=dev),
                  bits=torch.empty((n_lambda, rows, C), dtype=torch.float32, device=dev))

    def fn():
        ops.quantize_into(mu, sigma, q.all_code_points, q._packed, pen, length, None, bench.N_BITS,
                          totals=totals, workspace=ws, flags=flags, **kw)

    t = timeit(fn)
    return dict(case="sweep", n_lambda=n_lambda, flags=flags, full_outputs=bool(outputs), coords=rows * C,
                seconds=t, coord_lambda_per_s=rows * C * n_lambda / t)


def deep_case(N=16, C=320, rows=128 * 128 * 2, lamb=0.5, flags=0):
    dev = torch.device("cuda", 0)
    prior = vbq_b200.BMSHJ2018Prior(C, device=dev, seed=5)
    q = vbq_b200.ChannelwisePriorCDFQuantizer(C, N, device=dev)
    t0 = time.perf_counter()
    q.build_code_points(prior)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    u = torch.rand((rows, C), generator=g, device=dev, dtype=torch.float64) * 0.998 + 0.001
    mu = prior.inverse_cdf(u).contiguous()
    sigma = torch.exp(0.5 * (torch.randn((rows, C), generator=g, device=dev) * 1.5 - 3.0)).contiguous()
    pen, length = q._length_tables([lamb])
    qidx = torch.empty((1, rows, C), dtype=torch.int32, device=dev)
    bits = torch.empty((1, rows, C), dtype=torch.float32, device=dev)

    def fn():
        ops.quantize_into(mu, sigma, q.all_code_points, q._packed, pen, length, None, N, qidx=qidx, bits=bits,
                          flags=flags)

    t = timeit(fn, steps=5)
    return dict(case="deep", N=N, C=C, rows=rows, lamb=lamb, flags=flags, table_build_s=t_build, seconds=t,
                coords_per_s=rows * C / t, level_hist=torch.bincount(
                    (31 - torch.log2((qidx[0].flatten()[:100000] + 1).float())).long().clamp(0) * 0).tolist()[:1])


def embedding_case(V=1_000_000, K=300, beta=1.0, flags=0):
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    means = torch.randn((V, K), generator=g, device=dev) * 1.2329 - 0.08
    stds = torch.exp(torch.randn((V, K), generator=g, device=dev) * 0.7 + float(np.log(0.04)))
    cb = vbq_b200.GaussianCodebook(vbq_b200.word_embeddings.empirical_std(means), 10, device=dev)

    def fn():
        cb.quantize(means, stds, [beta], outputs=ops.OUT_ZHAT, flags=flags)

    t = timeit(fn, steps=5)
    return dict(case="embeddings", V=V, K=K, beta=beta, flags=flags, seconds=t, coords_per_s=V * K / t)


if __name__ == "__main__":
    which = sys.argv[1:] or ["sweep", "deep", "emb"]
    if "sweep" in which:
        for L in (16, 64):
            for flags in (0, ops.FLAG_NO_SWEEP, ops.FLAG_FAST):
                print(json.dumps(sweep_case(L, flags, outputs=False)), flush=True)
        print(json.dumps(sweep_case(16, 0, outputs=True)), flush=True)
    if "deep" in which:
        for lamb in (0.5, 2.0 ** -8):
            print(json.dumps(deep_case(lamb=lamb)), flush=True)
    if "emb" in which:
        for flags in (0, ops.FLAG_FAST):
            print(json.dumps(embedding_case(flags=flags)), flush=True)
