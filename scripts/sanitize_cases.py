"""Small invocations of every search kernel for compute-sanitizer (memcheck / racecheck / synccheck): the TMA pipelines
(raw lengths, arbitrary penalties + entropy-model gather, early exit, logvar inputs, ragged rows and channels), the cp.async
bisection, the sweeps, the deep-table path, the host pipeline and the peer-totals kernels (world = 1).

  compute-sanitizer --tool memcheck python scripts/sanitize_cases.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vbq_b200                      # noqa: E402
from vbq_b200 import ops, sharding   # noqa: E402

dev = torch.device("cuda", 0)


def case(rows, C, N, lambs, seed, corrected=False):
    pr = vbq_b200.BMSHJ2018Prior(C, device=dev, seed=seed)
    q = vbq_b200.ChannelwisePriorCDFQuantizer(C, N, device=dev)
    q.build_code_points(pr)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    u = torch.rand((rows, C), generator=g, device=dev, dtype=torch.float64) * 0.998 + 0.001
    mu = pr.inverse_cdf(u).contiguous()
    lv = (torch.randn((rows, C), generator=g, device=dev) * 1.5 - 3.0).contiguous()
    sg = torch.exp(0.5 * lv).contiguous()
    if corrected:
        q.build_entropy_models_from_latents(mu, lv, lambs, add_n_smoothing=1.0)
    outs = ops.OUT_ZHAT | ops.OUT_BITS | ops.OUT_TOTALS
    for fl in (0, ops.FLAG_NO_TMA, ops.FLAG_NO_SWEEP, ops.FLAG_BRACKET_WALK | ops.FLAG_NO_SWEEP):
        q.quantize(mu, sg, lambs, outputs=outs, flags=fl, entropy_bits=corrected)
        q.quantize(mu, lv, lambs, logvar=True, outputs=ops.OUT_QIDX | ops.OUT_TOTALS, flags=fl)
    q.quantize(mu, sg, lambs[:1], outputs=ops.OUT_QIDX | ops.OUT_BITS | ops.OUT_TOTALS)
    torch.cuda.synchronize()
    return q, mu, sg, lv


q, mu, sg, lv = case(700, 48, 10, [0.5], 1)                      # raw lengths: TMA kernel, cp.async kernel, bracket walk
case(333, 20, 10, [4.0, 0.02], 2)                                # ragged channels, early exit, two lambdas (sweeps)
case(515, 32, 10, [0.5, 2.0], 3, corrected=True)                 # arbitrary penalties + entropy-model gather
case(260, 16, 6, [0.3], 4)                                       # run-time depth
case(300, 16, 13, [0.5, 0.004], 5)                               # deep tables (global-memory depths)
Z, B = q.compress_batch_channel_latents(mu.cpu().numpy(), sg.cpu().numpy(), [0.5])      # host pipeline
peer = sharding.PeerTotals(n_lambda_max=1)                       # peer totals, world = 1
pen, length = q._length_tables([0.5])
tot = torch.zeros((1, 4), dtype=torch.float64, device=dev)
glob = torch.zeros((1, 4), dtype=torch.float64, device=dev)
plan = ops.QuantizePlan(mu, sg, q.all_code_points, q._packed, pen, length, None, 10,
                        qidx=torch.empty((1,) + tuple(mu.shape), dtype=torch.int32, device=dev), totals=tot,
                        flags=ops.search_flags([0.5]), peer=peer)
s1, s2, s3 = peer.next_seq(), peer.next_seq(), peer.next_seq()
plan.run_peer()
plan.run_peer(s1, tot)
plan.run_peer(s2, tot, s1, glob)
peer.push(s3, tot)
peer.collect(s2, 1, glob)
peer.collect(s3, 1, glob)
torch.cuda.synchronize()
assert torch.equal(glob, tot)
peer.close()
print("sanitize cases done")
