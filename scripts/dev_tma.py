"""Development check of the TMA-pipeline kernel against the cp.async bisection kernel and the literal reference walk on the
benchmark workload (and a few ragged shapes): bit-exact outputs, reproducible totals, and CUDA-event timing of each.

  python scripts/dev_tma.py [--steps 200] [--quick]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import vbq_b200  # noqa: E402
from vbq_b200 import ops  # noqa: E402


def run_once(q, mu, sigma, pen, length, flags, outs=("qidx", "bits"), N=bench.N_BITS):
    dev = mu.device
    L = pen.shape[0]
    rows, C = mu.shape
    o = {}
    dt = {"zhat": torch.float32, "qidx": torch.int32, "level": torch.int32, "bits": torch.float32}
    for k in outs:
        o[k] = torch.full((L, rows, C), -7, dtype=dt[k], device=dev)
    tot = torch.zeros((L, 4), dtype=torch.float64, device=dev)
    ws = ops.quantize_workspace(L, dev)
    ops.quantize_into(mu, sigma, q.all_code_points, q._packed, pen, length, None, N, totals=tot, workspace=ws,
                      flags=flags, **o)
    torch.cuda.synchronize()
    return o, tot


def time_plan(q, sets, pen, length, flags, steps):
    plans = []
    for b in sets:
        plans.append(ops.QuantizePlan(b["mu"], b["sigma"], q.all_code_points, q._packed, pen, length, None, bench.N_BITS,
                                      qidx=b["qidx"], bits=b["bits"], totals=b["tot"], flags=flags))
    for i in range(10):
        plans[i % len(plans)].run()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    t0 = time.perf_counter()
    ev[0].record()
    for i in range(steps):
        plans[i % len(plans)].run()
    ev[1].record()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / steps
    return ms, 1e6 * t_issue / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--variants", default="", help="comma-separated VBQ_TMA_VARIANT values (development builds)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    prior, q = bench.make_prior_and_quantizer(dev)
    pen, length = q._length_tables([bench.LAMB])
    base = ops.search_flags([bench.LAMB])

    # ---- correctness on the benchmark batch -----------------------------------------------------------------
    mu, sigma = bench.make_batch(prior, 1000, dev)
    ref, tot_ref = run_once(q, mu, sigma, pen, length, base | ops.FLAG_NO_TMA)
    new, tot_new = run_once(q, mu, sigma, pen, length, base)
    ok = all(torch.equal(ref[k], new[k]) for k in ref)
    print("bench batch: outputs equal:", ok, " totals ref", tot_ref.tolist(), "new", tot_new.tolist())
    rel = ((tot_new - tot_ref).abs() / tot_ref.abs().clamp_min(1e-300)).max().item()
    print("  totals max rel diff vs cp.async kernel: %.3e" % rel)
    again = [run_once(q, mu, sigma, pen, length, base)[1] for _ in range(5)]
    print("  totals reproducible over 5 runs:", all(torch.equal(t, tot_new) for t in again))
    assert ok

    # ---- ragged shapes, other depths, prune, several lambdas through NO_SWEEP ---------------------------------
    if not args.quick:
        import numpy as np
        for (rows, C, N, lambs) in [(1, 16, 10, [0.5]), (17, 20, 10, [0.5]), (1000, 36, 10, [0.01, 2.0]),
                                    (4099, 192, 10, [0.5, 8.0, 0.0]), (96 * 7 + 5, 48, 6, [0.3]), (128 * 3 + 48, 16, 10, [0.5]), (2000, 32, 10, [3.0]), (333, 12, 0, [1.0]),
                                    (50000, 64, 10, [4.0]), (36864, 192, 10, [2.0 ** -8])]:
            pr = vbq_b200.BMSHJ2018Prior(C, device=dev)
            qq = vbq_b200.ChannelwisePriorCDFQuantizer(C, N, device=dev)
            qq.build_code_points(pr)
            g = torch.Generator(device=dev)
            g.manual_seed(rows * 7 + C)
            u = torch.rand((rows, C), generator=g, device=dev, dtype=torch.float64) * 0.998 + 0.001
            m = pr.inverse_cdf(u).contiguous()
            # a third of the coordinates exactly on code points / midpoints: near-ties in bulk
            tab = qq.all_code_points
            idx = torch.randint(0, tab.shape[1], (rows, C), generator=g, device=dev)
            onpt = tab.t()[idx, torch.arange(C, device=dev)[None, :].expand(rows, C)]
            sel = torch.rand((rows, C), generator=g, device=dev) < 0.33
            m = torch.where(sel, onpt, m).contiguous()
            s = torch.exp(0.5 * (torch.randn((rows, C), generator=g, device=dev) * 1.5 - 3.0)).contiguous()
            p2, l2 = qq._length_tables(lambs)
            for outs in (("qidx", "bits"), ("zhat", "level"), ("zhat",), ("qidx",), ()):
                for fl in (0, ops.FLAG_NO_PRUNE):
                    f = fl | ops.FLAG_NO_SWEEP
                    r_, tr = run_once(qq, m, s, p2, l2, f | ops.FLAG_NO_TMA, outs, N)
                    for pen_ in (p2, p2.clone()):     # with / without the host copy of the penalties
                        n_, tn = run_once(qq, m, s, pen_, l2, f, outs, N)
                        same = all(torch.equal(r_[k], n_[k]) for k in r_)
                        relt = ((tn - tr).abs() / tr.abs().clamp_min(1e-300)).max().item()
                        n2_, tn2 = run_once(qq, m, s, pen_, l2, f, outs, N)
                        if not same or relt > 2e-7 or (pen_ is p2 and not torch.equal(tn, tn2)):
                            print("MISMATCH rows=%d C=%d N=%d lambs=%s outs=%s flags=%d same=%s totals rel %.2e repro %s" %
                                  (rows, C, N, lambs, outs, f, same, relt, torch.equal(tn, tn2)))
                            raise SystemExit(1)
            print("ok rows=%d C=%d N=%d lambs=%s" % (rows, C, N, lambs))

    # ---- timing ------------------------------------------------------------------------------------------------
    sets = []
    for s_ in range(4):
        m, s = bench.make_batch(prior, 2000 + s_, dev)
        sets.append(dict(mu=m, sigma=s, qidx=torch.empty((1, bench.ROWS, bench.C), dtype=torch.int32, device=dev),
                         bits=torch.empty((1, bench.ROWS, bench.C), dtype=torch.float32, device=dev),
                         tot=torch.zeros((1, 4), dtype=torch.float64, device=dev)))
    res = {}
    runs = [("tma", base, "0"), ("cp_async", base | ops.FLAG_NO_TMA, "0")]
    for v in args.variants.split(","):
        if v:
            runs.append(("tma_v" + v, base, v))
    for name, fl, var in runs:
        os.environ["VBQ_TMA_VARIANT"] = var
        ms, issue_us = time_plan(q, sets, pen, length, fl, args.steps)
        res[name] = {"us_per_step": 1e3 * ms, "G_coords_s": bench.COORDS / ms / 1e6, "host_issue_us": issue_us,
                     "roofline_frac": bench.COORDS * 16 / (ms * 1e-3) / 1e9 / 6533.8}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
