"""Development check of the both-ends TMA kernel (arbitrary penalties) against the literal reference walk and the bracket-walk
kernel: bit-exact outputs on near-tie-heavy inputs, reproducible totals, timing.  python scripts/dev_both.py [--quick]"""
import argparse, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, vbq_b200
from vbq_b200 import ops


def run(q, mu, sigma, pen, length, em, flags, outs, N=10, want_em=True):
    dev = mu.device
    L = pen.shape[0]
    rows, C = mu.shape
    dt = {"zhat": torch.float32, "qidx": torch.int32, "level": torch.int32, "bits": torch.float32}
    o = {k: torch.full((L, rows, C), -7, dtype=dt[k], device=dev) for k in outs}
    if em is not None and want_em and "zhat" in outs:
        o["em_bits"] = torch.full((L, rows, C), -7.0, dtype=torch.float32, device=dev)
    tot = torch.zeros((L, 4), dtype=torch.float64, device=dev)
    ws = ops.quantize_workspace(L, dev)
    ops.quantize_into(mu, sigma, q.all_code_points, q._packed, pen, length, em, N, totals=tot, workspace=ws, flags=flags, **o)
    torch.cuda.synchronize()
    return o, tot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--variants", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    if not args.quick:
        for (rows, C, lambs, seed) in [(1, 16, [0.5], 1), (37, 20, [0.3], 2), (1000, 36, [0.01, 2.0], 3), (4099, 192, [0.5, 8.0, 0.0], 4),
                                       (20000, 64, [0.05], 5), (36864, 192, [2.0 ** -8], 6)]:
            N = 10
            pr = vbq_b200.BMSHJ2018Prior(C, device=dev, seed=seed)
            q = vbq_b200.ChannelwisePriorCDFQuantizer(C, N, device=dev)
            q.build_code_points(pr)
            g = torch.Generator(device=dev); g.manual_seed(seed)
            u = torch.rand((rows, C), generator=g, device=dev, dtype=torch.float64) * 0.9998 + 0.0001
            m = pr.inverse_cdf(u).contiguous()
            tab = q.all_code_points
            idx = torch.randint(0, tab.shape[1], (rows, C), generator=g, device=dev)
            onpt = tab.t()[idx, torch.arange(C, device=dev)[None, :].expand(rows, C)]
            sel = torch.rand((rows, C), generator=g, device=dev) < 0.33
            m = torch.where(sel, onpt, m).contiguous()
            if rows > 8:
                m[0] = 1e4; m[1] = -1e4; m[2] = tab[:, -1] * 1.0000001 + 1e-3
            s = torch.exp(0.5 * (torch.randn((rows, C), generator=g, device=dev) * 1.5 - 3.0)).contiguous()
            rng = np.random.default_rng(seed)
            L = len(lambs)
            R = rng.gamma(2.0, 2.0, size=(L, C, N + 1)).astype(np.float32)          # non-monotone corrections
            length = (np.arange(N + 1, dtype=np.float32)[None, None, :] + R).astype(np.float32)
            pen = (np.asarray(lambs, dtype=np.float32)[:, None, None] * length).astype(np.float32)
            pen_t = ops.with_host_copy(pen, dev)
            len_t = torch.from_numpy(length).to(dev)
            em = torch.from_numpy(rng.gamma(2.0, 3.0, size=(L, C, 2 ** (N + 1) - 1)).astype(np.float32)).to(dev)
            for outs in (("zhat", "bits"), ("qidx",), ()):
                for em_ in (em, None):
                    if em_ is not None and outs == ("qidx",):
                        continue
                    f = ops.FLAG_NO_SWEEP
                    ref, tr = run(q, m, s, pen_t, len_t, em_, f | ops.FLAG_REFERENCE_WALK, outs)
                    new, tn = run(q, m, s, pen_t, len_t, em_, f, outs)
                    new2, tn2 = run(q, m, s, pen_t, len_t, em_, f, outs)
                    same = all(torch.equal(ref[k], new[k]) for k in ref)
                    rel = ((tn - tr).abs() / tr.abs().clamp_min(1e-300)).max().item()
                    if not same or rel > 2e-7 or not torch.equal(tn, tn2):
                        for k in ref:
                            bad = (ref[k] != new[k]).nonzero()
                            print(k, "mismatches", bad.shape[0], bad[:5].tolist())
                        print("MISMATCH rows=%d C=%d lambs=%s outs=%s em=%s same=%s rel=%.2e repro=%s" % (rows, C, lambs, outs, em_ is not None, same, rel, torch.equal(tn, tn2)))
                        print(tr.tolist(), tn.tolist())
                        raise SystemExit(1)
            print("ok rows=%d C=%d lambs=%s" % (rows, C, lambs))
    # timing on the Kodak batch with fitted entropy models
    prior, q0 = bench.make_prior_and_quantizer(dev)
    q = vbq_b200.ChannelwisePriorCDFQuantizer(bench.C, bench.N_BITS, device=dev)
    q.set_code_points(q0.all_code_points)
    sets = [bench.make_batch(prior, 50 + i, dev) for i in range(4)]
    grid = [float(l) for l in 2.0 ** np.linspace(-8, 7, 16)]
    q.build_entropy_models_from_latents(sets[0][0], (2.0 * torch.log(sets[0][1])).contiguous(), grid, add_n_smoothing=1.0)
    res = {}
    runs = [("both", 0, "0"), ("sweep", 0, "0"), ("both_every_depth", ops.FLAG_NEIGHBOUR_EVERY_DEPTH, "0"), ("bracket_walk", ops.FLAG_BRACKET_WALK, "0")] + [("both_v" + v, 0, v) for v in args.variants.split(",") if v]
    rows, C = sets[0][0].shape
    for name, fl, var in runs:
        os.environ["VBQ_TMA_VARIANT"] = var
        for lambs in ([0.5], grid):
            L = len(lambs)
            pen, length = q._length_tables(lambs)
            em = q._entropy_model_tensor(lambs)
            z = torch.empty((L, rows, C), dtype=torch.float32, device=dev)
            b = torch.empty((L, rows, C), dtype=torch.float32, device=dev)
            e = torch.empty((L, rows, C), dtype=torch.float32, device=dev)
            if os.environ.get("VBQ_NO_EM"):
                em, e = None, None
            plans = [ops.QuantizePlan(mu, sg, q.all_code_points, q._packed, pen, length, em, bench.N_BITS, zhat=z, bits=b, em_bits=e,
                                      totals=torch.zeros((L, 4), dtype=torch.float64, device=dev),
                                      flags=fl | (ops.FLAG_NO_SWEEP if name.startswith("both") else 0)) for mu, sg in sets]
            for i in range(3):
                plans[i % 4].run()
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            for i in range(20):
                plans[i % 4].run()
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / 20
            res["%s_L%d" % (name, L)] = {"ms_per_call": ms, "G_coord_lambda_s": bench.COORDS * L / ms / 1e6}
    # totals only (the rate-distortion sweep after build_entropy_models): bracket-walk sweep vs one both-ends launch per lambda
    for name, fl in (("totals_sweep", 0), ("totals_per_lambda", ops.FLAG_NO_SWEEP)):
        for with_em in (True, False):
            L = len(grid)
            pen, length = q._length_tables(grid)
            em = q._entropy_model_tensor(grid) if with_em else None
            plans = [ops.QuantizePlan(mu, sg, q.all_code_points, q._packed, pen, length, em, bench.N_BITS,
                                      totals=torch.zeros((L, 4), dtype=torch.float64, device=dev), flags=fl) for mu, sg in sets]
            for i in range(3):
                plans[i % 4].run()
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            for i in range(10):
                plans[i % 4].run()
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / 10
            res["%s_L16%s" % (name, "_em" if with_em else "")] = {"ms_per_call": ms, "G_coord_lambda_s": bench.COORDS * L / ms / 1e6,
                                                                  "totals": plans[0].run()[3].tolist()}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
