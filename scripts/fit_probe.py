"""Development: time of the two passes of the entropy-model fit (quantizer.py:82-150) on the Kodak batch, 16 lambdas."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, vbq_b200
dev = torch.device("cuda", 0)
prior, q0 = bench.make_prior_and_quantizer(dev)
q = vbq_b200.ChannelwisePriorCDFQuantizer(bench.C, bench.N_BITS, device=dev)
q.set_code_points(q0.all_code_points)
b = bench.make_batch(prior, 50, dev)
mu, sg = (b["mu"], b["sigma"]) if isinstance(b, dict) else b
lv = (2.0 * torch.log(sg)).contiguous()
grid = [float(l) for l in 2.0 ** np.linspace(-8, 7, 16)]
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c1 = q._histograms(mu, lv, grid, 'level')
    torch.cuda.synchronize(); t1 = time.perf_counter()
    q.build_entropy_models_from_latents(mu, lv, grid, add_n_smoothing=1.0)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    c2 = q._histograms(mu, lv, grid, 'qidx')
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print("pass 1 (depth counts) %.2f ms, whole fit %.2f ms, pass 2 (symbol counts) %.2f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)))
print(c1.shape, c1.sum(), c2.shape, c2.sum())
if os.environ.get("VBQ_FIT_PROFILE"):
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    q.build_entropy_models_from_latents(mu, lv, grid, add_n_smoothing=1.0)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
from vbq_b200 import ops
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = q.quantize(mu, lv, grid, logvar=True, outputs=ops.OUT_QIDX)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    counts = torch.zeros((16, bench.C, 2047), dtype=torch.int64, device=dev)
    for i in range(16):
        ops.symbol_histogram(out["qidx"][i], 10, counts[i])
    torch.cuda.synchronize(); t2 = time.perf_counter()
    h = counts.cpu()
    t3 = time.perf_counter()
    h32 = counts.to(torch.int32)
    pin = torch.empty(h32.shape, dtype=torch.int32, pin_memory=True)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    pin.copy_(h32, non_blocking=True); torch.cuda.synchronize()
    t5 = time.perf_counter()
    print("search (16 lambdas, qidx out) %.2f ms, 16 histograms %.2f ms, int64 .cpu() %.2f ms, int32 pinned copy %.2f ms (alloc %.2f ms)"
          % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t5 - t4), 1e3 * (t4 - t3)))
