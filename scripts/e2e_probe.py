"""Where does the host pipeline's time go?  Runs vbq_quantize_host with different output sets / chunk sizes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from vbq_b200 import ops

dev = torch.device("cuda", 0)
prior, q = bench.make_prior_and_quantizer(dev)
mu, sigma = bench.make_batch(prior, 7, dev)
pen, length = q._length_tables([0.5])
h_mu, h_sg = mu.cpu().pin_memory(), sigma.cpu().pin_memory()
R, C = mu.shape
h_q = torch.empty((1, R, C), dtype=torch.int32).pin_memory()
h_b = torch.empty((1, R, C), dtype=torch.float32).pin_memory()
h_t = torch.empty((1, 4), dtype=torch.float64).pin_memory()


def run(chunk, outs, kw, n=20):
    pipe = ops.HostPipeline(C, 10, 1, chunk, outs, device=dev)
    f = lambda: pipe.run(h_mu, h_sg, q.all_code_points, q._packed, pen, length, None, flags=2, **kw)
    f(); f()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    dt = (time.perf_counter() - t0) / n
    pipe.close()
    return dt * 1e3


for chunk in (4608, 9216, 18432, 36864):
    full = run(chunk, ops.OUT_QIDX | ops.OUT_BITS | ops.OUT_TOTALS, dict(qidx=h_q, bits=h_b, totals=h_t))
    one = run(chunk, ops.OUT_QIDX | ops.OUT_TOTALS, dict(qidx=h_q, totals=h_t))
    none = run(chunk, ops.OUT_TOTALS, dict(totals=h_t))
    print("chunk %6d: full %.3f ms | one output %.3f ms | totals only (H2D + kernel) %.3f ms" % (chunk, full, one, none))
