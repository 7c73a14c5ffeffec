"""Development: the N = 16 / C = 320 case through the different call levels."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vbq_b200
from vbq_b200 import ops
dev = torch.device("cuda", 0)
C5, N5, rows5 = 320, 16, 32 * 128 * 128
prior5 = vbq_b200.BMSHJ2018Prior(C5, device=dev, seed=5)
q5 = vbq_b200.ChannelwisePriorCDFQuantizer(C5, N5, device=dev)
q5.build_code_points(prior5)
g = torch.Generator(device=dev); g.manual_seed(1)
u = torch.rand((rows5, C5), generator=g, device=dev, dtype=torch.float64) * 0.998 + 0.001
mu = prior5.inverse_cdf(u).contiguous(); del u
sg = torch.exp(0.5 * (torch.randn((rows5, C5), generator=g, device=dev) * 1.5 - 3.0)).contiguous()
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
for lamb in (0.5, 2.0 ** -8):
    pen, length = q5._length_tables([lamb])
    qidx = torch.empty((1, rows5, C5), dtype=torch.int32, device=dev); bits = torch.empty((1, rows5, C5), dtype=torch.float32, device=dev)
    tot = torch.zeros((1, 4), dtype=torch.float64, device=dev); ws = ops.quantize_workspace(1, dev)
    fl = ops.search_flags([lamb])
    a = timeit(lambda: ops.quantize_into(mu, sg, q5.all_code_points, q5._packed, pen, length, None, N5, qidx=qidx, bits=bits, flags=fl))
    b = timeit(lambda: ops.quantize_into(mu, sg, q5.all_code_points, q5._packed, pen, length, None, N5, qidx=qidx, bits=bits, totals=tot, workspace=ws, flags=fl))
    c = timeit(lambda: q5.quantize(mu, sg, [lamb], outputs=ops.OUT_QIDX | ops.OUT_BITS | ops.OUT_TOTALS))
    d = timeit(lambda: q5.quantize(mu, sg, [lamb], outputs=ops.OUT_QIDX | ops.OUT_BITS))
    print("lambda %g: into no totals %.2f ms | into totals %.2f | facade totals %.2f | facade no totals %.2f  (%.1f G coords/s best)" % (lamb, a, b, c, d, rows5 * C5 / min(a, b, c, d) / 1e6))
