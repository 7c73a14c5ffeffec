"""Development: per-warp timeline of the bisection kernel (library built with -DVBQ_TRACE)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from vbq_b200 import ops
dev = torch.device("cuda", 0)
prior, q = bench.make_prior_and_quantizer(dev)
pen, length = q._length_tables([bench.LAMB])
mu, sigma = bench.make_batch(prior, 1000, dev)
qidx = torch.empty((1, bench.ROWS, bench.C), dtype=torch.int32, device=dev)
bits = torch.empty((1, bench.ROWS, bench.C), dtype=torch.float32, device=dev)
tot = torch.zeros((1, 4), dtype=torch.float64, device=dev)
ws = torch.zeros(1 << 20, dtype=torch.float64, device=dev)
import time
evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
NIT = int(os.environ.get("NIT", "3"))
for it in range(NIT):
    if it == NIT - 2: evs[0].record()
    if it == NIT - 1: evs[1].record()
    ops.quantize_into(mu, sigma, q.all_code_points, q._packed, pen, length, None, bench.N_BITS, qidx=qidx, bits=bits,
                      totals=tot, workspace=ws, flags=2 | 256)
evs[2].record()
torch.cuda.synchronize()
print("NIT", NIT, "event time of the last two launches us: %.2f %.2f" % (1e3 * evs[0].elapsed_time(evs[1]), 1e3 * evs[1].elapsed_time(evs[2])))
w = ws.cpu().numpy()
base = 256 // 8 + 1024 * 4
tr = w[base:base + 148 * 32 * 5].reshape(148, 32, 5)
nw = int((tr[0, :, 0] > 0).sum())
tr = tr[:, :nw]
t0 = tr[:, :, 0].min()
start = tr[:, 0, 0] - t0
wend = tr[:, :, 1] - t0
cend = tr[:, 0, 2] - t0
print("warps/CTA", nw, "kernel span us %.2f" % (cend.max() / 1e3))
print("CTA start us: min %.2f max %.2f" % (start.min() / 1e3, start.max() / 1e3))
print("CTA end us: min %.2f median %.2f max %.2f" % (cend.min() / 1e3, np.median(cend) / 1e3, cend.max() / 1e3))
print("warp end spread inside CTA us: median %.2f max %.2f" % (np.median(wend.max(1) - wend.min(1)) / 1e3, (wend.max(1) - wend.min(1)).max() / 1e3))
order = np.argsort(cend)
for i in list(order[:5]) + list(order[-12:]):
    print("cta %3d sm %3d segs %d start %.2f end %.2f warp-end min %.2f" % (i, tr[i, 0, 4], tr[i, 0, 3], start[i] / 1e3, cend[i] / 1e3, wend[i].min() / 1e3))
cyc = (tr[:, 0, 3] - np.floor(tr[:, 0, 3])) * 1e9
print("SM clock MHz from clock64/globaltimer: median %.0f" % np.median(cyc / (tr[:, 0, 2] - tr[:, 0, 0]) * 1e3))
segs = np.floor(tr[:, 0, 3])
for s_ in (1, 2):
    m = segs == s_
    if m.any():
        print("segments", s_, "n", int(m.sum()), "mean end %.2f" % (cend[m].mean() / 1e3))
