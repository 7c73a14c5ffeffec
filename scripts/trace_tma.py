"""Development: timeline of the TMA pipeline kernel (library built with VBQ_BUILD_DEFINES=-DVBQ_TRACE)."""
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
wss = [torch.zeros(1 << 21, dtype=torch.float64, device=dev) for _ in range(3)]
for it in range(9):
    ops.quantize_into(mu, sigma, q.all_code_points, q._packed, pen, length, None, bench.N_BITS, qidx=qidx, bits=bits,
                      totals=tot, workspace=wss[it % 3], flags=2 | 256)
torch.cuda.synchronize()
base = 256 // 8 + 1024 * 4
W = [w.cpu().numpy()[base:base + 148 * 1024].reshape(148, 1024) for w in wss]
T0 = min(w[:, 0].min() for w in W)
for k, w in enumerate(W):
    f = lambda col: "%.2f..%.2f" % ((w[:, col].min() - T0) / 1e3, (w[:, col].max() - T0) / 1e3)
    print("launch %d: CTA entry %s | after pdl_wait %s | consumers end (warp0) %s | producer end %s | exit %s us" % (
        k + 6, f(0), f(1), f(2), f(3), f(4)))
w = W[2]
dur = (w[:, 2] - w[:, 1]) / 1e3
print("consumer duration per CTA: min %.2f max %.2f mean %.2f; slowest CTAs %s" % (dur.min(), dur.max(), dur.mean(), np.argsort(dur)[-8:]))

# per-CTA: ranges (replicating row_cut of quantize_tma.cu) and tiles vs duration
SW, rows4, ng, grid = 448, (bench.ROWS + 3) // 4 * 4, (bench.C + 15) // 16, 148
def row_cut(v):
    vg = rows4 + SW
    g = min((v + SW) // vg, ng)
    o = v + SW - g * vg
    return g * rows4 + (((o - SW) // 4 * 4) if o > SW else 0)
vtotal = (rows4 + SW) * ng - SW
info = []
for b in range(grid):
    p0, p1 = row_cut(vtotal * b // grid), row_cut(vtotal * (b + 1) // grid)
    nr = (p1 - 1) // rows4 - p0 // rows4 + 1
    P = w[b, 8:264].reshape(64, 4)
    first_full = (P[0, 1] - w[b, 1]) / 1e3
    info.append((b, nr, p1 - p0, dur[b], (w[b, 3] - w[b, 2]) / 1e3, (w[b, 4] - w[b, 3]) / 1e3))
info = np.array(info)
for nr in (1, 2):
    m = info[:, 1] == nr
    if m.any():
        print("CTAs with %d range(s): %d, rows %.0f, consumer time mean %.2f min %.2f max %.2f; producer tail %.2f; exit tail %.2f" % (
            nr, m.sum(), info[m, 2].mean(), info[m, 3].mean(), info[m, 3].min(), info[m, 3].max(), info[m, 4].mean(), info[m, 5].mean()))
order = np.argsort(info[:, 3])
print("fastest:", [(int(info[i, 0]), int(info[i, 1]), round(info[i, 3], 1)) for i in order[:10]])
print("slowest:", [(int(info[i, 0]), int(info[i, 1]), round(info[i, 3], 1)) for i in order[-10:]])
