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
for cta in (0, 70, int(np.argmax(dur))):
    t0 = w[cta, 1]
    P = w[cta, 8:264].reshape(64, 4)
    Cc = w[cta, 264:264 + 3 * 64 * 3].reshape(3, 64, 3)
    nt = int((P[:, 0] > 0).sum())
    print("CTA %d: %d tiles, consumer time %.2f us" % (cta, nt, dur[cta]))
    for wi, name in enumerate(("warp 0", "warp W-1", "warp W/2")):
        c = Cc[wi][:nt]
        waits = c[:, 1] - c[:, 0]
        work = c[:, 2] - c[:, 1]
        print("   %s: wait per tile mean %.0f ns (max %.0f), work per tile mean %.0f ns; wait share %.1f%%" % (
            name, waits.mean(), waits.max(), work.mean(), 100 * waits.sum() / (waits.sum() + work.sum())))
    for t in range(min(nt, 14)):
        print("   tile %2d: load issued %.2f | full seen w0 %.2f wL %.2f | end w0 %.2f wL %.2f wM %.2f | done seen %.2f stored %.2f slot read %.2f" % (
            t, (P[t, 0] - t0) / 1e3, (Cc[0, t, 1] - t0) / 1e3, (Cc[1, t, 1] - t0) / 1e3, (Cc[0, t, 2] - t0) / 1e3,
            (Cc[1, t, 2] - t0) / 1e3, (Cc[2, t, 2] - t0) / 1e3, (P[t, 1] - t0) / 1e3, (P[t, 2] - t0) / 1e3, (P[t, 3] - t0) / 1e3))
