"""Python replica of vbq_quantize_host's stream pattern with timing events, to see the per-chunk timeline."""
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
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 4608
nch = (R + chunk - 1) // chunk
s_in, s_k, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
slots = [dict(mu=torch.empty((chunk, C), device=dev), sg=torch.empty((chunk, C), device=dev),
              q=torch.empty((1, chunk, C), dtype=torch.int32, device=dev),
              b=torch.empty((1, chunk, C), device=dev)) for _ in range(3)]


def once(timed=False):
    E = lambda: torch.cuda.Event(enable_timing=True)
    ev = [dict(i0=E(), i1=E(), k1=E(), o0=E(), o1=E()) for _ in range(nch)]
    t0 = E(); t0.record(s_in)
    for k in range(nch):
        a, b = k * chunk, min(R, (k + 1) * chunk)
        s = slots[k % 3]
        with torch.cuda.stream(s_in):
            if k >= 3: s_in.wait_event(ev[k - 3]["o1"])
            ev[k]["i0"].record(s_in)
            s["mu"][:b - a].copy_(h_mu[a:b], non_blocking=True); s["sg"][:b - a].copy_(h_sg[a:b], non_blocking=True)
            ev[k]["i1"].record(s_in)
        with torch.cuda.stream(s_k):
            s_k.wait_event(ev[k]["i1"])
            if k >= 3: s_k.wait_event(ev[k - 3]["o1"])
            ops.quantize_into(s["mu"][:b - a], s["sg"][:b - a], q.all_code_points, q._packed, pen, length, None, 10,
                              qidx=s["q"][:, :b - a], bits=s["b"][:, :b - a], flags=2)
            ev[k]["k1"].record(s_k)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev[k]["k1"])
            ev[k]["o0"].record(s_out)
            h_q[0, a:b].copy_(s["q"][0, :b - a], non_blocking=True); h_b[0, a:b].copy_(s["b"][0, :b - a], non_blocking=True)
            ev[k]["o1"].record(s_out)
    torch.cuda.synchronize()
    if timed:
        for k in range(nch):
            e = ev[k]
            print("chunk %d: H2D %6.0f-%6.0f us | kernel done %6.0f | D2H %6.0f-%6.0f" % (
                k, t0.elapsed_time(e["i0"]) * 1e3, t0.elapsed_time(e["i1"]) * 1e3, t0.elapsed_time(e["k1"]) * 1e3,
                t0.elapsed_time(e["o0"]) * 1e3, t0.elapsed_time(e["o1"]) * 1e3))


# quantize_into needs contiguous views: slices of the leading rows are contiguous
once(); once()
t = time.perf_counter()
for _ in range(10): once()
print("python replica: %.3f ms/step" % ((time.perf_counter() - t) / 10 * 1e3))
once(True)
