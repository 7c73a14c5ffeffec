"""Development: where the time of the facade's host path goes (per-call wall clock, alternating inputs, kept outputs)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
dev = torch.device("cuda", 0)
bench.bind_to_gpu_numa_node(0)
prior, q = bench.make_prior_and_quantizer(dev)
ins = []
for s in range(2):
    mu, sigma = bench.make_batch(prior, s, dev)
    ins.append((mu.cpu().pin_memory().numpy(), sigma.cpu().pin_memory().numpy()))
state = {}
ts = []
for i in range(12):
    t0 = time.perf_counter()
    Z, B = q.compress_batch_channel_latents(ins[i % 2][0], ins[i % 2][1], [bench.LAMB])
    state["last"] = (Z[bench.LAMB], B[bench.LAMB])
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
print("per call ms:", ["%.2f" % t for t in ts])
