"""Development: host-side profile of the per-image compress_latents call (tests/bench_per_image.py)."""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda", 0)
prior, q = bench.make_prior_and_quantizer(dev)
lambs = [float(l) for l in 2 ** np.linspace(-8, 7, 16)]
mu, sigma = bench.make_batch(prior, 3, dev)
means = mu.cpu().numpy().reshape(bench.IMAGES, bench.H, bench.W, bench.C)
logvars = (2 * torch.log(sigma)).cpu().numpy().reshape(means.shape)
q.build_entropy_models_from_latents(means, logvars, lambs, add_n_smoothing=1)
for i in range(3):
    q.compress_latents(means[i:i + 1], logvars[i:i + 1], lambs)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(24):
    q.compress_latents(means[i:i + 1], logvars[i:i + 1], lambs)
print("ms per image: %.3f" % ((time.perf_counter() - t0) / 24 * 1e3))
pr = cProfile.Profile()
pr.enable()
for i in range(24):
    q.compress_latents(means[i:i + 1], logvars[i:i + 1], lambs)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
