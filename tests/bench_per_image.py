#!/usr/bin/env python
"""The reference's evaluation pattern (utils.py:535-553): one Kodak-shaped image per call, the 16-lambda grid of
post_process.py:115, host (NumPy) latents in, NumPy results out, through ChannelwisePriorCDFQuantizer.compress_latents.
Prints seconds per image for the CUDA path and for the CPU oracle port."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                        # noqa: E402
from oracle import vbq_oracle as O  # noqa: E402

dev = torch.device("cuda", 0)
prior, q = bench.make_prior_and_quantizer(dev)
lambs = [float(l) for l in 2 ** np.linspace(-8, 7, 16)]
mu, sigma = bench.make_batch(prior, 3, dev)
means = mu.cpu().numpy().reshape(bench.IMAGES, bench.H, bench.W, bench.C)
logvars = (2 * torch.log(sigma)).cpu().numpy().reshape(means.shape)
q.build_entropy_models_from_latents(means, logvars, lambs, add_n_smoothing=1)      # two-pass fit on the 24 images
# The results are NumPy views of pinned buffers (no extra host copy); the caller keeps one result alive while the next call
# runs, so the pinned allocator needs two generations of them: warm up the same way (cudaHostAlloc takes milliseconds).
for i in range(8):
    out = q.compress_latents(means[i:i + 1], logvars[i:i + 1], lambs)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(bench.IMAGES):
    out = q.compress_latents(means[i:i + 1], logvars[i:i + 1], lambs)
t_gpu = (time.perf_counter() - t0) / bench.IMAGES

oq = O.QuantizerNP(bench.C, bench.N_BITS)
oq.set_code_points(q.all_code_points.cpu().numpy())
oq.raw_code_length_entropy_models = q.raw_code_length_entropy_models
oq.entropy_models = q.entropy_models
t0 = time.perf_counter()
n_cpu = 3
for i in range(n_cpu):
    ref = oq.compress_latents(means[i:i + 1], logvars[i:i + 1], lambs)
t_cpu = (time.perf_counter() - t0) / n_cpu
same = float(np.mean([np.mean(out_ == ref_) for out_, ref_ in
                      [(q.compress_latents(means[2:3], logvars[2:3], lambs)["Z_hat"][l], ref["Z_hat"][l]) for l in lambs]]))
coords = bench.H * bench.W * bench.C * len(lambs)
print(json.dumps({"case": "per-image compress_latents, 16 lambdas, host in/out", "s_per_image_gpu": t_gpu,
                  "s_per_image_cpu_oracle": t_cpu, "speedup": t_cpu / t_gpu,
                  "coord_lambda_per_s_gpu": coords / t_gpu, "z_hat_agreement_with_oracle": same}))
