"""Development: a few launches of the both-ends sweep kernel on the Kodak batch (for ncu): 16 lambdas, z_hat + code length."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, vbq_b200
from vbq_b200 import ops
dev = torch.device("cuda", 0)
prior, q0 = bench.make_prior_and_quantizer(dev)
q = vbq_b200.ChannelwisePriorCDFQuantizer(bench.C, bench.N_BITS, device=dev)
q.set_code_points(q0.all_code_points)
mu, sg = bench.make_batch(prior, 50, dev)
grid = [float(l) for l in 2.0 ** np.linspace(-8, 7, 16)]
q.build_entropy_models_from_latents(mu, (2.0 * torch.log(sg)).contiguous(), grid, add_n_smoothing=1.0)
outs = ops.OUT_TOTALS if os.environ.get("VBQ_TOTALS_ONLY") else ops.OUT_ZHAT | ops.OUT_BITS | ops.OUT_TOTALS
for i in range(3):
    q.quantize(mu, sg, grid, outputs=outs, entropy_bits=bool(os.environ.get("VBQ_WITH_EM")))
torch.cuda.synchronize()
