#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference (mandt-lab/vbq, mounted
read-only at /root/reference) in the build container.

  python tests/golden/gen_golden.py            # rewrites tests/golden/*.npz

* `utils.py` (NumPy/numba half) is imported as is.
* `quantizer.py`, `learned_prior.py`, `vae_models.py` are imported as is on top of `oracle/tf_shim` (a NumPy
  stand-in for the TF-1.15 eager ops they call; TensorFlow itself cannot be installed here).
* the notebook cells that build the code points and define `compress_coordinates`
  (word-embeddings/compress-trained-word-embeddings.ipynb, cells 26 and 28) are exec'd verbatim.
Nothing from /root/reference is copied into the repository; only inputs and outputs are stored."""
import io
import json
import os
import sys
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VBQ_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
sys.path.insert(0, os.path.join(REF, "img-compression"))
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_vbq")

import tensorflow as tf          # noqa: E402  (the shim)
import utils as ref_utils        # noqa: E402
import learned_prior as ref_lp   # noqa: E402
import quantizer as ref_q        # noqa: E402
import vae_models as ref_vm      # noqa: E402


class FakeVAE:
    """encode() returns stored (mean, logvar); decode() is a fixed linear map, so quantizer.compress can run."""

    def __init__(self, means, logvars):
        self.means, self.logvars = tf.constant(means), tf.constant(logvars)

    def encode(self, X):
        return self.means, self.logvars

    def decode(self, z):
        z = np.asarray(z)                                   # (B, H', W', C) -> (B, 16H', 16W', 3)
        m = (0.5 + 0.01 * z.mean(axis=-1, keepdims=True)).astype(np.float32)
        return tf.constant(m.repeat(16, axis=1).repeat(16, axis=2).repeat(3, axis=3))


def learned_case(name, C, N, shape, lambs, seed, factor_std):
    rng = np.random.default_rng(seed)
    np.random.seed(seed)                                   # the shim's initializers draw from np.random
    prior = ref_lp.BMSHJ2018Prior(C, dims=(3, 3, 3), init_scale=10.)
    if factor_std > 0:                                     # exercise the tanh gates: perturb the transformed factors
        prior._factors = [tf.constant(np.tanh(factor_std * rng.standard_normal(np.asarray(f).shape)).astype(np.float32))
                          for f in prior._factors]
    q = ref_q.ChannelwisePriorCDFQuantizer(C, N)
    with contextlib.redirect_stdout(io.StringIO()):
        q.build_code_points(prior)
    table = np.asarray(q.all_code_points)
    # latents: means spread over the prior's support (incl. far tails and exact code points), logvar ~ N(-3, 1.5^2)
    B = int(np.prod(shape))
    srt = np.asarray(q.code_points_by_channel)
    u = rng.uniform(0, 1, (B, C))
    pos = u * (srt.shape[1] - 1)
    lo = np.floor(pos).astype(int)
    frac = (pos - lo).astype(np.float32)
    means = np.stack([srt[c, lo[:, c]] * (1 - frac[:, c]) + srt[c, np.minimum(lo[:, c] + 1, srt.shape[1] - 1)] * frac[:, c]
                      for c in range(C)], axis=1).astype(np.float32)
    means[0] = srt[:, 0] - 7.0
    means[1] = srt[:, -1] + 7.0
    means[2] = table[:, 0]
    means[3] = table[:, -1]
    means[4] = table[:, 2 ** N - 1]
    means[5] = table[:, 5 % table.shape[1]]
    logvars = rng.normal(-3.0, 1.5, (B, C)).astype(np.float32)
    means_l, logvars_l = means.reshape(shape + (C,)), logvars.reshape(shape + (C,))
    stds = (np.exp(logvars) ** np.float32(0.5)).astype(np.float32)

    out = dict(C=C, N=N, lambs=np.array(lambs, dtype=np.float64), table=table,
               sorted_table=srt, means=means_l, logvars=logvars_l, stds=stds,
               cdf_in=means, cdf_out=np.asarray(prior.cdf(tf.constant(means), stop_gradient=True)))
    for k in range(4):
        out["matrix%d" % k] = np.asarray(prior._matrices[k])
        out["bias%d" % k] = np.asarray(prior._biases[k])
        if k < 3:
            out["factor%d" % k] = np.asarray(prior._factors[k])
    # brackets (quantizer.py:65-80)
    left, right = q.get_all_N_bit_intervals(tf.constant(means))
    out["left"], out["right"] = np.asarray(left), np.asarray(right)
    # raw code lengths (quantizer.py:156-188)
    Zh, nb = q.compress_batch_channel_latents(tf.constant(means), tf.constant(stds), lambs)
    for i, l in enumerate(lambs):
        out["raw_zhat_%d" % i] = np.asarray(Zh[l])
        out["raw_bits_%d" % i] = np.asarray(nb[l])
    # two-pass entropy models (quantizer.py:82-150), then the corrected-length mode
    X = np.zeros((shape[0], 16 * shape[1], 16 * shape[2], 3), dtype=np.float32)
    vae = FakeVAE(means_l, logvars_l)
    q.build_entropy_models(X, vae, lambs, add_n_smoothing=1)
    for i, l in enumerate(lambs):
        out["rcl_%d" % i] = np.asarray(q.raw_code_length_entropy_models[l])
        out["em_%d" % i] = np.asarray(q.entropy_models[l])
    Zh, nb = q.compress_batch_channel_latents(tf.constant(means), tf.constant(stds), lambs)
    res = q.compress(X, vae, lambs, clip=True)
    for i, l in enumerate(lambs):
        out["cl_zhat_%d" % i] = np.asarray(Zh[l])
        out["cl_bits_%d" % i] = np.asarray(nb[l])
        for key in ("Z_hat", "raw_num_bits", "num_bits_cl", "num_bits", "X_hat"):
            out["compress_%s_%d" % (key, i)] = np.asarray(res[key][l])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "table", table.shape, "latents", means_l.shape)


def utils_case():
    """The NumPy/numba half of utils.py: xi-space brackets, Algorithm 1 (`encode_vectorized`), the generic
    operator `batch_quantize_indep_dims(backend=np)` and the docstring known-answer."""
    from scipy.stats import norm
    out = {}
    out["kat_interval"] = np.array(ref_utils.get_n_bit_interval(0.4375, 2))         # utils.py:31-32 -> [.375,.625]
    x = np.array([0, .03, .0625, .4375, .5, .97, 1.0])
    N = 4
    L, R = np.empty((N + 1, len(x))), np.empty((N + 1, len(x)))
    ref_utils.get_all_N_bit_intervals(x, N, L, R)
    out["xi_x"], out["xi_left"], out["xi_right"] = x, L, R
    out["bin_floats_3"] = np.array(ref_utils.n_bit_binary_floats(3))
    rng = np.random.default_rng(42)
    K, N = 4000, 10
    std = 1.2329
    mu = rng.normal(-0.08, std, K)
    sig = np.exp(rng.normal(np.log(0.04), 0.7, K))
    prior_std = float(np.sqrt(np.mean(mu ** 2)))
    for i, lamb in enumerate([0.01, 1.0, 30.0]):
        fun = lambda z: -0.5 * ((z - mu) / sig) ** 2                                 # noqa: E731
        r = ref_utils.encode_vectorized(fun, mu, lamb, squash=lambda z: norm.cdf(z, scale=prior_std),
                                        unsquash=lambda xi: norm.ppf(xi, scale=prior_std), max_bits_per_coord=N)
        out["ev_zhat_%d" % i], out["ev_bits_%d" % i], out["ev_xi_%d" % i] = r["z_hat"], r["num_bits"], r["xi_hat"]
    out["ev_mu"], out["ev_sigma"], out["ev_prior_std"], out["ev_lambs"] = mu, sig, prior_std, np.array([0.01, 1.0, 30.0])
    # generic operator on explicit candidates, NumPy backend (utils.py:363-423)
    B, Kc, M = 50, 6, 9
    P = np.sort(rng.normal(0, 2, (M, B, Kc)).astype(np.float32), axis=0)
    Lc = rng.integers(0, 8, (M, B, Kc)).astype(np.int32)
    loc = rng.normal(0, 1, (B, Kc)).astype(np.float32)
    scale = np.exp(rng.normal(-1, 0.5, (B, Kc))).astype(np.float32)
    fun = lambda z: np.float32(-0.5) * ((z - loc) / scale) ** 2                      # noqa: E731
    lambs = [0.1, 1.0]
    Zh, nb = ref_utils.batch_quantize_indep_dims((B, Kc), P, Lc, fun, lambs, backend=np)
    out.update(bq_P=P, bq_L=Lc, bq_loc=loc, bq_scale=scale, bq_lambs=np.array(lambs))
    for i, l in enumerate(lambs):
        out["bq_zhat_%d" % i], out["bq_bits_%d" % i] = Zh[l], nb[l]
    np.savez_compressed(os.path.join(HERE, "utils_numpy.npz"), **out)
    print("utils_numpy")


def notebook_case():
    """Cells 26 (code points) and 28 (compress_coordinates) of the notebook, exec'd verbatim."""
    nb = json.load(open(os.path.join(REF, "word-embeddings", "compress-trained-word-embeddings.ipynb")))
    cells = ["".join(c["source"]) for c in nb["cells"] if c["cell_type"] == "code"]
    cp_cell = next(c for c in cells if c.startswith("max_codepoint_length"))
    cc_cell = next(c for c in cells if c.startswith("def compress_coordinates"))
    rng = np.random.default_rng(1)
    V, K = 300, 100                                                                   # SURVEY.md §8d C1 statistics
    vecs_u = rng.normal(-0.08, 1.2329, (V, K)).astype(np.float32)
    stds_u = np.exp(rng.normal(np.log(0.04), 0.7, (V, K))).astype(np.float32)
    import scipy.stats
    ns = dict(np=np, scipy=__import__("scipy"), vecs_u=vecs_u)
    ns["empirical_std"] = np.sqrt(np.mean(vecs_u.ravel() ** 2))                       # cell 25
    exec(cp_cell, ns)
    exec(cc_cell, ns)
    out = dict(means=vecs_u, stds=stds_u, empirical_std=float(ns["empirical_std"]),
               codepoints=ns["codepoints"], lengths=ns["lengths"], betas=np.array([0.01, 1.0, 300.0]))
    for i, beta in enumerate([0.01, 1.0, 300.0]):
        with contextlib.redirect_stdout(io.StringIO()):
            optima, _ = ns["compress_coordinates"](vecs_u, stds_u, beta)
        out["optima_%d" % i] = optima
    np.savez_compressed(os.path.join(HERE, "notebook_embeddings.npz"), **out)
    print("notebook_embeddings", vecs_u.shape)


if __name__ == "__main__":
    lambs16 = [float(l) for l in 2 ** np.linspace(-8, 7, 16)]
    learned_case("learned_c6_n10", C=6, N=10, shape=(2, 8, 12), lambs=[lambs16[0], lambs16[7], lambs16[9], lambs16[15]],
                 seed=3, factor_std=0.0)
    learned_case("learned_c20_n6_gated", C=20, N=6, shape=(1, 10, 13), lambs=[2.0 ** -6, 0.5, 8.0], seed=4,
                 factor_std=0.5)
    utils_case()
    notebook_case()
