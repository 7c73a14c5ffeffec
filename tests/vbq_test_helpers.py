"""Shared generators / comparison helpers for the tests (SURVEY.md §8d synthetic inputs)."""
import numpy as np

from oracle import vbq_oracle as O


def make_prior(C, seed, factor_std=0.5, init_scale=10.0):
    return O.LearnedPriorNP.init(C, init_scale=init_scale, rng=np.random.default_rng(seed), factor_std=factor_std)


def make_latents(prior, rows, seed, edge_cases=True, table=None):
    """mu = F_c^-1(U(0.001, 0.999)), logvar ~ N(-3, 1.5^2)  (SURVEY.md §8d, config C2)."""
    rng = np.random.default_rng(seed)
    C = prior.channels
    u = rng.uniform(0.001, 0.999, (rows, C))
    mu = prior.inverse_cdf_f64(u).astype(np.float32)
    logvar = rng.normal(-3.0, 1.5, (rows, C)).astype(np.float32)
    if edge_cases and rows >= 16:
        mu[0] = -1e4          # far below every code point
        mu[1] = 1e4           # far above every code point
        mu[2] = 0.0
        if table is not None:  # exactly on code points of several depths, and one ulp around them
            Q = table.shape[1]
            for r, h in zip(range(3, 12), [0, 1, 2, 5, Q // 2, Q - 1, Q - 2, (Q - 1) // 2, 3]):
                mu[r] = table[:, h % Q]
            mu[12] = np.nextafter(table[:, Q - 1], np.float32(np.inf))
            mu[13] = np.nextafter(table[:, (Q - 1) // 2], np.float32(-np.inf))
            mu[14] = np.nextafter(table[:, 0], np.float32(np.inf))
            mu[15] = np.nextafter(table[:, 0], np.float32(-np.inf))
    sigma = (np.exp(logvar) ** np.float32(0.5)).astype(np.float32)
    return mu, sigma, logvar


def classify_mismatches(det, zhat_k, bits_k, rel=1e-6):
    """Compare kernel output with the oracle's per-lambda details (QuantizerNP(..., details=True)).

    Returns (n_mismatch, n_outside_tie_band): a mismatch is a coordinate whose z_hat or code length differs; it is
    inside the tie band when the oracle's score of the kernel's pick is within `rel` relative of the maximum."""
    P, scores, k_o = det["P"], det["scores"], det["cand"]
    zo = np.take_along_axis(P, k_o[None], axis=0)[0]
    mism = (zo != zhat_k) | np.isnan(zhat_k)
    n_mis = int(mism.sum())
    if n_mis == 0:
        return 0, 0
    best = np.take_along_axis(scores, k_o[None], axis=0)[0]
    same = (P == zhat_k[None])                       # candidates equal to the kernel's pick
    s_k = np.where(same, scores, -np.inf).max(axis=0)
    gap = np.abs(best - s_k)
    outside = mism & ~(gap <= rel * np.maximum(np.abs(best), np.finfo(np.float32).tiny))
    return n_mis, int(outside.sum())
