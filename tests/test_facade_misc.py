"""Facade pieces the hot path leans on but that had no test of their own (VERDICT r1 "close the small holes"):
pickling as in post_process.py:106-107,163-164, `code_points_by_bits` (quantizer.py:40-46), the Gaussian priors and the
GaussianVAE contract of vae_models.py:14-72, and curry_normal_logpdf with its constant (utils.py:307-327)."""
import io
import pickle

import numpy as np
import pytest
import torch

import vbq_b200
from vbq_b200 import ops, utils
from vbq_test_helpers import make_latents, make_prior

pytestmark = pytest.mark.gpu


def _quantizer_with_models(C=20, N=6, rows=3000, lambs=(0.05, 1.0)):
    pr = make_prior(C, 5)
    q = vbq_b200.ChannelwisePriorCDFQuantizer(C, N)
    q.set_code_points(ops.build_code_points_learned(torch.from_numpy(pr.packed()).cuda(), N))
    mu, sigma, logvar = make_latents(pr, rows, 9, table=q.all_code_points.cpu().numpy())
    q.build_entropy_models_from_latents(mu, logvar, list(lambs), add_n_smoothing=0.1)
    return q, mu, logvar


def test_pickle_round_trip_keeps_every_result():
    lambs = [0.05, 1.0]
    q, mu, logvar = _quantizer_with_models(lambs=lambs)
    buf = io.BytesIO()
    pickle.dump(q, buf)                                   # post_process.py:106-107
    state = pickle.loads(pickle.dumps(q.__getstate__()))
    assert all(not isinstance(v, torch.Tensor) for v in state.values()), "the pickle must not carry CUDA tensors"
    q2 = pickle.loads(buf.getvalue())                     # post_process.py:163-164
    assert q2.lambs == q.lambs and q2.max_bits_per_coord == q.max_bits_per_coord
    assert torch.equal(q2.all_code_points, q.all_code_points)
    assert torch.equal(q2.code_points_by_channel, q.code_points_by_channel)
    a = q.compress_latents(mu[None], logvar[None], lambs)
    b = q2.compress_latents(mu[None], logvar[None], lambs)
    for key in ('Z_hat', 'raw_num_bits', 'num_bits_cl', 'num_bits'):
        for lamb in lambs:
            assert np.array_equal(a[key][lamb], b[key][lamb]), (key, lamb)
    for lamb in lambs:
        assert np.array_equal(q.entropy_models[lamb], q2.entropy_models[lamb])
        assert np.array_equal(q.raw_code_length_entropy_models[lamb], q2.raw_code_length_entropy_models[lamb])


def test_code_points_by_bits_is_the_heap_order_split():
    C, N = 5, 4
    pr = make_prior(C, 3)
    q = vbq_b200.ChannelwisePriorCDFQuantizer(C, N)
    q.set_code_points(ops.build_code_points_learned(torch.from_numpy(pr.packed()).cuda(), N))
    by_bits = q.code_points_by_bits
    table = q.all_code_points.cpu().numpy()
    assert len(by_bits) == C and all(len(by_bits[c]) == N + 1 for c in range(C))
    for c in range(C):
        flat = []
        for n in range(N + 1):
            pts = by_bits[c][n].cpu().numpy()
            assert pts.shape == (2 ** n,)
            assert np.all(np.diff(pts) > 0)               # each level ascends
            flat.append(pts)
        assert np.array_equal(np.concatenate(flat), table[c])        # quantizer.py:40-46: consecutive heap slices
        assert np.array_equal(np.sort(table[c]), q.code_points_by_channel[c].cpu().numpy())


def test_gaussian_priors_match_scipy():
    from scipy.stats import norm
    xi = np.array(utils.all_bin_floats(7))
    z = vbq_b200.vae_models.StandardGaussianPrior.inverse_cdf(xi[:, None])
    assert z.dtype == np.float64 and np.allclose(z[:, 0], norm.ppf(xi), rtol=1e-12, atol=1e-14)   # vae_models.py:23-25
    mean = np.array([0.3, -2.0, 10.0, 0.0])
    std = np.array([1.0, 0.01, 30.0, 2.5])
    pr = vbq_b200.vae_models.FactoredGaussianPrior(mean, std)
    xi_rep = np.repeat(xi[:, None], 4, axis=1)
    want = norm.ppf(xi_rep, loc=mean, scale=std)          # vae_models.py:40-43
    got = pr.inverse_cdf(xi_rep)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-13)
    assert np.allclose(pr.logvar, 2 * np.log(std))
    with pytest.raises(AssertionError):
        pr.inverse_cdf(xi_rep[:, :3])
    # torch in, torch out (device), and as the quantizer's prior: the table is float32(norm.ppf) in heap order
    assert isinstance(pr.inverse_cdf(torch.from_numpy(xi_rep)), torch.Tensor)
    q = vbq_b200.ChannelwisePriorCDFQuantizer(4, 7)
    q.build_code_points(pr)
    table = q.all_code_points.cpu().numpy()
    ulp = np.spacing(np.abs(want.T.astype(np.float32)))
    assert np.all(np.abs(table - want.T) <= ulp)
    q2 = vbq_b200.ChannelwisePriorCDFQuantizer(4, 7)      # the generic route (any object with inverse_cdf)
    q2.build_code_points(type("P", (), {"inverse_cdf": staticmethod(lambda x: norm.ppf(x, loc=mean, scale=std))})())
    assert np.array_equal(q2.all_code_points.cpu().numpy(), want.T.astype(np.float32))


def test_gaussian_vae_contract_and_compress():
    C, N, lambs = 16, 6, [0.1, 2.0]
    pr = make_prior(C, 11)
    q = vbq_b200.ChannelwisePriorCDFQuantizer(C, N)
    q.set_code_points(ops.build_code_points_learned(torch.from_numpy(pr.packed()).cuda(), N))
    g = torch.Generator().manual_seed(1)
    Wenc = torch.randn(3, 2 * C, generator=g).cuda()
    Wdec = torch.randn(C, 3, generator=g).cuda()
    vae = vbq_b200.vae_models.GaussianVAE(prior=None, inference_net=lambda x: torch.as_tensor(x).cuda().float() @ Wenc,
                                          generative_net=lambda z: z @ Wdec, decode_sigmoid=True)
    X = torch.rand(2, 5, 7, 3, generator=g).numpy().astype(np.float32)
    mean, logvar = vae.encode(X)                          # vae_models.py:55-58: split along the last axis
    full = torch.as_tensor(X).cuda() @ Wenc
    assert torch.equal(mean, full[..., :C]) and torch.equal(logvar, full[..., C:])
    z = torch.randn(4, 5, 7, C, generator=g).cuda()
    assert torch.equal(vae.decode(z), torch.sigmoid(z @ Wdec))       # vae_models.py:60-70
    vae_nosig = vbq_b200.vae_models.GaussianVAE(None, vae.inference_net, vae.generative_net)
    assert torch.equal(vae_nosig.decode(z), z @ Wdec)
    q.build_entropy_models(X, vae, lambs, add_n_smoothing=0.1)
    out = q.compress(X, vae, lambs)                       # quantizer.py:242-256
    lat = q.compress_latents(mean, logvar, lambs)
    for lamb in lambs:
        assert np.array_equal(out['Z_hat'][lamb], lat['Z_hat'][lamb])
        xh = torch.sigmoid(torch.from_numpy(lat['Z_hat'][lamb]).cuda() @ Wdec).cpu().numpy()
        assert out['X_hat'][lamb].shape == X.shape and np.allclose(out['X_hat'][lamb], np.clip(xh, 0, 1), atol=1e-6)


def test_logpdf_constant_branch():
    from scipy.stats import norm
    rng = np.random.default_rng(0)
    loc = torch.from_numpy(rng.normal(size=(4, 3)).astype(np.float32))
    scale = torch.from_numpy(np.exp(rng.normal(size=(4, 3))).astype(np.float32))
    z = torch.from_numpy(rng.normal(size=(5, 4, 3)).astype(np.float32))
    full = utils.curry_normal_logpdf(loc, scale, ignore_const=False)(z).numpy()          # utils.py:318-324
    assert np.allclose(full, norm.logpdf(z.numpy(), loc.numpy(), scale.numpy()), rtol=2e-5, atol=2e-5)
    kern = utils.curry_normal_logpdf(loc, scale, ignore_const=True)(z).numpy()
    assert np.allclose(kern, -0.5 * ((z.numpy() - loc.numpy()) / scale.numpy()) ** 2, rtol=1e-6)
    # through batch_quantize_indep_dims the constant shifts every candidate of a dimension equally: same optimum
    P = torch.from_numpy(np.sort(rng.normal(size=(3, 9)).astype(np.float32), axis=1))    # (K, M)
    L = torch.arange(9, dtype=torch.int32)[None, :].repeat(3, 1)
    loc_d, scale_d = loc.cuda(), scale.cuda()
    a = utils.batch_quantize_indep_dims((4, 3), P, L, utils.curry_normal_logpdf(loc_d, scale_d, ignore_const=False), [0.3])
    b = utils.batch_quantize_indep_dims((4, 3), P, L, utils.curry_normal_logpdf(loc_d, scale_d, ignore_const=True), [0.3])
    assert np.array_equal(a[0][0.3], b[0][0.3]) and np.array_equal(a[1][0.3], b[1][0.3])


def test_beta_sweep_equals_the_notebook_loop():
    """ipynb:464-473, 1102-1103: compressed_bitlength of every beta = empirical_entropy(compress_coordinates(...))."""
    from vbq_b200.word_embeddings import GaussianCodebook, empirical_entropy, empirical_std
    rng = np.random.default_rng(3)
    means = (rng.normal(size=(1237, 37)) * 1.2).astype(np.float32)
    stds = np.exp(rng.normal(size=means.shape) * 0.7 + np.log(0.05)).astype(np.float32)
    cb = GaussianCodebook(empirical_std(means), 10)
    betas = np.exp(np.linspace(np.log(0.01), np.log(100000), 7))
    for exact in (False, True):
        want = np.array([empirical_entropy(cb.compress_coordinates(means, stds, float(b), exact=exact)[0]) for b in betas])
        got = cb.beta_sweep(means, stds, betas, exact=exact)
        assert got.shape == (7,) and np.allclose(got, want, rtol=1e-12, atol=1e-6), (exact, got, want)
        small = cb.beta_sweep(means, stds, betas, exact=exact, max_chunk_symbols=7 * 4096)     # several row chunks
        assert np.allclose(small, want, rtol=1e-12, atol=1e-6)
    assert np.all(np.diff(cb.beta_sweep(means, stds, betas)) <= 1e-6)      # the rate falls as beta grows
    # row-sharded: the counts of two halves add up (what all_reduce_counts does across ranks)
    halves = []
    cb.beta_sweep(means[:600], stds[:600], betas, reduce_fn=lambda c: (halves.append(c.clone()), c)[1])
    got2 = cb.beta_sweep(means[600:], stds[600:], betas, reduce_fn=lambda c: c + halves[0])
    assert np.allclose(got2, cb.beta_sweep(means, stds, betas), rtol=1e-12)
