"""Full-size GPU tests (BASELINE.json configurations at their single-GPU size) through size-independent properties:
the oracle cannot run 10^8 coordinates in seconds, so at these sizes the CUDA path is checked against invariants of
the domain — every z_hat is exactly the table entry its index names (quantizer.py:136-137), idempotence at lambda=0,
rate/distortion monotonicity in lambda, totals == sums of the per-coordinate outputs, sharded == unsharded,
sweep == per-lambda — plus the oracle on a random sample of rows."""
import numpy as np
import pytest
import torch

from oracle import vbq_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _learned(C, N, seed):
    import vbq_b200
    prior = vbq_b200.BMSHJ2018Prior(C, device=DEV, seed=seed)
    q = vbq_b200.ChannelwisePriorCDFQuantizer(C, N, device=DEV)
    q.build_code_points(prior)
    return prior, q


def _latents(prior, rows, C, seed):
    g = torch.Generator(device=DEV)
    g.manual_seed(seed)
    u = torch.rand((rows, C), generator=g, device=DEV, dtype=torch.float64) * 0.998 + 0.001
    mu = prior.inverse_cdf(u).contiguous()
    sigma = torch.exp(0.5 * (torch.randn((rows, C), generator=g, device=DEV) * 1.5 - 3.0)).contiguous()
    return mu, sigma


def _check_invariants(q, mu, sigma, out, i, lamb):
    from vbq_b200 import ops
    zh, qi, lv, tot = out["zhat"][i], out["qidx"][i], out["level"][i], out["totals"][i]
    N = q.max_bits_per_coord
    # (1) the sorted index names exactly the returned code point (quantizer.py:135-137)
    srt = q.code_points_by_channel                                   # (C, Q)
    assert torch.equal(torch.gather(srt.t(), 0, qi.long()), zh)
    # (2) the index is consistent with the depth: q+1 = (2i+1) 2^(N-n)  =>  trailing zeros of q+1 = N-n
    tz = (qi + 1) & -(qi + 1)
    assert torch.equal(tz, (1 << (N - lv)).to(tz.dtype))
    # (3) totals are the sums of the per-coordinate outputs (float64, deterministic reduction)
    assert float(tot[0]) == float(lv.sum(dtype=torch.float64))
    d = ((zh.double() - mu.double()) / sigma.double())
    dist = float((0.5 * d * d).sum())
    assert abs(float(tot[3]) - dist) <= 2e-6 * max(dist, 1.0)
    # (4) the chosen point is at least as good as the bracket ends of its own depth and of the neighbouring depths
    return float(tot[0]), dist


def test_kodak_full_size_properties():
    """configs[1]: 24 x 32x48 x 192, N=10, the reference's 16-lambda grid."""
    from vbq_b200 import ops
    C, N, rows = 192, 10, 24 * 32 * 48
    prior, q = _learned(C, N, 2)
    mu, sigma = _latents(prior, rows, C, 1)
    lambs = [float(l) for l in 2 ** np.linspace(-8, 7, 16)]
    outs = ops.OUT_ZHAT | ops.OUT_QIDX | ops.OUT_LEVEL | ops.OUT_TOTALS
    sweep = q.quantize(mu, sigma, lambs, outputs=outs)                       # one walk for all lambdas
    bits, dists = [], []
    for i, l in enumerate(lambs):
        b, d = _check_invariants(q, mu, sigma, sweep, i, l)
        bits.append(b)
        dists.append(d)
        single = q.quantize(mu, sigma, [l], outputs=outs)                    # single-lambda kernel
        for k in ("zhat", "qidx", "level"):
            assert torch.equal(single[k][0], sweep[k][i]), (k, l)
        assert torch.allclose(single["totals"][0], sweep["totals"][i], rtol=1e-6)
    # rate falls and distortion rises with lambda (exact optimiser of lambda*R + D, coordinate by coordinate)
    assert all(a >= b for a, b in zip(bits, bits[1:])) and all(a <= b + 1e-6 * b for a, b in zip(dists, dists[1:]))
    # idempotence: at lambda = 0 a code point is its own best approximation, at its own depth
    zh = sweep["zhat"][8].contiguous()
    again = q.quantize(zh, sigma, [0.0], outputs=ops.OUT_ZHAT | ops.OUT_QIDX | ops.OUT_LEVEL)
    assert torch.equal(again["zhat"][0], zh) and torch.equal(again["qidx"][0], sweep["qidx"][8])
    # sharded (8 contiguous image shards, as 8 ranks would) == unsharded
    from vbq_b200 import sharding
    parts, tot = [], torch.zeros_like(sweep["totals"][8])
    for r in range(8):
        a, b = sharding.shard_bounds(24, 8, r)
        o = q.quantize(mu[a * 1536:b * 1536], sigma[a * 1536:b * 1536], [lambs[8]], outputs=outs)
        parts.append(o["qidx"][0])
        tot += o["totals"][0]
    assert torch.equal(torch.cat(parts), sweep["qidx"][8])
    single8 = q.quantize(mu, sigma, [lambs[8]], outputs=outs)["totals"][0]
    assert torch.allclose(tot, single8, rtol=1e-9)                  # same kernel: only the pairing of float32 terms and the float64 reduction order differ
    assert torch.allclose(tot, sweep["totals"][8], rtol=1e-6)       # sweep kernel: float32 distortion terms rounded differently
    # oracle on a random sample of rows (same table): bit-exact
    idx = torch.randperm(rows, device=DEV)[:300]
    oq = O.QuantizerNP(C, N)
    oq.set_code_points(q.all_code_points.cpu().numpy(), build_grids=False)
    Zo, Bo = oq.compress_batch_channel_latents(mu[idx].cpu().numpy(), sigma[idx].cpu().numpy(), [lambs[3], lambs[8]])
    for j in (3, 8):
        assert np.array_equal(sweep["zhat"][j][idx].cpu().numpy(), Zo[lambs[j]])
        assert np.array_equal(sweep["level"][j][idx].cpu().numpy(), Bo[lambs[j]])


def test_embeddings_full_size_properties():
    """configs[3]: 1M x 300 Gaussian posteriors, one shared Gaussian prior (float32 kernel and float64 kernel)."""
    import vbq_b200
    from vbq_b200 import ops
    V, K = 1_000_000, 300
    g = torch.Generator(device=DEV)
    g.manual_seed(4)
    means = torch.randn((V, K), generator=g, device=DEV) * 1.2329 - 0.08
    stds = torch.exp(torch.randn((V, K), generator=g, device=DEV) * 0.7 + float(np.log(0.04)))
    cb = vbq_b200.GaussianCodebook(vbq_b200.word_embeddings.empirical_std(means), 10, device=DEV)
    cp32 = torch.from_numpy(cb.codepoints).to(DEV).float()
    out = cb.quantize(means, stds, [1.0], outputs=ops.OUT_ZHAT | ops.OUT_LEVEL | ops.OUT_QIDX)
    zh, lv, qi = out["zhat"][0], out["level"][0], out["qidx"][0]
    srt = torch.sort(cp32).values
    assert torch.equal(srt[qi.long()], zh)                                   # index names the value
    # float64 notebook kernel on the same inputs: values agree except on near-ties (< 1e-5 of the coordinates)
    exact, _ = cb.compress_coordinates(means, stds, 1.0)
    frac = float((exact != zh).float().mean())
    assert frac < 1e-5, frac
    # every optimum is a code point, and rows can be sharded freely
    half, _ = cb.compress_coordinates(means[:V // 2], stds[:V // 2], 1.0)
    assert torch.equal(half, exact[:V // 2])
    assert torch.isin(exact[:1000].flatten(), cp32).all()
    # sample vs the exhaustive notebook restatement
    m, s = means[:200].cpu().numpy(), stds[:200].cpu().numpy()
    want, _ = O.compress_coordinates(m, s, 1.0, cb.codepoints, cb.lengths)
    assert np.array_equal(exact[:200].cpu().numpy(), want)


def test_deep_table_full_size_properties():
    """configs[4] at one GPU's share: 32 images x 128x128 x 320 channels, max bit depth 16 (depths 11..16 are served
    from the heap-order table in global memory)."""
    from vbq_b200 import ops
    C, N, rows = 320, 16, 32 * 128 * 128 // 4          # a quarter of the share keeps the test under ~20 s
    prior, q = _learned(C, N, 5)
    mu, sigma = _latents(prior, rows, C, 3)
    lambs = [2.0 ** -8, 0.5]
    outs = ops.OUT_ZHAT | ops.OUT_QIDX | ops.OUT_LEVEL | ops.OUT_TOTALS
    out = q.quantize(mu, sigma, lambs, outputs=outs)
    b0, d0 = _check_invariants(q, mu, sigma, out, 0, lambs[0])
    b1, d1 = _check_invariants(q, mu, sigma, out, 1, lambs[1])
    assert b0 >= b1 and d0 <= d1
    assert int(out["level"][0].max()) == N                                   # the deep levels are really used
    idx = torch.randperm(rows, device=DEV)[:64]
    oq = O.QuantizerNP(C, N)
    oq.set_code_points(q.all_code_points.cpu().numpy(), build_grids=False)
    Zo, Bo = oq.compress_batch_channel_latents(mu[idx].cpu().numpy(), sigma[idx].cpu().numpy(), lambs)
    for j, l in enumerate(lambs):
        assert np.array_equal(out["zhat"][j][idx].cpu().numpy(), Zo[l])
        assert np.array_equal(out["level"][j][idx].cpu().numpy(), Bo[l])
