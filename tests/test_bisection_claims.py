"""CPU checks of the two claims the certified-bisection kernels rest on, against the oracle:

1. with non-decreasing penalties the reference's first argmax over the 2N+1 bracket candidates is always a PATH NODE of
   the tree walk (so one candidate per depth suffices);
2. the approximate loss A = fma(t, t, pen), t = (z - mu) * (sqrt(1/2) * rcp(sigma)), differs from the reference's
   float32 loss E = fl(0.5 * fl(fl(fl(z - mu) / sigma)^2) + pen) by less than 32 units of the float32 bit pattern, for
   a reciprocal that is off by up to one ulp either way (MUFU.RCP), which is what the guard band of 192 assumes;
plus the share of coordinates the certificate rejects on bench-shaped inputs (they take the literal search)."""
import numpy as np

from oracle import vbq_oracle as O
import vbq_test_helpers as H

F32 = np.float32


def path_nodes(table_heap, mu):
    """(C, Q) heap-order table, mu (B, C) -> (N+1, B, C): the node the walk `mu > z ? right : left` visits per depth."""
    C, Q = table_heap.shape
    N = int(np.log2(Q + 1)) - 1
    B = mu.shape[0]
    cidx = np.broadcast_to(np.arange(C)[None, :], (B, C))
    ip = np.zeros((B, C), dtype=np.int64)
    out = np.empty((N + 1, B, C), dtype=F32)
    for n in range(N + 1):
        z = table_heap[cidx, (1 << n) - 1 + ip]
        out[n] = z
        ip = 2 * ip + (mu > z)
    return out


def _setup(C, N, rows, seed):
    pr = H.make_prior(C, seed=seed)
    xi = O.xi_heap(N)
    table = pr.inverse_cdf_f64(np.repeat(xi[:, None], C, axis=1)).astype(F32).T.copy()   # (C, Q) heap order
    mu, sigma, _ = H.make_latents(pr, rows, seed + 1, table=table)
    srt = np.sort(table, axis=1)
    rng = np.random.default_rng(seed + 2)
    idx = rng.integers(0, srt.shape[1], (rows, C))
    on = srt[np.arange(C)[None, :], idx]
    nxt = srt[np.arange(C)[None, :], np.minimum(idx + 1, srt.shape[1] - 1)]
    pick = rng.integers(0, 6, (rows, C))
    mu = np.where(pick == 0, on, np.where(pick == 1, (0.5 * (on.astype(np.float64) + nxt)).astype(F32), mu)).astype(F32)
    oq = O.QuantizerNP(C, N)
    oq.set_code_points(table, build_grids=False)
    return table, mu, sigma, oq


def test_reference_winner_is_always_a_path_node():
    for (C, N, rows, seed) in ((24, 10, 3000, 1), (7, 6, 2000, 2), (3, 1, 500, 3)):
        table, mu, sigma, oq = _setup(C, N, rows, seed)
        lambs = [0.0, 1e-30, 2.0 ** -8, 0.1, 0.5, 3.0, 16.0, 1e4]
        Zo, Bo = oq.compress_batch_channel_latents(mu, sigma, lambs)
        nodes = path_nodes(table, mu)
        for l in lambs:
            at_depth = np.take_along_axis(nodes, Bo[l][None].astype(np.int64), axis=0)[0]
            assert np.array_equal(at_depth, Zo[l]), "lambda=%g: the winner is not the path node of its depth" % l


def test_non_monotone_penalties_break_the_claim():
    """The monotonicity condition is necessary: with corrected (non-monotone) lengths some winners are neighbours."""
    table, mu, sigma, oq = _setup(24, 10, 3000, 5)
    rng = np.random.default_rng(0)
    oq.raw_code_length_entropy_models = {0.5: rng.uniform(0.25, 6.0, (24, 11)).astype(F32)}
    Zo, Bo, det = oq.compress_batch_channel_latents(mu, sigma, [0.5], details=True)
    nodes = path_nodes(table, mu)
    lvl = det[0.5]["level"]
    at_depth = np.take_along_axis(nodes, lvl[None].astype(np.int64), axis=0)[0]
    assert (at_depth != Zo[0.5]).any()


def test_neighbour_wins_only_below_the_running_maximum_of_the_penalties():
    """DESIGN.md §2c (round 2): with arbitrary per-channel penalties the reference's winner is a path node, or the in-level
    neighbour at a depth n whose penalty is below max(pen[0..n-1]) — elsewhere the neighbour loses to the ancestor it lies
    beyond, which has a penalty no larger, is at least as near and comes first in the candidate order.  Penalties of
    several shapes, ties included (repeated values, lambda = 0)."""
    table, mu, sigma, oq = _setup(24, 10, 3000, 9)
    nodes = path_nodes(table, mu)
    rng = np.random.default_rng(3)
    n_ = np.arange(11, dtype=F32)[None, :]
    tables = {"noise": rng.uniform(0.25, 6.0, (24, 11)),
              "fitted": 5.0 + 0.08 * (n_ - 3.0) ** 2 + rng.normal(0.0, 0.35, (24, 11)) ,
              "steps": np.repeat(rng.uniform(0.0, 4.0, (24, 4)), 3, axis=1)[:, :11] - n_,      # plateaus: equal lengths
              "dip": np.where((n_ == 8) & (rng.random((24, 1)) < 0.5), -3.0, 0.5 * n_)}
    neighbours = 0
    for name, R in tables.items():
        for lam in (0.0, 2.0 ** -8, 0.5, 8.0):
            oq.raw_code_length_entropy_models = {lam: R.astype(F32)}
            Zo, Bo, det = oq.compress_batch_channel_latents(mu, sigma, [lam], details=True)
            lvl = det[lam]["level"].astype(np.int64)
            pen = (F32(lam) * oq.code_lengths([lam])[0].astype(F32)).astype(F32)          # (N+1, C), as the reference scores
            below = np.zeros_like(pen, dtype=bool)
            below[1:] = pen[1:] < np.maximum.accumulate(pen, axis=0)[:-1]
            is_path = np.take_along_axis(nodes, lvl[None], axis=0)[0] == Zo[lam]
            allowed = np.take_along_axis(below, lvl, axis=0)                            # (B, C): winner's depth is flagged
            assert (is_path | allowed).all(), "%s, lambda=%g: a neighbour won at a depth where it should be dominated" % (name, lam)
            neighbours += int((~is_path).sum())
    assert neighbours > 0


def _keys(z, mu, sigma, pen, rcp_ulps):
    """Bit patterns of the exact loss E and of the approximate loss A (float64 emulation of the single-rounding FMA)."""
    d = (z - mu).astype(F32)
    t = (d / sigma).astype(F32)
    E = (F32(0.5) * (t * t).astype(F32) + pen).astype(F32)
    r = (F32(1.0) / sigma).astype(F32)
    r = (r.view(np.int32) + rcp_ulps).view(F32)                       # reciprocal off by up to one ulp
    r2 = (r * F32(0.70710678)).astype(F32)
    ta = (d * r2).astype(F32)
    A = (ta.astype(np.float64) * ta.astype(np.float64) + pen.astype(np.float64)).astype(F32)
    return E.view(np.int32).astype(np.int64), A.view(np.int32).astype(np.int64)


def test_key_error_bound():
    rng = np.random.default_rng(7)
    n = 2_000_000
    worst = 0
    for scale_mu, scale_sig, lam in ((10.0, 1.0, 0.5), (100.0, 0.01, 2.0 ** -8), (1.0, 30.0, 16.0), (1e-3, 1e-5, 0.0),
                                     (50.0, 1e-3, 1e4), (1e-20, 1e-18, 1e-30)):
        z = (rng.standard_normal(n) * scale_mu).astype(F32)
        mu = (z + rng.standard_normal(n).astype(F32) * F32(scale_sig) * rng.choice([1e-3, 1.0, 30.0], n).astype(F32)).astype(F32)
        sigma = np.exp(rng.normal(np.log(scale_sig), 1.5, n)).astype(F32)
        pen = (F32(lam) * rng.integers(0, 11, n).astype(F32)).astype(F32)
        for ulps in (-1, 0, 1):
            E, A = _keys(z, mu, sigma, pen, ulps)
            ok = np.isfinite(E.astype(np.int32).view(F32)) & np.isfinite(A.astype(np.int32).view(F32))
            worst = max(worst, int(np.abs(E - A)[ok].max()))
    print("largest |bits(A) - bits(E)| over 36 M samples: %d" % worst)
    assert worst < 32


def test_share_of_uncertified_coordinates_on_bench_shaped_inputs():
    """Kodak-shaped synthetic latents, lambda = 0.5: how many coordinates have their two best keys within
    the guard of 192 (they are redone by the literal search on the GPU)."""
    C, N, rows = 48, 10, 4000
    pr = H.make_prior(C, seed=11, factor_std=0.0)
    xi = O.xi_heap(N)
    table = pr.inverse_cdf_f64(np.repeat(xi[:, None], C, axis=1)).astype(F32).T.copy()
    mu, sigma, _ = H.make_latents(pr, rows, 12, edge_cases=False)
    nodes = path_nodes(table, mu)                                     # (N+1, B, C)
    shares = {}
    for lam in (2.0 ** -8, 0.5, 8.0):
        pen = (F32(lam) * np.arange(N + 1, dtype=F32))[:, None, None] * np.ones_like(nodes)
        _, A = _keys(nodes, mu[None], sigma[None], pen.astype(F32), 0)
        keys = (A & ~np.int64(15)) | np.arange(N + 1, dtype=np.int64)[:, None, None]
        two = np.sort(keys, axis=0)[:2]
        shares[lam] = float(((two[1] - two[0] - 1) <= 192).mean())
    print("uncertified share:", shares)
    assert all(v < 2e-3 for v in shares.values())


def test_certified_winner_equals_reference_winner():
    """End-to-end model of the kernel's decision in NumPy: integer keys of the approximate losses of the path nodes,
    winner = minimum, certificate = gap to the runner-up > 192.  Every certified coordinate must carry the oracle's
    depth and code point — on inputs where a third of the coordinates sit on code points or exactly between two — and
    the reciprocal may be off by an ulp either way."""
    total = cert = 0
    for (C, N, rows, seed) in ((24, 10, 4000, 21), (9, 7, 3000, 22)):
        table, mu, sigma, oq = _setup(C, N, rows, seed)
        nodes = path_nodes(table, mu)
        lambs = [0.0, 2.0 ** -8, 0.1, 0.5, 3.0, 16.0]
        Zo, Bo = oq.compress_batch_channel_latents(mu, sigma, lambs)
        for l in lambs:
            pen = ((F32(l) * np.arange(N + 1, dtype=F32))[:, None, None] * np.ones_like(nodes)).astype(F32)
            for ulps in (-1, 0, 1):
                _, A = _keys(nodes, mu[None], sigma[None], pen, ulps)
                keys = (A & ~np.int64(15)) | np.arange(N + 1, dtype=np.int64)[:, None, None]
                order = np.sort(keys, axis=0)
                winner = (order[0] & 15).astype(np.int64)
                certified = (order[1] - order[0] - 1 > 192) if N > 0 else np.ones_like(winner, dtype=bool)
                certified &= np.isfinite(A.astype(np.int32).view(F32)).all(axis=0)
                z_w = np.take_along_axis(nodes, winner[None], axis=0)[0]
                assert np.array_equal(winner[certified], Bo[l][certified].astype(np.int64)), (l, ulps)
                assert np.array_equal(z_w[certified], Zo[l][certified]), (l, ulps)
                total += certified.size
                cert += int(certified.sum())
    print("certified %d of %d (coordinate, lambda, rcp error) cases; the rest take the literal search" % (cert, total))
    assert cert > 0.6 * total
