"""The TMA pipeline kernels (csrc/quantize_tma.cuh) against the literal reference walk (VBQ_FLAG_REFERENCE_WALK: both
bracket ends of every depth, IEEE float32 scores, first maximum in the reference's candidate order — quantizer.py:156-188,
utils.py:392-415) and against the cp.async bisection kernel, on inputs with near-ties in bulk: a third of the coordinates
sit exactly on code points, rows far outside the table, ragged row counts and channel counts, N < 10, early exit on/off.
Outputs must be identical; the totals identical between repeated runs and within 2e-7 of the other kernels' (the
distortion total is an exact sum of terms rounded to 2^-24)."""
import numpy as np
import pytest
import torch

import vbq_b200
from vbq_b200 import ops

pytestmark = pytest.mark.gpu
DT = {"zhat": torch.float32, "qidx": torch.int32, "level": torch.int32, "bits": torch.float32, "em_bits": torch.float32}


def _case(rows, C, N, seed):
    dev = torch.device("cuda", 0)
    pr = vbq_b200.BMSHJ2018Prior(C, device=dev, seed=seed)
    q = vbq_b200.ChannelwisePriorCDFQuantizer(C, N, device=dev)
    q.build_code_points(pr)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    u = torch.rand((rows, C), generator=g, device=dev, dtype=torch.float64) * 0.9998 + 0.0001
    m = pr.inverse_cdf(u).contiguous()
    tab = q.all_code_points
    idx = torch.randint(0, tab.shape[1], (rows, C), generator=g, device=dev)
    onpt = tab.t()[idx, torch.arange(C, device=dev)[None, :].expand(rows, C)]
    m = torch.where(torch.rand((rows, C), generator=g, device=dev) < 0.33, onpt, m).contiguous()
    if rows > 8:
        m[0] = 1e4
        m[1] = -1e4
        m[2] = tab[:, -1] * 1.0000001 + 1e-3          # just above the highest code point of the deepest level
        m[3] = tab[:, tab.shape[1] // 2]
    s = torch.exp(0.5 * (torch.randn((rows, C), generator=g, device=dev) * 1.5 - 3.0)).contiguous()
    return q, m, s


def _run(q, m, s, pen, length, em, flags, outs, N):
    L, (rows, C) = pen.shape[0], m.shape
    o = {k: torch.full((L, rows, C), -7, dtype=DT[k], device=m.device) for k in outs}
    tot = torch.zeros((L, 4), dtype=torch.float64, device=m.device)
    ops.quantize_into(m, s, q.all_code_points, q._packed, pen, length, em, N, totals=tot,
                      workspace=ops.quantize_workspace(L, m.device), flags=flags, **o)
    torch.cuda.synchronize()
    return o, tot


def _same(a, b):
    return all(torch.equal(a[k], b[k]) for k in a)


def _close(t, r, rel=2e-7):
    # terms are rounded to 2^-24 before the exact sum: relative for large totals, an absolute floor for tiny ones
    # (the sweep kernels sum float32 tiles of 64 terms in units of 2^-16: rel = 2e-6)
    return bool(((t - r).abs() <= rel * r.abs() + 1e-5 * (rel / 2e-7)).all())


@pytest.mark.parametrize("rows,C,N,lambs", [(1, 16, 10, [0.5]), (17, 20, 10, [0.5]), (1000, 36, 10, [0.01, 2.0]),
                                            (4099, 192, 10, [0.5, 8.0, 0.0]), (677, 48, 6, [0.3]), (333, 12, 0, [1.0]),
                                            (30000, 64, 10, [4.0])])
def test_raw_lengths_tma_equals_reference_walk(rows, C, N, lambs):
    q, m, s = _case(rows, C, N, rows + C)
    pen, length = q._length_tables(lambs)            # carries its host copy: the TMA kernel applies
    for outs in (("qidx", "bits"), ("zhat", "level"), ("zhat",), ("qidx",), ()):
        for fl in (0, ops.FLAG_NO_PRUNE):
            f = fl | ops.FLAG_NO_SWEEP
            ref, tr = _run(q, m, s, pen, length, None, f | ops.FLAG_REFERENCE_WALK, outs, N)
            old, to = _run(q, m, s, pen, length, None, f | ops.FLAG_NO_TMA, outs, N)
            new, tn = _run(q, m, s, pen, length, None, f, outs, N)
            new2, tn2 = _run(q, m, s, pen, length, None, f, outs, N)
            assert _same(ref, new) and _same(old, new), (outs, fl)
            assert torch.equal(tn, tn2) and _close(tn, tr) and _close(to, tr)
            assert torch.equal(tn[:, 0], tr[:, 0])                       # depth sums are integers: exact


def _lengths(style, rng, L, C, N):
    """Corrected code lengths n + R[c, n] of several shapes: R pure noise (every depth can have a cheaper deeper
    neighbour), like fitted tables (almost monotone, dips at a few shallow depths), monotone per channel, monotone with
    a dip at one deep depth in a few channels, and all equal (ties between depths everywhere)."""
    n = np.arange(N + 1, dtype=np.float32)[None, None, :]
    if style == "noisy":
        R = rng.gamma(2.0, 2.0, size=(L, C, N + 1))
    elif style == "fitted":
        R = 5.0 + 0.08 * (n - 3.0) ** 2 + rng.normal(0.0, 0.35, size=(L, C, N + 1))
    elif style == "monotone":
        return np.cumsum(rng.gamma(1.0, 1.0, size=(L, C, N + 1)), axis=2).astype(np.float32)
    elif style == "deep_dip":
        length = np.cumsum(rng.gamma(1.0, 1.0, size=(L, C, N + 1)) + 0.5, axis=2)
        dip = rng.random((L, C)) < 0.2
        length[..., 8] = np.where(dip, length[..., 6] - 0.25, length[..., 8])
        return length.astype(np.float32)
    else:
        return np.full((L, C, N + 1), 3.0, dtype=np.float32)
    return (n + R).astype(np.float32)


@pytest.mark.parametrize("style", ["noisy", "fitted", "monotone", "deep_dip", "flat"])
@pytest.mark.parametrize("rows,C,lambs", [(1, 16, [0.5]), (37, 20, [0.3]), (1000, 36, [0.01, 2.0]),
                                          (4099, 192, [0.5, 8.0, 0.0]), (20000, 64, [0.05])])
def test_arbitrary_penalties_tma_equals_reference_walk(rows, C, lambs, style):
    """Corrected code lengths n + R_lambda[c, n] and the entropy-model bits.  FLAG_NEIGHBOUR_EVERY_DEPTH makes the
    both-ends kernel score the neighbour at every depth; by default it does so only where the penalties of a channel
    group allow the neighbour to win (FLAG_NO_PRUNE, which the facade sets for small lambdas, must not change that)."""
    N = 10
    q, m, s = _case(rows, C, N, 7 * rows + C)
    rng = np.random.default_rng(rows)
    L = len(lambs)
    length = _lengths(style, rng, L, C, N)
    pen = ops.with_host_copy(np.asarray(lambs, dtype=np.float32)[:, None, None] * length, m.device)
    len_t = torch.from_numpy(length).to(m.device)
    em = torch.from_numpy(rng.gamma(2.0, 3.0, size=(L, C, 2 ** (N + 1) - 1)).astype(np.float32)).to(m.device)
    combos = ((("zhat", "bits", "em_bits"), em), (("zhat", "bits"), em), (("zhat", "bits"), None),
              (("qidx",), None), ((), em), ((), None))
    for outs, em_ in combos if style in ("noisy", "fitted") else combos[:1] + combos[3:4]:
        f = ops.FLAG_NO_SWEEP
        ref, tr = _run(q, m, s, pen, len_t, em_, f | ops.FLAG_REFERENCE_WALK, outs, N)
        for fl in (0, ops.FLAG_NEIGHBOUR_EVERY_DEPTH, ops.FLAG_NO_PRUNE):
            new, tn = _run(q, m, s, pen, len_t, em_, f | fl, outs, N)
            new2, tn2 = _run(q, m, s, pen, len_t, em_, f | fl, outs, N)
            assert _same(ref, new), (outs, fl)
            assert torch.equal(tn, tn2) and _close(tn, tr)
    # several lambdas in one call (no NO_SWEEP): the both-ends sweep kernel (csrc/sweep_both.cu), one walk for all lambdas
    for outs, em_ in ((("zhat", "bits", "em_bits"), em), (("zhat", "qidx", "level", "bits"), None), ((), em), ((), None)):
        ref, tr = _run(q, m, s, pen, len_t, em_, ops.FLAG_REFERENCE_WALK, outs, N)
        for fl in (0, ops.FLAG_NEIGHBOUR_EVERY_DEPTH):
            new, tn = _run(q, m, s, pen, len_t, em_, fl, outs, N)
            new2, tn2 = _run(q, m, s, pen, len_t, em_, fl, outs, N)
            assert _same(ref, new), (outs, fl)
            assert torch.equal(tn, tn2) and _close(tn, tr, 2e-6), (outs, fl, tn, tr)


@pytest.mark.parametrize("rows,C,N,L", [(700, 20, 6, 5), (3000, 48, 10, 40), (64, 16, 1, 2), (5000, 36, 9, 3)])
def test_both_ends_sweep_small_depths_and_lambda_chunks(rows, C, N, L):
    """csrc/sweep_both.cu with max_bits_per_coord < 10 (entropy-model bits gathered in the kernel: em_gather_kernel serves
    N = 10 only), with more lambdas than one launch holds (18), with entropy-model bits but no code-length output, and with
    code lengths beyond 512 bits."""
    q, m, s = _case(rows, C, N, rows + N)
    rng = np.random.default_rng(rows + L)
    lambs = [float(l) for l in 2.0 ** np.linspace(-8, 6, L)]
    length = _lengths("fitted" if N > 3 else "noisy", rng, L, C, N)
    if N == 9:
        length = (length * np.float32(200.0)).astype(np.float32)   # beyond 512 bits: the totals take the float32 butterfly
    pen = ops.with_host_copy(np.asarray(lambs, dtype=np.float32)[:, None, None] * length, m.device)
    len_t = torch.from_numpy(length).to(m.device)
    em = torch.from_numpy(rng.gamma(2.0, 3.0, size=(L, C, 2 ** (N + 1) - 1)).astype(np.float32)).to(m.device)
    for outs, em_ in ((("zhat", "bits", "em_bits"), em), (("zhat", "em_bits"), em), (("qidx", "level"), None), ((), em)):
        ref, tr = _run(q, m, s, pen, len_t, em_, ops.FLAG_REFERENCE_WALK, outs, N)
        new, tn = _run(q, m, s, pen, len_t, em_, 0, outs, N)
        new2, tn2 = _run(q, m, s, pen, len_t, em_, 0, outs, N)
        assert _same(ref, new), outs
        assert torch.equal(tn, tn2) and _close(tn, tr, 2e-6), (outs, tn, tr)


@pytest.mark.parametrize("lambs", [[0.5], [0.1, 0.5, 4.0]])
def test_entropy_model_bits_with_and_without_totals(lambs):
    """Every instantiation of em_gather_kernel in one process (with / without totals, after the single-lambda kernel and
    after the sweep): each needs its own shared-memory attribute."""
    N, rows, C = 10, 3000, 48
    q, m, s = _case(rows, C, N, 5)
    rng = np.random.default_rng(5)
    L = len(lambs)
    length = _lengths("fitted", rng, L, C, N)
    pen = ops.with_host_copy(np.asarray(lambs, dtype=np.float32)[:, None, None] * length, m.device)
    len_t = torch.from_numpy(length).to(m.device)
    em = torch.from_numpy(rng.gamma(2.0, 3.0, size=(L, C, 2 ** (N + 1) - 1)).astype(np.float32)).to(m.device)
    outs = ("zhat", "bits", "em_bits")
    ref, _ = _run(q, m, s, pen, len_t, em, ops.FLAG_REFERENCE_WALK, outs, N)
    for with_totals in (True, False, True, False):
        if with_totals:
            new, _ = _run(q, m, s, pen, len_t, em, 0, outs, N)
        else:
            new = {k: torch.full((L, rows, C), -7, dtype=DT[k], device=m.device) for k in outs}
            ops.quantize_into(m, s, q.all_code_points, q._packed, pen, len_t, em, N, **new)
            torch.cuda.synchronize()
        assert _same(ref, new), with_totals


@pytest.mark.parametrize("rows,C", [(300, 1024), (50000, 16), (129, 4080), (5, 2000)])
def test_work_distribution_variants(rows, C):
    """More channel groups than tile queues (static row ranges, CTAs that span several groups), one group for all CTAs,
    fewer tiles than CTAs; with and without totals (no workspace: static ranges)."""
    N, lambs = 10, [0.5]
    q, m, s = _case(rows, C, N, rows + C)
    pen, length = q._length_tables(lambs)
    ref, tr = _run(q, m, s, pen, length, None, ops.FLAG_NO_PRUNE | ops.FLAG_NO_TMA, ("qidx", "bits"), N)
    new, tn = _run(q, m, s, pen, length, None, ops.FLAG_NO_PRUNE, ("qidx", "bits"), N)
    assert _same(ref, new) and _close(tn, tr) and torch.equal(tn[:, :2], tr[:, :2])
    o = {k: torch.full((1, rows, C), -7, dtype=DT[k], device=m.device) for k in ("qidx", "bits")}
    ops.quantize_into(m, s, q.all_code_points, q._packed, pen, length, None, N, flags=ops.FLAG_NO_PRUNE, **o)
    torch.cuda.synchronize()
    assert _same(ref, o)
