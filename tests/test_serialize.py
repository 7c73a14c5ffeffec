"""Symbol serialisation: CPU tests of the oracle restatement and the ABI, GPU tests of the kernels."""
import numpy as np
import pytest
import torch

from oracle import vbq_oracle as O


def test_oracle_pack_known_answer_and_round_trip():
    # N = 2: 3-bit symbols 1,2,3,4,5,6,0,6,1,2,3 -> 33 bits -> 2 words, written out by hand
    q = [1, 2, 3, 4, 5, 6, 0, 6, 1, 2, 3]
    bits = "".join(format(v, "03b")[::-1] for v in q)            # LSB first
    bits += "0" * (64 - len(bits))
    want = [int(bits[32 * k:32 * k + 32][::-1], 2) for k in range(2)]
    got = O.pack_indices(q, 2)
    assert got.dtype == np.uint32 and got.tolist() == want
    assert O.unpack_indices(got, len(q), 2).tolist() == q
    rng = np.random.default_rng(0)
    for N in (0, 1, 4, 10, 16, 20):
        for n in (0, 1, 31, 32, 33, 1000):
            s = rng.integers(0, 2 ** (N + 1) - 1, n, endpoint=True if N == 0 else False).astype(np.int32)
            s = np.minimum(s, 2 ** (N + 1) - 2)
            w = O.pack_indices(s, N)
            assert w.size == (n * (N + 1) + 31) // 32
            assert np.array_equal(O.unpack_indices(w, n, N), s)


def test_abi_sizes_need_no_gpu():
    from vbq_b200 import _lib
    lib = _lib.load()
    assert lib.vbq_packed_index_words(7077888, 10) == 7077888 * 11 // 32
    assert lib.vbq_packed_index_words(33, 2) == 4 and lib.vbq_packed_index_words(0, 10) == 0
    assert lib.vbq_packed_index_words(-1, 10) == -1 and lib.vbq_packed_index_words(5, 21) == -1
    assert lib.vbq_pack_indices(None, -1, 10, None, None) == 2
    assert lib.vbq_symbol_histogram(None, 4, 0, 10, None, None) == 2


@pytest.mark.gpu
@pytest.mark.parametrize("N", [0, 1, 4, 10, 13, 20])
@pytest.mark.parametrize("n", [1, 31, 32, 33, 4097, 300001])
def test_pack_unpack_match_oracle(N, n):
    from vbq_b200 import ops
    rng = np.random.default_rng(N * 1000 + n)
    s = rng.integers(0, 2 ** (N + 1) - 1, n).astype(np.int32) if N > 0 else np.zeros(n, np.int32)
    d = torch.from_numpy(s).cuda()
    w = ops.pack_indices(d, N)
    assert np.array_equal(w.cpu().numpy().view(np.uint32), O.pack_indices(s, N))      # bit-exact stream
    assert np.array_equal(ops.unpack_indices(w, n, N).cpu().numpy(), s)


@pytest.mark.gpu
def test_histogram_and_stream_of_kernel_output():
    """Quantize a batch, serialise its indices with frequency tables, read them back: identical symbols, counts equal
    to the reference's per-channel bincount, and the ideal code length is below the fixed-width size."""
    import vbq_b200
    from vbq_b200 import ops, serialize
    import vbq_test_helpers as H
    C, N, rows = 37, 10, 5000
    pr = H.make_prior(C, seed=4)
    q = vbq_b200.ChannelwisePriorCDFQuantizer(C, N)
    q.set_code_points(ops.build_code_points_learned(torch.from_numpy(pr.packed()).cuda(), N))
    mu, sigma, _ = H.make_latents(pr, rows, 5)
    lambs = [2.0 ** -6, 0.5, 8.0]
    out = q.quantize(torch.from_numpy(mu).cuda(), torch.from_numpy(sigma).cuda(), lambs, outputs=ops.OUT_QIDX)
    qi = out["qidx"]
    Q = q.quantization_levels
    for i in range(len(lambs)):
        counts = ops.symbol_histogram(qi[i], N)
        assert np.array_equal(counts.cpu().numpy(), O.symbol_histogram(qi[i].cpu().numpy(), Q))
    blob = serialize.dumps(qi, N, with_counts=True)
    back = serialize.loads(blob)
    assert back["max_bits"] == N and torch.equal(back["qidx"], qi)
    assert np.array_equal(back["counts"][1].numpy(), O.symbol_histogram(qi[1].cpu().numpy(), Q))
    n_words = ops.packed_index_words(rows * C, N)
    assert len(blob) == 28 + 3 * C * Q * 4 + 3 * n_words * 4
    ideal = serialize.ideal_code_length_bits(back["counts"][1])
    assert 0 < ideal < rows * C * (N + 1)
    # accumulation over row shards == one pass (what a sharded fit all-reduces)
    acc = torch.zeros((C, Q), dtype=torch.int64, device="cuda")
    for a, b in ((0, 1234), (1234, 5000)):
        ops.symbol_histogram(qi[0][a:b].contiguous(), N, acc)
    assert torch.equal(acc, ops.symbol_histogram(qi[0], N))
