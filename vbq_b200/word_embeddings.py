"""Word-embedding VBQ: the notebook cells of the reference's
word-embeddings/compress-trained-word-embeddings.ipynb:373-390 (code points) and :429-443
(`compress_coordinates`), on the GPU.

All coordinates share one zero-mean Gaussian prior N(0, empirical_std^2).  The shared code-point tree is
replicated over 16 virtual channels so the same sm_100a kernel as the image path walks it
(bank-conflict-free interleaving, include/vbq_b200.h)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops

_VC = 16  # virtual channels = VBQ_GROUP


def empirical_std(means):
    """ipynb:373-374: sqrt(mean(mu^2))."""
    m = means if isinstance(means, torch.Tensor) else torch.as_tensor(np.asarray(means))
    return float(torch.sqrt(torch.mean(m.double().reshape(-1) ** 2)))


def empirical_entropy(values):
    """ipynb:452-455: total_counts*log2(total_counts) - sum(counts*log2(counts)) over the distinct values, i.e. the
    bit length of an ideal entropy code of the quantized coordinates.  Device tensors stay on the device."""
    t = values if isinstance(values, torch.Tensor) else torch.as_tensor(np.asarray(values))
    _, counts = torch.unique(t.reshape(-1), return_counts=True)
    c = counts.double()
    total = c.sum()
    return float(total * torch.log2(total) - (c * torch.log2(c)).sum())


class GaussianCodebook:
    """`codepoints` / `lengths` of the notebook (heap order, float64 / int64) plus the device tables."""

    def __init__(self, empirical_std, max_codepoint_length=10, device="cuda"):
        self.max_codepoint_length = int(max_codepoint_length)
        self.empirical_std = float(empirical_std)
        self.device = torch.device(device)
        N = self.max_codepoint_length
        # codepoint_xi = np.arange(0.5**(length+1), 1, 0.5**length), length-major (ipynb:383-388)
        xi = np.concatenate([np.arange(0.5 ** (n + 1), 1, 0.5 ** n) for n in range(N + 1)])
        xi_d = torch.from_numpy(xi).to(self.device).reshape(-1, 1)
        std_d = torch.tensor([self.empirical_std], dtype=torch.float64, device=self.device)
        pts = ops.gaussian_inverse_cdf(xi_d, None, std_d).reshape(-1)          # norm.ppf(xi, scale=std), float64
        self.codepoints = pts.cpu().numpy()
        self.lengths = np.concatenate([np.full(2 ** n, n, dtype=np.int64) for n in range(N + 1)])
        self._table = pts.to(torch.float32).reshape(1, -1).repeat(_VC, 1).contiguous()   # (16, Q)
        self._packed = ops.pack_code_points(self._table, N)
        torch.cuda.current_stream(self._table.device).synchronize()
        ops.stable_packed(self._packed)

    def _level_lengths(self, bitlengths):
        N = self.max_codepoint_length
        bl = np.asarray(bitlengths)
        per_level = np.array([bl[2 ** n - 1] for n in range(N + 1)], dtype=np.float64)
        if not np.array_equal(np.repeat(per_level, [2 ** n for n in range(N + 1)]), bl.astype(np.float64)):
            raise NotImplementedError("bitlengths must be constant within a bit depth for the bracketing search")
        return per_level

    def quantize(self, means, stds, betas, bitlengths=None, outputs=ops.OUT_ZHAT, flags=0):
        """means, stds: float32 CUDA tensors of equal shape -> dict of (len(betas),) + means.shape tensors."""
        N = self.max_codepoint_length
        per_level = self._level_lengths(self.lengths if bitlengths is None else bitlengths).astype(np.float32)
        pen = np.stack([np.float32(b) * per_level for b in betas])[:, None, :]
        pen = ops.with_host_copy(pen, self.device)   # channel-independent: launch constants of the TMA kernel
        # the notebook's default lengths are the bit depths themselves: no length table, which lets vbq_quantize use
        # the certified-bisection kernels (raw code lengths); custom bitlengths go through the length-table path
        length = None
        if bitlengths is not None and not np.array_equal(per_level, np.arange(N + 1, dtype=np.float32)):
            length = torch.from_numpy(np.broadcast_to(per_level, (len(betas), 1, N + 1)).copy()).to(self.device)
        m = means.reshape(-1)
        s = stds.reshape(-1)
        n = m.numel()
        pad = (-n) % _VC
        if pad:
            m = torch.cat([m, m.new_zeros(pad)])
            s = torch.cat([s, s.new_ones(pad)])
        z, q, lv, b, _, tot = ops.quantize(m.reshape(-1, _VC).contiguous(), s.reshape(-1, _VC).contiguous(),
                                           self._table, self._packed, pen, length, None, N, outputs,
                                           ops.search_flags(betas, flags))
        L = len(betas)

        def unpad(t):
            return t.reshape(L, -1)[:, :n].reshape((L,) + tuple(means.shape)) if t.numel() else t

        return dict(zhat=unpad(z), qidx=unpad(q), level=unpad(lv), bits=unpad(b), totals=tot)

    def beta_sweep(self, means, stds, betas, exact=False, reduce_fn=None, max_chunk_symbols=1 << 29):
        """The rate side of the notebook's beta sweep (ipynb:464-473, :1102-1103: `test_beta` for every beta of
        `np.exp(np.linspace(np.log(0.01), np.log(100000), 50))`): compresses all coordinates with every beta and
        returns the `compressed_bitlength` = `empirical_entropy(compressed)` (ipynb:452-455) of each, as a float64
        array (the embedding-quality metrics of `test_beta` need the skip-gram model and are out of scope).

        The coordinates are walked ONCE for all betas (sweep kernel) in row chunks of at most ``max_chunk_symbols``
        (beta, coordinate) pairs; the symbols of a chunk are counted on the device (`vbq_symbol_histogram`) and only the
        (len(betas), Q) count table leaves the loop.  ``exact=True`` runs the notebook's float64 search once per beta
        instead.  ``reduce_fn`` (optional) all-reduces the int64 counts over row-sharded ranks
        (`vbq_b200.sharding.all_reduce_counts`) before the entropies are formed."""
        N, Q = self.max_codepoint_length, 2 ** (self.max_codepoint_length + 1) - 1
        m = (means if isinstance(means, torch.Tensor) else torch.as_tensor(np.asarray(means)))
        s = (stds if isinstance(stds, torch.Tensor) else torch.as_tensor(np.asarray(stds)))
        m = m.to(device=self.device, dtype=torch.float32).reshape(-1)
        s = s.to(device=self.device, dtype=torch.float32).reshape(-1)
        betas = [float(b) for b in betas]
        L, n = len(betas), m.numel()
        counts = torch.zeros((L, _VC, Q), dtype=torch.int64, device=self.device)
        if exact:
            cp = torch.from_numpy(self.codepoints).to(self.device)
            per_level = torch.from_numpy(self._level_lengths(self.lengths).astype(np.float64)).to(self.device)
            step = max(_VC, (max_chunk_symbols // _VC) * _VC)
            for i, beta in enumerate(betas):
                for a in range(0, n, step):
                    b = min(n, a + step)
                    _, heap, level = ops.compress_coordinates_f64(m[a:b], s[a:b], cp, per_level, beta, True,
                                                                  want_index=True, want_level=True)
                    # heap index h of depth d -> sorted index (2 i + 1) 2^(N-d) - 1 with i = h - (2^d - 1)
                    i_in = heap - ((1 << level) - 1)
                    q = ((2 * i_in + 1) << (N - level)) - 1
                    flat = torch.bincount(q.reshape(-1).long(), minlength=Q)
                    counts[i, 0] += flat
        else:
            step = max(_VC, (max_chunk_symbols // max(L, 1) // _VC) * _VC)
            for a in range(0, n, step):
                b = min(n, a + step)
                k = ((b - a) // _VC) * _VC
                q = self.quantize(m[a:b], s[a:b], betas, outputs=ops.OUT_QIDX)['qidx']        # (L, b - a)
                for i in range(L):
                    if k:
                        ops.symbol_histogram(q[i, :k].reshape(-1, _VC).contiguous(), N, counts[i])
                    if k < b - a:     # the ragged tail (fewer than 16 symbols)
                        counts[i, 0] += torch.bincount(q[i, k:].long(), minlength=Q)
        counts = counts.sum(dim=1)
        if reduce_fn is not None:
            counts = reduce_fn(counts)
        c = counts.double()
        total = c.sum(dim=1)
        plogp = torch.where(c > 0, c * torch.log2(c.clamp_min(1.0)), torch.zeros_like(c)).sum(dim=1)
        return (total * torch.log2(total) - plogp).cpu().numpy()

    def compress_coordinates(self, means, stds, beta, bitlengths=None, exact=True):
        """Notebook signature (ipynb:429-443): returns (optima shaped and typed like `means`, None).
        Minimises (c - mu)^2 + 2 beta sigma^2 len(c) over the code points.

        ``exact=True`` (default) runs the float64 kernel that reproduces the notebook's arithmetic bit for bit
        (float64 code points and losses; `(2*beta)*stds**2` in float32 unless `beta` is a NumPy float64 scalar, which
        NumPy >= 2 promotes to float64).  ``exact=False`` runs the float32 image-path kernel (lambda = beta), which
        is faster and differs only on near-ties."""
        is_np = not isinstance(means, torch.Tensor)
        m = torch.as_tensor(np.asarray(means)) if is_np else means
        s = torch.as_tensor(np.asarray(stds)) if is_np else stds
        m32 = m.to(device=self.device, dtype=torch.float32).contiguous()
        s32 = s.to(device=self.device, dtype=torch.float32).contiguous()
        if exact:
            per_level = self._level_lengths(self.lengths if bitlengths is None else bitlengths)
            pen_f32 = not (isinstance(beta, np.floating) and np.lib.NumpyVersion(np.__version__) >= "2.0.0")
            out, _, _ = ops.compress_coordinates_f64(
                m32, s32, torch.from_numpy(self.codepoints).to(self.device),
                torch.from_numpy(per_level.astype(np.float64)).to(self.device), float(beta), pen_f32)
        else:
            out = self.quantize(m32, s32, [beta], bitlengths)['zhat'][0]
        if is_np:
            return out.cpu().numpy().astype(np.asarray(means).dtype), None
        return out.to(means.dtype), None
