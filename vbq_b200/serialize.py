"""Wire format for the emitted symbols (arithmetic coding itself stays out of scope).

The reference stops at counting its sorted quantile indices per channel (quantizer.py:135-146) and reporting ideal
code lengths (ipynb:452-455).  This module turns the kernel's `qidx` output into bytes an external entropy coder can
consume: a fixed-width bit stream (max_bits_per_coord+1 bits per symbol, packed on the GPU) plus, optionally, the
per-channel frequency tables the coder needs, and reads it back.

Layout (little endian):  magic b"VBQ1" | uint32 N, C, L, flags | uint64 rows | [flags & 1: L*C*Q uint32 counts]
| L streams, each ceil(rows*C*(N+1)/32) uint32 words (symbols in (row, channel) order).
"""
from __future__ import annotations

import struct

import numpy as np
import torch

from . import ops

MAGIC = b"VBQ1"
_HDR = struct.Struct("<4sIIIIQ")


def dumps(qidx, max_bits: int, with_counts: bool = False) -> bytes:
    """qidx: (rows, C) or (L, rows, C) int32 CUDA tensor of sorted quantile indices -> bytes."""
    q = qidx if qidx.dim() == 3 else qidx.unsqueeze(0)
    L, rows, C = q.shape
    Q = ops.num_levels(max_bits)
    out = [_HDR.pack(MAGIC, max_bits, C, L, 1 if with_counts else 0, rows)]
    if with_counts:
        for i in range(L):
            counts = ops.symbol_histogram(q[i], max_bits)
            out.append(counts.to(torch.int32).cpu().numpy().astype("<u4").tobytes())
    for i in range(L):
        out.append(ops.pack_indices(q[i], max_bits).cpu().numpy().astype("<i4").tobytes())
    return b"".join(out)


def loads(data: bytes, device="cuda"):
    """Inverse of dumps: dict(max_bits, qidx (L, rows, C) int32 on `device`, counts (L, C, Q) int64 or None)."""
    magic, N, C, L, flags, rows = _HDR.unpack_from(data, 0)
    if magic != MAGIC:
        raise ValueError("vbq_b200.serialize: bad magic %r" % magic)
    Q = ops.num_levels(N)
    pos = _HDR.size
    counts = None
    if flags & 1:
        n = L * C * Q
        counts = torch.from_numpy(np.frombuffer(data, dtype="<u4", count=n, offset=pos).astype(np.int64))
        counts = counts.reshape(L, C, Q)
        pos += 4 * n
    n_sym = rows * C
    n_words = ops.packed_index_words(n_sym, N)
    qs = []
    for _ in range(L):
        w = np.frombuffer(data, dtype="<i4", count=n_words, offset=pos)
        pos += 4 * n_words
        words = torch.from_numpy(w.copy()).to(device)
        qs.append(ops.unpack_indices(words, n_sym, N).reshape(rows, C))
    return dict(max_bits=N, qidx=torch.stack(qs) if qs else torch.empty((0, rows, C), dtype=torch.int32, device=device),
                counts=counts)


def ideal_code_length_bits(counts) -> float:
    """Σ_c [ n_c log2 n_c - Σ_s k_cs log2 k_cs ]: the length of an ideal adaptive-free entropy code with one frequency
    table per channel (the per-channel analogue of ipynb:452-455)."""
    c = counts.double()
    tot = c.sum(dim=-1)
    t1 = torch.where(tot > 0, tot * torch.log2(tot.clamp(min=1)), torch.zeros_like(tot)).sum()
    t2 = torch.where(c > 0, c * torch.log2(c.clamp(min=1)), torch.zeros_like(c)).sum()
    return float(t1 - t2)
