"""Batch/row sharding over the GPUs of one node.

Every coordinate is independent given the (replicated, <= 55 KB) prior and the lambda list, so ranks take
contiguous slices of the leading axis (images or embedding rows) and never exchange latents.  The only
collective is one all-reduce(SUM) of the (n_lambda, 4) float64 rate/distortion totals; entropy-model fitting adds
an int64 histogram all-reduce.  One process per GPU, `torch.distributed` (NCCL over NVLink on the B200 box, gloo in
the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items, world_size, rank):
    """Contiguous split of ``n_items`` leading-axis units: the first ``n_items % world_size`` ranks get one
    extra.  Returns (start, stop)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank %d / world size %d" % (rank, world_size))
    base, rem = divmod(int(n_items), world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_leading_axis(x, world_size=None, rank=None):
    """This rank's contiguous slice of ``x`` along axis 0."""
    world_size = dist.get_world_size() if world_size is None else world_size
    rank = dist.get_rank() if rank is None else rank
    a, b = shard_bounds(x.shape[0], world_size, rank)
    return x[a:b]


def all_reduce_totals(totals, group=None, async_op=False):
    """Sum the per-shard (n_lambda, 4) float64 totals over all ranks, in place; returns ``totals`` (or, with
    ``async_op``, the work handle to wait on — None when there is nothing to reduce).
    Columns: sum of raw depth n, sum of code length, sum of entropy-model bits, sum (z_hat-mu)^2/(2 sigma^2)."""
    if totals.dtype != torch.float64:
        raise TypeError("totals must be float64")
    work = None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        work = dist.all_reduce(totals, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return work if async_op else totals


def all_reduce_counts(counts, group=None):
    """Sum int64 histogram counts over all ranks (entropy-model fitting, reference quantizer.py:104-105,138-140)."""
    if counts.dtype != torch.int64:
        raise TypeError("counts must be int64")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


class PeerTotals:
    """All-reduce of the (n_lambda, 4) totals WITHOUT a collective launch (include/vbq_b200.h, csrc/peer.cu): the last CTA
    of the search kernel writes the call's sums into every rank's inbox over NVLink; `collect(seq)` enqueues the wait for
    all ranks' contributions to call `seq` and their sum in rank order.  One object per rank; the constructor exchanges
    the 64-byte memory handles of the inboxes through torch.distributed (any backend).

        peer = PeerTotals(n_lambda_max=64)
        plan = ops.QuantizePlan(..., totals=local_totals, peer=peer)
        plan.run(); seq = peer.next_seq()       # call number seq = 1, 2, 3, ... identical on all ranks
        peer.push(seq, local_totals)            # or: the NEXT call does it while it runs, plan.run_peer(seq, local_totals, ...)
        peer.collect(seq, n_lambda, global_totals)

    Up to 8 calls may be delivered before the oldest is collected."""

    def __init__(self, n_lambda_max=64, group=None, device=None):
        import ctypes
        from . import _lib
        self._lib = _lib.load()
        self.group = group
        on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if on else 0
        self.world = dist.get_world_size(group) if on else 1
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.seq = 0
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.vbq_peer_ctx_create(self.rank, self.world, int(n_lambda_max), ctypes.byref(h)),
                       "vbq_peer_ctx_create")
            self._h = h
            if self.world > 1:
                mine = torch.zeros(64, dtype=torch.uint8)
                _lib.check(self._lib.vbq_peer_ctx_handle(self._h, mine.data_ptr()), "vbq_peer_ctx_handle")
                if dist.get_backend(group) == "nccl":
                    every = [torch.zeros(64, dtype=torch.uint8, device=self.device) for _ in range(self.world)]
                    dist.all_gather(every, mine.to(self.device), group=group)
                    every = torch.stack(every).cpu()
                else:
                    every = [torch.zeros(64, dtype=torch.uint8) for _ in range(self.world)]
                    dist.all_gather(every, mine, group=group)
                    every = torch.stack(every)
                every = every.contiguous()
                _lib.check(self._lib.vbq_peer_ctx_connect(self._h, every.data_ptr()), "vbq_peer_ctx_connect")
                dist.barrier(group)     # nobody writes into an inbox that is not mapped everywhere yet

    def next_seq(self):
        self.seq += 1
        return self.seq

    def push(self, seq, totals):
        """Enqueue (on the current stream) the delivery of the (n_lambda, 4) totals of call `seq` to every rank."""
        from . import _lib, ops
        with torch.cuda.device(self.device):
            _lib.check(self._lib.vbq_peer_push(self._h, int(seq), int(totals.shape[0]), totals.data_ptr(),
                                               ops._stream(self.device)), "vbq_peer_push")

    def collect(self, seq, n_lambda, out):
        """Enqueue (on the current stream) the sum over ranks of call `seq` into `out` (n_lambda, 4) float64."""
        from . import _lib, ops
        if out.dtype != torch.float64 or tuple(out.shape) != (n_lambda, _lib.TOTALS) or not out.is_cuda:
            raise ValueError("vbq_b200: `out` must be a (n_lambda, %d) float64 CUDA tensor" % _lib.TOTALS)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.vbq_peer_collect(self._h, int(seq), int(n_lambda), out.data_ptr(),
                                                  ops._stream(self.device)), "vbq_peer_collect")
        return out

    def close(self):
        if getattr(self, "_h", None) is not None:
            torch.cuda.synchronize(self.device)
            if self.world > 1 and dist.is_initialized():
                dist.barrier(self.group)    # nobody unmaps an inbox that a peer may still write to
            self._lib.vbq_peer_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                self._lib.vbq_peer_ctx_destroy(self._h)
                self._h = None
        except Exception:
            pass


class ShardedQuantizer:
    """Data-parallel wrapper: each rank quantizes its slice and the per-lambda totals are all-reduced."""

    def __init__(self, quantizer, group=None):
        self.quantizer = quantizer
        self.group = group

    def rd_sweep(self, local_means, local_scales, lambs, logvar=False, entropy_bits=False, flags=0):
        """Totals-only rate-distortion sweep over this rank's shard; returns the global (n_lambda, 4) totals."""
        from . import ops
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            flags |= ops.FLAG_RESERVE_SM      # leave an SM to NCCL kernels of neighbouring calls
        out = self.quantizer.quantize(local_means, local_scales, lambs, logvar=logvar, outputs=ops.OUT_TOTALS,
                                      flags=flags, entropy_bits=entropy_bits)
        return all_reduce_totals(out['totals'], self.group)

    def quantize(self, local_means, local_scales, lambs, outputs, logvar=False, entropy_bits=False, flags=0):
        """Full outputs for this rank's shard (they stay on the rank) plus the GLOBAL totals: one all-reduce per call."""
        from . import ops
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            flags |= ops.FLAG_RESERVE_SM
        out = self.quantizer.quantize(local_means, local_scales, lambs, logvar=logvar, outputs=outputs | ops.OUT_TOTALS,
                                      flags=flags, entropy_bits=entropy_bits)
        all_reduce_totals(out['totals'], self.group)
        return out

    def build_entropy_models_from_latents(self, local_means, local_logvars, lambs, add_n_smoothing):
        return self.quantizer.build_entropy_models_from_latents(
            local_means, local_logvars, lambs, add_n_smoothing,
            reduce_fn=lambda c: all_reduce_counts(c, self.group))
