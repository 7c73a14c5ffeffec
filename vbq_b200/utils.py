"""Host-side helpers with the names of the reference's img-compression/utils.py (hot-path subset)."""
from __future__ import annotations

import numpy as np
import torch


def n_bit_binary_floats(n):
    """Level-n quantiles xi = (i + 1/2) 2^-n, i < 2^n  (reference utils.py:23-24)."""
    return [i * 2 ** (-n) + 2 ** (-n - 1) for i in range(2 ** n)]


def all_bin_floats(max_bits):
    """All Q = 2^(N+1)-1 quantiles in heap order (reference quantizer.py:30)."""
    return np.hstack([n_bit_binary_floats(n) for n in range(max_bits + 1)])


def heap_to_sorted_index(level, index, max_bits):
    """Sorted rank q = (2i+1) 2^(N-n) - 1 of heap entry (n, i) (reference quantizer.py:37,135)."""
    return (2 * index + 1) * (1 << (max_bits - level)) - 1


class NormalLogpdf:
    """Callable returned by `curry_normal_logpdf` (reference utils.py:307-327): f(z) = -0.5*((z-loc)/scale)^2
    (+ const).  It carries loc/scale so that `batch_quantize_indep_dims` can hand them to the CUDA kernel
    instead of materialising f(P)."""

    def __init__(self, loc, scale, ignore_const):
        self.loc, self.scale, self.ignore_const = loc, scale, ignore_const

    def __call__(self, z):
        out = -0.5 * ((z - self.loc) / self.scale) ** 2
        if not self.ignore_const:
            out = out - torch.log(torch.as_tensor(self.scale)) - 0.5 * float(np.log(2 * np.pi))
        return out


def curry_normal_logpdf(loc, scale, ignore_const=False, backend=None):
    """Reference utils.py:307-327.  ``backend`` is accepted for signature compatibility and ignored."""
    return NormalLogpdf(loc, scale, ignore_const)


def as_device_f32(x, device):
    """numpy / torch (any device) -> contiguous float32 CUDA tensor."""
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float32).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).to(device)


def to_host_numpy(t):
    """Device tensor -> NumPy array through a pinned staging buffer (PyTorch caches pinned blocks, so repeated calls
    of the same size pay one DMA at PCIe speed instead of a pageable copy)."""
    if not t.is_cuda:
        return t.numpy()
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return h.numpy()




def batch_quantize_indep_dims(Z_shape, code_points, code_lengths, fun, lambs, backend=None, return_np=True,
                              device="cuda"):
    """Reference utils.py:363-423: for every batch element and dimension pick the candidate maximising
    fun(P) - lamb * L, first maximum, for every lamb.  ``code_points`` / ``code_lengths`` are (K, M) or (M, B, K)
    (lengths may also be (Lambda, M, B, K)).  Runs on the GPU (`vbq_argmax_candidates`); ``backend`` is accepted for
    signature compatibility and ignored.  ``fun`` may be the object returned by `curry_normal_logpdf` (then the
    float32 Gaussian score is computed inside the kernel) or any callable on torch tensors."""
    from . import ops
    B, K = Z_shape
    dev = torch.device(device)

    def to_dev(x, dtype=None):
        t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(x))
        if dtype is None:
            dtype = torch.float32 if t.dtype.is_floating_point else torch.int32
        return t.to(device=dev, dtype=dtype)

    P = to_dev(code_points, torch.float32)
    L = to_dev(code_lengths)
    if P.dim() == 2:                                             # (K, M) -> (M, B, K), utils.py:385-386
        P = P.t()[:, None, :].expand(-1, B, -1)
        L = L.t()[:, None, :].expand(-1, B, -1)
    else:
        assert P.dim() == 3
    P, L = P.contiguous(), L.contiguous()
    if isinstance(fun, NormalLogpdf) and fun.ignore_const:
        loc = to_dev(fun.loc, torch.float32).expand(B, K).contiguous()
        scale = to_dev(fun.scale, torch.float32).expand(B, K).contiguous()
        zhat, bits, _ = ops.argmax_candidates(P, L, lambs, loc=loc, scale=scale)
    else:
        fun_P = fun(P)
        fun_P = to_dev(fun_P, torch.float32).contiguous()
        zhat, bits, _ = ops.argmax_candidates(P, L, lambs, fun_P=fun_P)
    Z_hat_dict, num_bits_dict = {}, {}
    for i, lamb in enumerate(lambs):
        z, nb = zhat[i], bits[i]
        if return_np:
            z, nb = z.cpu().numpy(), nb.cpu().numpy()
        Z_hat_dict[lamb] = z
        num_bits_dict[lamb] = nb
    return Z_hat_dict, num_bits_dict
