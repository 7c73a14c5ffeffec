"""BMSHJ2018Prior — the learned per-channel factorized prior, with the interface of the reference's
img-compression/learned_prior.py:6-218 (parameter layout, `cdf`, `inverse_cdf`).  Parameter fitting
(`train`, the CLI) is out of scope."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


class BMSHJ2018Prior:
    """Flexible prior of Ballé et al. 2018 (appendix 6.1): per channel a monotone 1-3-3-3-1 network whose
    sigmoid is the CDF.

    Attributes mirror the reference: ``_matrices[k]`` (C, d_{k+1}, d_k) = softplus(raw), ``_biases[k]``
    (C, d_{k+1}, 1), ``_factors[k]`` (C, d_{k+1}, 1) = tanh(raw)  (learned_prior.py:30-58)."""

    def __init__(self, channels, dims=(3, 3, 3), init_scale=10., device="cuda", seed=None, **kwargs):
        self._channels = int(channels)
        self._init_scale = float(init_scale)
        self._dims = tuple(int(f) for f in dims)
        if self._dims != (3, 3, 3):
            raise NotImplementedError("vbq_b200 kernels implement dims=(3,3,3), the reference default "
                                      "(learned_prior.py:10, post_process.py:75)")
        self.device = torch.device(device)
        fdims = (1,) + self._dims + (1,)
        scale = self._init_scale ** (1 / (len(self._dims) + 1))
        rng = np.random.default_rng(seed)
        self.raw_matrices, self.raw_biases, self.raw_factors = [], [], []
        for i in range(len(self._dims) + 1):
            init = np.log(np.expm1(1 / scale / fdims[i + 1]))
            self.raw_matrices.append(np.full((channels, fdims[i + 1], fdims[i]), init, dtype=np.float32))
            self.raw_biases.append(rng.uniform(-.5, .5, size=(channels, fdims[i + 1], 1)).astype(np.float32))
            if i < len(self._dims):
                self.raw_factors.append(np.zeros((channels, fdims[i + 1], 1), dtype=np.float32))
        self._refresh()

    # -- parameters ------------------------------------------------------------------------------------
    def _refresh(self):
        self._matrices = [np.logaddexp(0.0, m).astype(np.float32) for m in self.raw_matrices]
        self._biases = [b.astype(np.float32) for b in self.raw_biases]
        self._factors = [np.tanh(f).astype(np.float32) for f in self.raw_factors]
        self._packed = None

    def set_raw_parameters(self, matrices=None, biases=None, factors=None):
        """Install trained raw variables (`matrix_k`, `bias_k`, `factor_k` of the reference checkpoint)."""
        if matrices is not None:
            self.raw_matrices = [np.asarray(m, dtype=np.float32) for m in matrices]
        if biases is not None:
            self.raw_biases = [np.asarray(b, dtype=np.float32) for b in biases]
        if factors is not None:
            self.raw_factors = [np.asarray(f, dtype=np.float32) for f in factors]
        self._refresh()

    def set_transformed_parameters(self, matrices, biases, factors):
        """Install already-transformed `_matrices`, `_biases`, `_factors` (what the reference's CDF reads)."""
        self._matrices = [np.asarray(m, dtype=np.float32) for m in matrices]
        self._biases = [np.asarray(b, dtype=np.float32) for b in biases]
        self._factors = [np.asarray(f, dtype=np.float32) for f in factors]
        self._packed = None

    @property
    def init_scale(self):
        return self._init_scale

    @property
    def dims(self):
        return self._dims

    @property
    def channels(self):
        return self._channels

    def packed_params(self):
        """(C, 43) float32 CUDA tensor in the layout of include/vbq_b200.h."""
        if self._packed is None:
            C = self._channels
            cols = []
            for k in range(4):
                cols.append(self._matrices[k].reshape(C, -1))
                cols.append(self._biases[k].reshape(C, -1))
                if k < 3:
                    cols.append(self._factors[k].reshape(C, -1))
            host = np.ascontiguousarray(np.concatenate(cols, axis=1), dtype=np.float32)
            assert host.shape == (C, 43)
            self._packed = torch.from_numpy(host).to(self.device)
        return self._packed

    # -- reference interface ---------------------------------------------------------------------------
    def cdf(self, inputs, stop_gradient=True):
        """CDF of channel-last inputs (..., C)  (learned_prior.py:109-148)."""
        x = inputs if isinstance(inputs, torch.Tensor) else torch.as_tensor(np.asarray(inputs))
        assert int(x.shape[-1]) == self._channels, \
            'Innermost dimension of inputs = %d, does not match number of channels = %d' % \
            (int(x.shape[-1]), self._channels)
        x2 = x.to(device=self.device, dtype=torch.float32).reshape(-1, self._channels).contiguous()
        return ops.learned_cdf(self.packed_params(), x2).reshape(x.shape)

    def inverse_cdf(self, xi, method='bisection', max_iterations=1000, tol=1e-9, **kwargs):
        """Quantile function for xi (..., C) in (0,1)  (learned_prior.py:173-218).  ``max_iterations`` and ``tol``
        are accepted for compatibility: the kernel iterates to float64 convergence and rounds to float32.
        Returns a float32 CUDA tensor, or an ndarray with ``return_np=True``."""
        if method != 'bisection':
            raise NotImplementedError  # as the reference (learned_prior.py:231)
        x = xi if isinstance(xi, torch.Tensor) else torch.as_tensor(np.asarray(xi, dtype=np.float64))
        assert int(x.shape[-1]) == self._channels
        x2 = x.to(device=self.device, dtype=torch.float64).reshape(-1, self._channels).contiguous()
        z = ops.learned_inverse_cdf(self.packed_params(), x2).reshape(x.shape)
        if kwargs.get('return_np', False):
            return z.cpu().numpy()
        return z

    def build_code_points_device(self, max_bits):
        """(C, Q) heap-order table from the same device routine as `inverse_cdf` (vbq_build_code_points_learned)."""
        return ops.build_code_points_learned(self.packed_params(), int(max_bits))
