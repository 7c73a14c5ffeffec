"""Gaussian priors and the VAE contract of the reference's img-compression/vae_models.py:14-72 (interfaces
only: the conv nets and their training are out of scope)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _xi_to_device(xi, device):
    x = xi if isinstance(xi, torch.Tensor) else torch.as_tensor(np.asarray(xi, dtype=np.float64))
    return x.to(device=device, dtype=torch.float64)


class StandardGaussianPrior:
    """N(0,1) prior (vae_models.py:14-25); `inverse_cdf(xi)` = norm.ppf(xi) evaluated on the GPU in float64."""
    device = "cuda"

    @classmethod
    def inverse_cdf(cls, xi):
        x = _xi_to_device(xi, cls.device)
        z = ops.gaussian_inverse_cdf(x.reshape(-1, x.shape[-1]).contiguous(), None, None).reshape(x.shape)
        return z if isinstance(xi, torch.Tensor) else z.cpu().numpy()


class FactoredGaussianPrior:
    """Per-channel N(mean[c], std[c]^2) (vae_models.py:28-43)."""

    def __init__(self, mean, std, device="cuda"):
        self.mean = np.asarray(mean, dtype=np.float64)
        self.std = np.asarray(std, dtype=np.float64)
        self.logvar = 2 * np.log(self.std)
        self.device = torch.device(device)

    def inverse_cdf(self, xi):
        assert xi.shape[-1] == len(self.mean)
        x = _xi_to_device(xi, self.device)
        m = torch.from_numpy(self.mean).to(self.device)
        s = torch.from_numpy(self.std).to(self.device)
        z = ops.gaussian_inverse_cdf(x.reshape(-1, x.shape[-1]).contiguous(), m, s).reshape(x.shape)
        return z if isinstance(xi, torch.Tensor) else z.cpu().numpy()

    def build_code_points_device(self, max_bits):
        m = torch.from_numpy(self.mean).to(self.device)
        s = torch.from_numpy(self.std).to(self.device)
        return ops.build_code_points_gaussian(m, s, int(max_bits))


class GaussianVAE:
    """The encode/decode contract the quantizer relies on (vae_models.py:46-72): `encode(x)` returns channel-last
    (mean, logvar); `decode(z)` maps latents of the same shape back to data space.  The networks are supplied by
    the caller as callables on torch tensors."""

    def __init__(self, prior, inference_net, generative_net, decode_sigmoid=False):
        self.prior = prior
        self.inference_net = inference_net
        self.generative_net = generative_net
        self.decode_sigmoid = decode_sigmoid

    def encode(self, x):
        out = self.inference_net(x)
        mean, logvar = torch.chunk(out, 2, dim=-1)
        return mean, logvar

    def decode(self, z):
        mean = self.generative_net(z)
        if self.decode_sigmoid:
            mean = torch.sigmoid(mean)
        return mean
