"""Batched evaluation driver — the caller of the hot path in the reference
(img-compression/utils.py:502-634 `evaluate_compression_quantizer`, post_process.py:159-188).

The reference loops over images one at a time ("TODO: parallelize", utils.py:535) and per image over the settings.
Here all same-sized images go through ONE `quantizer.compress`-style call (one kernel launch for every image and every
lambda), and the per-image bit totals are reduced on the device.  Only the rate statistics are produced
(`B`, `BPP`, `BPL`, `BPPCL`, the keys of utils.py:523-527); the image-quality metrics (PSNR, MS-SSIM) are out of scope
 and can be computed by the caller from the returned reconstructions."""
from __future__ import annotations

import numpy as np
import torch

from . import utils


def evaluate_compression_quantizer(quantizer, vae, images, settings, return_reconstructions=False, clip=True):
    """``images``: (N, H, W, 3) float array in [0, 1] (same-sized images, e.g. the 24 Kodak images of one
    orientation).  ``settings``: the lambdas.  Returns a dict with (N, len(settings)) arrays ``B`` (total bits),
    ``BPP`` (bits per pixel), ``BPL`` (bits per latent dimension), ``BPPCL`` (bits per pixel from the raw-depth +
    depth-entropy code lengths, utils.py:552-553) and, optionally, ``reconstructions`` (len(settings), N, H, W, 3)."""
    X = images if isinstance(images, torch.Tensor) else torch.as_tensor(np.asarray(images))
    N = int(X.shape[0])
    num_pixels = int(X.shape[1]) * int(X.shape[2])
    M = len(settings)
    means, logvars = vae.encode(X)
    dev = quantizer._compress_latents_device(means, logvars, settings)          # (M, N, H', W', C) tensors on the GPU
    num_bits = dev['num_bits'].reshape(M, N, -1)
    nbits = num_bits.sum(dim=2, dtype=torch.float64)                            # (M, N)
    latent_dims = num_bits.shape[2]
    cl = dev['raw_num_bits'].reshape(M, N, -1).sum(dim=2, dtype=torch.float64) if dev['corrected'] else nbits
    results = {
        'B': utils.to_host_numpy(nbits).T.copy(),
        'BPP': utils.to_host_numpy(nbits / num_pixels).T.copy(),
        'BPL': utils.to_host_numpy(nbits / latent_dims).T.copy(),
        'BPPCL': utils.to_host_numpy(cl / num_pixels).T.copy(),
    }
    if return_reconstructions:
        z = dev['Z_hat'].reshape((-1,) + tuple(means.shape[1:]))
        xh = vae.decode(z)
        xh = xh.detach() if isinstance(xh, torch.Tensor) else torch.as_tensor(np.asarray(xh))
        xh = xh.reshape((M,) + tuple(X.shape))
        if clip:
            xh = xh.clamp(0, 1)
        results['reconstructions'] = utils.to_host_numpy(xh.contiguous())
    return results
