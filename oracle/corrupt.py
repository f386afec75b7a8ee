"""Oracle: AddSensorNoise (reference src/data/transforms.py:70-154), two restatements.

``add_sensor_noise_torch`` follows the reference line by line with torch ops and the torch generators; run on a
CUDA tensor it yields exactly what the reference yields there (this is the "reference on a CUDA tensor" the CUDA
kernel must match bit for bit).  ``apply_noise_numpy`` is the arithmetic alone, with the random draws passed in,
each fp32 operation rounded separately.  Pinned by tests/golden/corrupt_*.npz (reference run on CPU, seed 42).
"""
from __future__ import annotations

import numpy as np
import torch


def add_sensor_noise_torch(x: torch.Tensor, gaussian_sigma=0.05, speckle_sigma=0.03, dead_band_prob=0.25,
                           thermal_scale=0.01, enable_striping=False, stripe_prob=0.1, return_mask=False):
    is_batched = x.ndim == 4
    if not is_batched:
        x = x.unsqueeze(0)
    B, C, H, W = x.shape
    c = x.clone()
    if gaussian_sigma > 0:                                             # transforms.py:90-92
        c = c + torch.randn_like(c) * gaussian_sigma
    if speckle_sigma > 0:                                              # transforms.py:95-97
        c = c * (1.0 + torch.randn_like(c) * speckle_sigma)
    masks = torch.zeros(B, C, dtype=torch.bool, device=x.device)
    if dead_band_prob > 0:                                             # transforms.py:100-106
        for b in range(B):
            m = torch.rand(C, device=x.device) < dead_band_prob
            masks[b] = m
            if m.sum().item() > 0:
                c[b, m, :, :] = 0.0
    if thermal_scale > 0:                                              # transforms.py:110-121
        w = torch.linspace(1.0, 2.0, C, device=x.device).view(1, C, 1, 1)
        c = c + torch.randn_like(c) * thermal_scale * w
    if enable_striping and torch.rand(1).item() < stripe_prob:         # transforms.py:124-125, 136-154
        c = c + torch.randn(B, C, 1, W, device=x.device) * 0.02
    c = torch.clamp(c, -3.0, 3.0)                                      # transforms.py:128
    if not is_batched:
        c = c.squeeze(0)
    return (c, masks) if return_mask else c


def apply_noise_numpy(x, n1, n2, mask, n3, w, sg, ss, st, stripe=None):
    """x,n1,n2,n3: fp32 [B,C,H,W]; mask bool [B,C]; w fp32 [C]; stripe fp32 [B,C,1,W] or None."""
    f = np.float32
    c = x.astype(f).copy()
    if sg > 0:
        c = (c + (n1 * f(sg)).astype(f)).astype(f)
    if ss > 0:
        c = (c * (f(1.0) + (n2 * f(ss)).astype(f)).astype(f)).astype(f)
    if mask is not None:
        c[mask] = f(0.0)
    if st > 0:
        c = (c + ((n3 * f(st)).astype(f) * w.reshape(1, -1, 1, 1).astype(f)).astype(f)).astype(f)
    if stripe is not None:
        c = (c + (stripe * f(0.02)).astype(f)).astype(f)
    return np.clip(c, f(-3.0), f(3.0))
