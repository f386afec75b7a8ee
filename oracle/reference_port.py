"""Oracle / CPU baseline: the reference's hot path restated op-for-op in torch eager fp32, the way the reference
itself computes it on the CPU (depthwise F.conv2d SSIM, permute+reshape SAM, one .item() per metric per image).

Used ONLY as (a) bench.py's ``cpu_baseline`` and ``--impl reference`` arm -- the Python reference cannot travel to the
GPU box, so its CPU path is timed through this port (kind "port") -- and (b) an extra cross-check of the numpy oracle.
Citations: src/training/metrics.py:19-287, src/data/transforms.py:70-134 (oracle/corrupt.py), src/models (oracle/model.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import corrupt as _corrupt
from . import model as _model


def calculate_psnr(pred, target, data_range=6.0):                      # metrics.py:39-46
    mse = F.mse_loss(pred, target).item()
    if mse == 0:
        return float("inf")
    return 10 * np.log10((data_range ** 2) / mse)


def calculate_ssim(pred, target, data_range=6.0, window_size=11):      # metrics.py:68-121
    if pred.ndim == 3:
        pred, target = pred.unsqueeze(0), target.unsqueeze(0)
    C = pred.shape[1]
    x = torch.arange(window_size).float() - window_size // 2
    g = torch.exp(-(x ** 2) / (2 * 1.5 ** 2))
    g = g / g.sum()
    w = g.unsqueeze(1).mm(g.unsqueeze(0)).unsqueeze(0).unsqueeze(0).expand(C, 1, window_size, window_size)
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    pad = window_size // 2
    mu1, mu2 = F.conv2d(pred, w, padding=pad, groups=C), F.conv2d(target, w, padding=pad, groups=C)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 ** 2, mu2 ** 2, mu1 * mu2
    s1 = F.conv2d(pred ** 2, w, padding=pad, groups=C) - mu1_sq
    s2 = F.conv2d(target ** 2, w, padding=pad, groups=C) - mu2_sq
    s12 = F.conv2d(pred * target, w, padding=pad, groups=C) - mu1_mu2
    m = ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return m.mean().item()


def calculate_sam(pred, target, epsilon=1e-8):                         # metrics.py:145-172
    if pred.ndim == 3:
        pred, target = pred.unsqueeze(0), target.unsqueeze(0)
    C = pred.shape[1]
    p = pred.permute(0, 2, 3, 1).reshape(-1, C)
    t = target.permute(0, 2, 3, 1).reshape(-1, C)
    cos = (p * t).sum(dim=1) / ((torch.norm(p, dim=1) + epsilon) * (torch.norm(t, dim=1) + epsilon))
    return torch.rad2deg(torch.acos(torch.clamp(cos, -1.0, 1.0))).mean().item()


def calculate_rmse(pred, target):                                      # metrics.py:189-191
    return np.sqrt(F.mse_loss(pred, target).item())


class MetricsTracker:                                                  # metrics.py:212-287
    def __init__(self, data_range=6.0):
        self.data_range = data_range
        self.psnr_sum = self.ssim_sum = self.sam_sum = self.rmse_sum = 0.0
        self.count = 0

    @torch.no_grad()
    def update(self, pred, target):
        for i in range(pred.shape[0]):
            p, t = pred[i:i + 1], target[i:i + 1]
            self.psnr_sum += calculate_psnr(p, t, self.data_range)
            self.ssim_sum += calculate_ssim(p, t, self.data_range)
            self.sam_sum += calculate_sam(p, t)
            self.rmse_sum += calculate_rmse(p, t)
            self.count += 1

    def compute(self):
        n = max(self.count, 1)
        return {"psnr": self.psnr_sum / n, "ssim": self.ssim_sum / n, "sam": self.sam_sum / n, "rmse": self.rmse_sum / n}


@torch.no_grad()
def hot_path_step(sd: dict, clean: torch.Tensor, noise_params, tracker: MetricsTracker) -> torch.Tensor:
    """One pass of the reference's evaluation hot path on the CPU: corrupt -> restore -> per-image metrics."""
    noisy = _corrupt.add_sensor_noise_torch(clean, *noise_params)
    restored = _model.restoration_forward(sd, noisy)
    tracker.update(restored, clean)
    return restored
