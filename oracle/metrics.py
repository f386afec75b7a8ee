"""Oracle: PSNR / SSIM / SAM / RMSE / MAE and MetricsTracker (reference src/training/metrics.py:19-287) in numpy.

fp32 arithmetic where the reference computes in fp32 (the window, the five blurs, the SSIM map, the SAM angle),
fp64 where the reference leaves torch (np.log10, np.sqrt on the .item() value).  The 11x11 window is applied as
the reference applies it -- a full 2-D correlation with zero padding 5 -- not separably, so that the CUDA kernel's
separable evaluation is checked against an independent formulation.
Pinned by tests/golden/metrics_*.npz (reference values, CPU fp32).
"""
from __future__ import annotations

import numpy as np

F = np.float32


def gaussian_window(size: int = 11, sigma: float = 1.5) -> np.ndarray:
    """metrics.py:76-79 / losses.py:111-114 in fp32."""
    x = np.arange(size, dtype=F) - F(size // 2)
    g = np.exp(-(x ** 2) / F(2 * sigma ** 2)).astype(F)
    return (g / g.sum(dtype=F)).astype(F)


def _blur2d(img: np.ndarray, w2: np.ndarray) -> np.ndarray:
    """zero-padded 'same' correlation of [..., H, W] with the 2-D window (F.conv2d(padding=5, groups=C))."""
    r = w2.shape[0] // 2
    H, W = img.shape[-2:]
    pad = np.zeros(img.shape[:-2] + (H + 2 * r, W + 2 * r), dtype=F)
    pad[..., r:r + H, r:r + W] = img
    out = np.zeros_like(img, dtype=F)
    for i in range(w2.shape[0]):
        for j in range(w2.shape[1]):
            out += w2[i, j] * pad[..., i:i + H, j:j + W]
    return out


def ssim_map(pred: np.ndarray, target: np.ndarray, data_range: float = 6.0, window_size: int = 11) -> np.ndarray:
    pred, target = pred.astype(F), target.astype(F)
    g = gaussian_window(window_size)
    w2 = np.outer(g, g).astype(F)
    C1, C2 = F((0.01 * data_range) ** 2), F((0.03 * data_range) ** 2)
    mu1, mu2 = _blur2d(pred, w2), _blur2d(target, w2)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = _blur2d(pred * pred, w2) - mu1_sq
    s2 = _blur2d(target * target, w2) - mu2_sq
    s12 = _blur2d(pred * target, w2) - mu1_mu2
    return ((F(2) * mu1_mu2 + C1) * (F(2) * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))


def calculate_ssim(pred, target, data_range: float = 6.0, window_size: int = 11) -> float:
    return float(ssim_map(np.asarray(pred), np.asarray(target), data_range, window_size).mean(dtype=np.float64).astype(F))


def _mse(pred, target) -> float:
    d = np.asarray(pred, dtype=F) - np.asarray(target, dtype=F)
    return float(F((d.astype(np.float64) ** 2).mean()))


def calculate_psnr(pred, target, data_range: float = 6.0) -> float:
    mse = _mse(pred, target)
    if mse == 0:
        return float("inf")
    return float(10 * np.log10((data_range ** 2) / mse))


def calculate_rmse(pred, target) -> float:
    return float(np.sqrt(_mse(pred, target)))


def calculate_mae(pred, target) -> float:
    d = np.abs(np.asarray(pred, dtype=F) - np.asarray(target, dtype=F))
    return float(F(d.mean(dtype=np.float64)))


def calculate_sam(pred, target, epsilon: float = 1e-8) -> float:
    p, t = np.asarray(pred, dtype=F), np.asarray(target, dtype=F)
    if p.ndim == 3:
        p, t = p[None], t[None]
    dot = (p * t).sum(axis=1, dtype=F)
    pn = np.sqrt((p * p).sum(axis=1, dtype=F)).astype(F) + F(epsilon)
    tn = np.sqrt((t * t).sum(axis=1, dtype=F)).astype(F) + F(epsilon)
    cos = np.clip(dot / (pn * tn), F(-1), F(1))
    ang = np.arccos(cos).astype(F) * F(180.0 / np.pi)
    return float(F(ang.mean(dtype=np.float64)))


class MetricsTracker:
    """metrics.py:212-287: per-image metrics, running sums, means."""

    def __init__(self, data_range: float = 6.0):
        self.data_range = data_range
        self.reset()

    def reset(self):
        self.psnr_sum = self.ssim_sum = self.sam_sum = self.rmse_sum = 0.0
        self.count = 0

    def update(self, pred, target):
        pred, target = np.asarray(pred), np.asarray(target)
        for i in range(pred.shape[0]):
            self.psnr_sum += calculate_psnr(pred[i:i + 1], target[i:i + 1], self.data_range)
            self.ssim_sum += calculate_ssim(pred[i:i + 1], target[i:i + 1], self.data_range)
            self.sam_sum += calculate_sam(pred[i:i + 1], target[i:i + 1])
            self.rmse_sum += calculate_rmse(pred[i:i + 1], target[i:i + 1])
            self.count += 1

    def compute(self) -> dict:
        if self.count == 0:
            return {"psnr": 0.0, "ssim": 0.0, "sam": 0.0, "rmse": 0.0}
        n = self.count
        return {"psnr": self.psnr_sum / n, "ssim": self.ssim_sum / n, "sam": self.sam_sum / n, "rmse": self.rmse_sum / n}
