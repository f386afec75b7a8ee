"""Oracle: CombinedLoss / SSIMLoss (reference src/training/losses.py:16-189) as a torch fp32 function, autograd for
the gradient.  (A torch fp32 reference is kept here because this is a floating-point kernel with a backward pass.)
Pinned by tests/golden/loss_*.npz."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _window(channel: int, size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    x = torch.arange(size).float() - size // 2
    g = torch.exp(-(x ** 2) / (2 * sigma ** 2))
    g = (g / g.sum()).unsqueeze(1)
    return g.mm(g.t()).unsqueeze(0).unsqueeze(0).expand(channel, 1, size, size).contiguous()


def ssim_mean(img1: torch.Tensor, img2: torch.Tensor, data_range: float = 6.0) -> torch.Tensor:
    ch = img1.shape[1]
    w = _window(ch).to(img1)
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mu1, mu2 = F.conv2d(img1, w, padding=5, groups=ch), F.conv2d(img2, w, padding=5, groups=ch)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, w, padding=5, groups=ch) - mu1_sq
    s2 = F.conv2d(img2 * img2, w, padding=5, groups=ch) - mu2_sq
    s12 = F.conv2d(img1 * img2, w, padding=5, groups=ch) - mu1_mu2
    m = ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return m.mean()


def combined_loss(pred: torch.Tensor, target: torch.Tensor, mse_weight=1.0, ssim_weight=0.1, data_range=6.0):
    """returns (total, mse, 1-ssim) as 0-dim tensors (losses.py:67-77)."""
    mse = F.mse_loss(pred, target)
    ssim_loss = 1 - ssim_mean(pred, target, data_range)
    return mse_weight * mse + ssim_weight * ssim_loss, mse, ssim_loss


def combined_loss_and_grad(pred: torch.Tensor, target: torch.Tensor, mse_weight=1.0, ssim_weight=0.1, data_range=6.0,
                           dtype=torch.float32):
    p = pred.detach().to(dtype).clone().requires_grad_(True)
    total, mse, sl = combined_loss(p, target.detach().to(dtype), mse_weight, ssim_weight, data_range)
    total.backward()
    return total.item(), mse.item(), sl.item(), p.grad.detach()
