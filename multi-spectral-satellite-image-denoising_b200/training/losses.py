"""Training loss -- drop-in for the reference's ``src.training.losses.CombinedLoss`` / ``SSIMLoss``
(src/training/losses.py:16-189): ``CombinedLoss(mse_weight, ssim_weight, data_range)(pred, target)`` returns the
0-dim tensor ``mse_weight * MSE + ssim_weight * (1 - SSIM)`` with autograd.

Forward and the analytic gradient wrt ``pred`` come from ONE library call (csrc/loss.cu): the gradient is produced
eagerly while the blurred statistics are still on chip / in L2 and is handed to autograd in ``backward`` scaled by
the incoming grad_output, instead of replaying five depthwise convolutions through autograd.
Always evaluated in fp32 (the FP32 result is the parity target; the reference's fp16-autocast SSIM is not,
SURVEY 3.1).
"""
from __future__ import annotations

import logging

import torch
import torch.nn as nn

from .. import _lib

logger = logging.getLogger(__name__)


class _LossFn(torch.autograd.Function):
    """returns fp32 [3] = {total, mse, 1 - ssim}; only element 0 carries gradient."""

    @staticmethod
    def forward(ctx, pred, target, w_mse, w_ssim, data_range):
        if not (pred.is_cuda and target.is_cuda):
            raise _lib.MsdError("loss: CUDA tensors required (libmsd_b200 has no CPU path)")
        if pred.shape != target.shape or pred.ndim != 4:
            raise _lib.MsdError(f"loss: expected matching [B,C,H,W], got {tuple(pred.shape)} / {tuple(target.shape)}")
        p32 = pred.detach().float().contiguous()
        t32 = target.detach().float().contiguous()
        B, C_, H, W = p32.shape
        lib = _lib.load()
        _lib.require_device(p32.device)
        need_grad = ctx.needs_input_grad[0]
        with torch.cuda.device(p32.device):
            out = torch.empty(3, dtype=torch.float32, device=p32.device)
            ws_bytes = lib.msd_loss_workspace_bytes(B, C_, H, W) if need_grad else 256
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=p32.device)
            grad = torch.empty_like(p32) if need_grad else None
            _lib.check(lib.msd_loss_fwd_bwd_f32(_lib.ptr(p32), _lib.ptr(t32), B, C_, H, W, float(w_mse), float(w_ssim),
                                                float(data_range), 1.0, _lib.ptr(out), _lib.ptr(grad), _lib.ptr(ws),
                                                ws_bytes, _lib.stream_ptr()))
        ctx.pred_dtype = pred.dtype
        ctx.save_for_backward(grad)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        g = (grad * grad_out[0]).to(ctx.pred_dtype) if grad is not None else None
        return g, None, None, None, None


class SSIMLoss(nn.Module):
    """1 - SSIM with an 11x11 Gaussian window, sigma 1.5, zero padding (losses.py:80-189)."""

    def __init__(self, window_size: int = 11, size_average: bool = True, data_range: float = 6.0, channel: int = 13):
        super().__init__()
        if window_size != 11:
            raise _lib.MsdError("SSIMLoss: only window_size=11 is built into libmsd_b200")
        if not size_average:
            raise _lib.MsdError("SSIMLoss: only size_average=True is built into libmsd_b200")
        self.window_size = window_size
        self.size_average = size_average
        self.channel = channel
        self.data_range = data_range

    def forward(self, img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
        # total = 0*MSE + 1*(1-SSIM)
        return _LossFn.apply(img1, img2, 0.0, 1.0, self.data_range)[0]


class CombinedLoss(nn.Module):
    """mse_weight * MSE + ssim_weight * (1 - SSIM) (losses.py:16-77)."""

    def __init__(self, mse_weight: float = 1.0, ssim_weight: float = 0.1, data_range: float = 6.0):
        super().__init__()
        self.mse_weight = mse_weight
        self.ssim_weight = ssim_weight
        self.mse = nn.MSELoss()          # kept for attribute parity with the reference; not used in forward
        self.ssim = SSIMLoss(data_range=data_range, channel=13)
        self.last_terms = None           # device fp32 [3] = {total, mse, 1-ssim} of the last call
        logger.info("Initialized CombinedLoss: MSE weight=%s, SSIM weight=%s", mse_weight, ssim_weight)

    def forward(self, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        out = _LossFn.apply(pred, target, self.mse_weight, self.ssim_weight, self.ssim.data_range)
        self.last_terms = out.detach()
        return out[0]
