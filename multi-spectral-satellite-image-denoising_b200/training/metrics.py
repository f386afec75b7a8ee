"""Evaluation metrics -- drop-in for the reference's ``src.training.metrics`` (src/training/metrics.py:19-287):
``calculate_psnr / calculate_ssim / calculate_sam / calculate_rmse / calculate_mae`` return Python floats,
``MetricsTracker`` keeps per-image running sums and ``compute()`` returns their means.

All arithmetic runs in csrc/metrics.cu.  Where the reference issues 4*B ``.item()`` syncs per batch
(metrics.py:247-256), the tracker here accumulates {sum psnr, sum ssim, sum sam, sum rmse, count} in an fp64[5]
DEVICE tensor and touches the host only in ``compute()`` (one 40-byte copy); ``all_reduce()`` sums that tensor
over ranks for the batch-sharded evaluation sweep.
"""
from __future__ import annotations

import math

import torch

from .. import _lib


def _as_bchw(pred: torch.Tensor, target: torch.Tensor):
    if pred.shape != target.shape:
        raise _lib.MsdError(f"metrics: shape mismatch {tuple(pred.shape)} vs {tuple(target.shape)}")
    if not (pred.is_cuda and target.is_cuda):
        raise _lib.MsdError("metrics: CUDA tensors required (libmsd_b200 has no CPU path)")
    if pred.ndim == 3:
        pred, target = pred.unsqueeze(0), target.unsqueeze(0)
    elif pred.ndim != 4:
        pred, target = pred.reshape(1, 1, 1, -1), target.reshape(1, 1, 1, -1)
    # the reference promotes mismatched / half inputs to fp32 (metrics.py:84-93; trainer.py:217 calls .float())
    return pred.detach().float().contiguous(), target.detach().float().contiguous()


def fused_metrics(pred: torch.Tensor, target: torch.Tensor, data_range: float = 6.0):
    """One pass of the fused kernels over a batch.

    Returns (per_image fp64 [B,4] = psnr/ssim/sam/rmse, sums fp64 [5], raw fp64 [B,4] = sum sq err, sum ssim_map,
    sum sam degrees, sum |err|), all on the device, no host sync.
    """
    pred, target = _as_bchw(pred, target)
    B, C_, H, W = pred.shape
    lib = _lib.load()
    _lib.require_device(pred.device)
    with torch.cuda.device(pred.device):
        per_image = torch.empty(B, 4, dtype=torch.float64, device=pred.device)
        sums = torch.empty(5, dtype=torch.float64, device=pred.device)
        raw = torch.empty(B, 4, dtype=torch.float64, device=pred.device)
        _lib.check(lib.msd_metrics_f32(_lib.ptr(pred), _lib.ptr(target), B, C_, H, W, float(data_range),
                                       _lib.ptr(per_image), _lib.ptr(sums), _lib.ptr(raw), _lib.stream_ptr()))
    return per_image, sums, raw


def _f32(v: float) -> float:
    return float(torch.tensor(v, dtype=torch.float32).item())


def calculate_psnr(pred: torch.Tensor, target: torch.Tensor, data_range: float = 6.0) -> float:
    """10*log10(data_range^2 / mse) over ALL elements passed; inf if mse == 0 (metrics.py:39-46)."""
    _, _, raw = fused_metrics(pred, target, data_range)
    mse = _f32(raw[:, 0].sum().item() / pred.numel())
    if mse == 0:
        return float("inf")
    return 10 * math.log10((data_range ** 2) / mse)


def calculate_ssim(pred: torch.Tensor, target: torch.Tensor, data_range: float = 6.0, window_size: int = 11) -> float:
    """Mean of the Gaussian-window SSIM map (metrics.py:49-121)."""
    if window_size != 11:
        raise _lib.MsdError("calculate_ssim: only window_size=11 is built into libmsd_b200")
    _, _, raw = fused_metrics(pred, target, data_range)
    return _f32(raw[:, 1].sum().item() / pred.numel())


def calculate_sam(pred: torch.Tensor, target: torch.Tensor, epsilon: float = 1e-8) -> float:
    """Mean spectral angle in degrees over all pixels (metrics.py:124-172)."""
    if epsilon != 1e-8:
        raise _lib.MsdError("calculate_sam: only epsilon=1e-8 is built into libmsd_b200")
    if pred.ndim not in (3, 4):
        raise _lib.MsdError("calculate_sam: expected [B,C,H,W] or [C,H,W]")
    _, _, raw = fused_metrics(pred, target)
    npix = pred.numel() // pred.shape[-3]
    return _f32(raw[:, 2].sum().item() / npix)


def calculate_rmse(pred: torch.Tensor, target: torch.Tensor) -> float:
    """sqrt(mse) (metrics.py:175-191)."""
    _, _, raw = fused_metrics(pred, target)
    return math.sqrt(_f32(raw[:, 0].sum().item() / pred.numel()))


def calculate_mae(pred: torch.Tensor, target: torch.Tensor) -> float:
    """mean |pred - target| (metrics.py:194-209)."""
    _, _, raw = fused_metrics(pred, target)
    return _f32(raw[:, 3].sum().item() / pred.numel())


class MetricsTracker:
    """Per-image PSNR / SSIM / SAM / RMSE running sums (metrics.py:212-287); means of per-image values."""

    def __init__(self, data_range: float = 6.0):
        self.data_range = data_range
        self.reset()

    def reset(self):
        self._acc = None        # fp64 [5] on the device: sum psnr, sum ssim, sum sam, sum rmse, count
        self._host = [0.0, 0.0, 0.0, 0.0, 0.0]
        self._dirty = False
        self._pending = None    # event of an in-flight prefetch()

    @torch.no_grad()
    def update(self, pred: torch.Tensor, target: torch.Tensor):
        """pred, target: [B, C, H, W]; no host synchronisation."""
        if pred.shape[0] == 0:                  # empty batch: the reference's per-sample loop (metrics.py:247-256) adds nothing
            return
        _, sums, _ = fused_metrics(pred, target, self.data_range)
        if self._acc is None:
            self._acc = sums
        else:
            if self._acc.device != sums.device:
                sums = sums.to(self._acc.device)
            self._acc += sums
        self._dirty = True
        self._pending = None

    def all_reduce(self, group=None):
        """Sum the running sums over all ranks (NCCL; the only collective of the evaluation sweep)."""
        import torch.distributed as dist
        if self._acc is None:
            raise _lib.MsdError("MetricsTracker.all_reduce: nothing accumulated on this rank")
        dist.all_reduce(self._acc, op=dist.ReduceOp.SUM, group=group)
        self._dirty = True
        self._pending = None      # a prefetch() issued before the collective holds the LOCAL sums: it must not be read afterwards

    def prefetch(self):
        """Start the device -> host copy of the running sums on the current stream (pinned buffer + event).  A later compute()
        / *_sum then waits for that 40-byte copy only, not for whatever the stream was given afterwards: an evaluation loop
        can enqueue batch i+1 before it reads the result of batch i."""
        if self._acc is None or not self._dirty:
            return self
        if getattr(self, "_pinned", None) is None:
            self._pinned = torch.empty(5, dtype=torch.float64, pin_memory=True)
        with torch.cuda.device(self._acc.device):
            self._pinned.copy_(self._acc, non_blocking=True)
            self._pending = torch.cuda.Event()
            self._pending.record()
        return self

    def _sync(self):
        pending = getattr(self, "_pending", None)
        if pending is not None:
            pending.synchronize()
            self._host = self._pinned.tolist()
            self._pending = None
            self._dirty = False
        elif self._dirty:
            self._host = self._acc.cpu().tolist()
            self._dirty = False
        return self._host

    psnr_sum = property(lambda self: self._sync()[0])
    ssim_sum = property(lambda self: self._sync()[1])
    sam_sum = property(lambda self: self._sync()[2])
    rmse_sum = property(lambda self: self._sync()[3])
    count = property(lambda self: int(self._sync()[4]))

    def compute(self) -> dict:
        s = self._sync()
        n = int(s[4])
        if n == 0:
            return {"psnr": 0.0, "ssim": 0.0, "sam": 0.0, "rmse": 0.0}
        return {"psnr": s[0] / n, "ssim": s[1] / n, "sam": s[2] / n, "rmse": s[3] / n}

    def __repr__(self) -> str:
        m = self.compute()
        return f"PSNR: {m['psnr']:.2f} dB, SSIM: {m['ssim']:.4f}, SAM: {m['sam']:.2f}°, RMSE: {m['rmse']:.4f}"
