"""Mirror of the reference's ``src.training`` for the hot path (src/training/__init__.py:7-18).  ``Trainer`` is the
reference's own orchestration code and is not re-implemented (caller, out of scope)."""
from .losses import CombinedLoss, SSIMLoss
from .metrics import (MetricsTracker, calculate_mae, calculate_psnr, calculate_rmse, calculate_sam, calculate_ssim)
from .eval import evaluate_model
from .optim import FusedAdamW

__all__ = ["CombinedLoss", "SSIMLoss", "MetricsTracker", "calculate_psnr", "calculate_ssim", "calculate_sam",
           "calculate_rmse", "calculate_mae", "evaluate_model", "FusedAdamW"]
