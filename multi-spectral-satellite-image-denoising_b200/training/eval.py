"""``evaluate_model`` -- same signature and return value as the reference's ``src.training.eval.evaluate_model``
(src/training/eval.py:12-70), driving the libmsd model and the device-resident MetricsTracker.  ``world`` /
``all_reduce=True`` adds the one collective of the batch-sharded sweep."""
from __future__ import annotations

import logging
from typing import Dict, Optional

import torch

from .metrics import MetricsTracker

logger = logging.getLogger(__name__)


def evaluate_model(model: torch.nn.Module, dataloader, device, metrics_tracker: Optional[MetricsTracker] = None,
                   verbose: bool = True, all_reduce: bool = False) -> Dict[str, float]:
    model.eval()
    model.to(device)
    if metrics_tracker is None:
        metrics_tracker = MetricsTracker()
    else:
        metrics_tracker.reset()
    if verbose:
        logger.info("Evaluating on %d batches...", len(dataloader))
    with torch.no_grad():
        for clean, noisy in dataloader:
            clean = clean.to(device, non_blocking=True)
            noisy = noisy.to(device, non_blocking=True)
            restored = model(noisy)
            metrics_tracker.update(restored, clean)
    if all_reduce:
        metrics_tracker.all_reduce()
    metrics = metrics_tracker.compute()
    if verbose:
        logger.info("Evaluation Results:")
        for k, v in metrics.items():
            if "psnr" in k:
                logger.info("  %s: %.2f dB", k.upper(), v)
            elif "ssim" in k:
                logger.info("  %s: %.4f", k.upper(), v)
            elif "sam" in k:
                logger.info("  %s: %.2f°", k.upper(), v)
            else:
                logger.info("  %s: %.4f", k, v)
    return metrics
