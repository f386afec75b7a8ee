"""Corruption simulator -- drop-in for the reference's ``src.data.transforms.AddSensorNoise``
(src/data/transforms.py:20-163): same constructor, same call semantics (3-D or 4-D input, same shape / dtype /
device out), same ``__repr__``.  The work is one fused sm_100a kernel (csrc/corrupt.cu) that regenerates PyTorch's
CUDA Philox stream in place, so on a CUDA tensor the result is bit-identical to the reference's and the default
CUDA generator is left at the same offset the reference would leave it at.
"""
from __future__ import annotations

import ctypes as C
import logging

import torch
import torch.nn as nn

from .. import _lib

logger = logging.getLogger(__name__)


def _torch_grid(numel: int, props) -> int:
    """Grid PyTorch launches for a `numel`-element distribution kernel (DistributionTemplates.h:50-61)."""
    blocks_per_sm = props.max_threads_per_multi_processor // 256
    return max(1, min(props.multi_processor_count * blocks_per_sm, (numel + 255) // 256))


class AddSensorNoise(nn.Module):
    """Gaussian + speckle + dead-band + wavelength-dependent thermal noise (+ optional striping), clamp to [-3, 3]."""

    def __init__(self, gaussian_sigma: float = 0.05, speckle_sigma: float = 0.03, dead_band_prob: float = 0.25,
                 thermal_scale: float = 0.01, enable_striping: bool = False, stripe_prob: float = 0.1):
        super().__init__()
        self.gaussian_sigma = gaussian_sigma
        self.speckle_sigma = speckle_sigma
        self.dead_band_prob = dead_band_prob
        self.thermal_scale = thermal_scale
        self.enable_striping = enable_striping
        self.stripe_prob = stripe_prob
        self._thermal_w = {}
        self.last_dead_mask = None  # [B, C] bool of the last call (exposed for the mask-parity tests)
        logger.debug("Initialized AddSensorNoise: gaussian=%.3f, dead_bands=%.2f", gaussian_sigma, dead_band_prob)

    def _weights(self, C_: int, device) -> torch.Tensor:
        key = (C_, device)
        w = self._thermal_w.get(key)
        if w is None:
            # built exactly as the reference builds it (transforms.py:112-114), once per (C, device)
            w = torch.linspace(1.0, 2.0, C_, device=device).contiguous()
            self._thermal_w[key] = w
        return w

    @torch.no_grad()
    def forward(self, x: torch.Tensor, patch_out: torch.Tensor | None = None) -> torch.Tensor:
        """x: clean image [C,H,W] or [B,C,H,W] on a CUDA device -> corrupted image of the same shape.

        patch_out: optional preallocated bf16 [B*(H/16)*(W/16), C*256] that receives the patch-embed operand
        of the corrupted image (lets the restoration model skip its patchify pass).
        """
        if not x.is_cuda:
            # The reference calls the transform per sample on a CPU tensor inside Dataset.__getitem__ (src/data/dataset.py:188-190).
            # To keep that call site working unchanged the sample is staged through the current CUDA device: H2D copy, the same
            # kernel, D2H copy.  This is NOT a CPU implementation -- without a CUDA device (or in a forked DataLoader worker, where
            # CUDA cannot be initialised) it raises, as every other entry point does.
            if not torch.cuda.is_available():
                raise _lib.MsdError("AddSensorNoise: a CUDA device is required (libmsd_b200 has no CPU path); got " + str(x.device))
            return self.forward(x.cuda(), patch_out).to(x.device)
        if x.dtype != torch.float32:
            raise _lib.MsdError(f"AddSensorNoise: float32 required, got {x.dtype}")
        is_batched = x.ndim == 4
        if not is_batched:
            x = x.unsqueeze(0)
        if x.ndim != 4:
            raise _lib.MsdError(f"AddSensorNoise: expected [C,H,W] or [B,C,H,W], got {tuple(x.shape)}")
        x = x.contiguous()
        B, C_, H, W = x.shape
        if x.numel() == 0:                      # empty batch: the reference's clone / randn_like chain returns an empty tensor
            self.last_dead_mask = torch.zeros(B, C_, dtype=torch.bool, device=x.device) if self.dead_band_prob > 0 else None
            return x.clone() if is_batched else x.clone().squeeze(0)
        lib = _lib.load()
        _lib.require_device(x.device)
        with torch.cuda.device(x.device):
            props = torch.cuda.get_device_properties(x.device)
            gen = torch.cuda.default_generators[x.device.index if x.device.index is not None else torch.cuda.current_device()]
            seed, offset = gen.initial_seed(), gen.get_offset()
            out = torch.empty_like(x)
            mask = torch.zeros(B * C_, dtype=torch.uint8, device=x.device) if self.dead_band_prob > 0 else None
            # striping coin flip comes from the CPU generator, as in the reference (transforms.py:124)
            stripe = bool(self.enable_striping and torch.rand(1).item() < self.stripe_prob)
            stripe_ws = torch.empty(B * C_ * W, dtype=torch.float32, device=x.device) if stripe else None
            consumed = C.c_uint64(0)
            _lib.check(lib.msd_corrupt_f32(
                _lib.ptr(x), _lib.ptr(out), _lib.ptr(self._weights(C_, x.device)) if self.thermal_scale > 0 else None,
                _lib.ptr(mask), _lib.ptr(stripe_ws), _lib.ptr(patch_out), B, C_, H, W,
                float(self.gaussian_sigma), float(self.speckle_sigma), float(self.dead_band_prob),
                float(self.thermal_scale), int(stripe), seed, offset,
                _torch_grid(x.numel(), props), _torch_grid(C_, props), _torch_grid(B * C_ * W, props),
                C.byref(consumed), _lib.stream_ptr()))
            gen.set_offset(offset + consumed.value)
        self.last_dead_mask = mask.view(B, C_).bool() if mask is not None else None
        if not is_batched:
            out = out.squeeze(0)
        return out

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(gaussian_sigma={self.gaussian_sigma:.3f}, "
                f"speckle_sigma={self.speckle_sigma:.3f}, dead_band_prob={self.dead_band_prob:.2f}, "
                f"thermal_scale={self.thermal_scale:.3f})")
