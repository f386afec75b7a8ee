"""Fused optimizer step for the training loop: global-norm gradient clipping + AdamW in two multi-tensor
kernels of libmsd_b200 (csrc/optim.cu).  Same hyper-parameters, parameter groups and state layout (``step``, ``exp_avg``,
``exp_avg_sq``) as ``torch.optim.AdamW`` (reference scripts/train.py:144-160), so optimizer state dicts interchange; the
clip is the arithmetic of ``torch.nn.utils.clip_grad_norm_`` (src/training/trainer.py:188-193) folded into the update."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from .. import _lib


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 max_grad_norm: Optional[float] = None):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("FusedAdamW: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = max_grad_norm
        self.last_grad_sqnorm = None          # device fp64 [1]: squared global gradient norm of the last step (before clipping)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        # pass 1: validate everything BEFORE any state is touched (a raise must not leave half-advanced step counters)
        todo, hyper, dev = [], None, None
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise _lib.MsdError("FusedAdamW: contiguous fp32 CUDA parameters required (libmsd_b200 has no CPU path)")
                if p.grad.dtype != torch.float32 or p.grad.device != p.device:
                    raise _lib.MsdError("FusedAdamW: fp32 gradients on the parameter's device required")
                h = (b1, b2, group["eps"])
                if hyper is not None and (h != hyper or p.device != dev):
                    raise _lib.MsdError("FusedAdamW: all parameters must share betas / eps and one device")
                hyper, dev = h, p.device
                todo.append((p, group))
        # pass 2: per-parameter state and step (torch.optim.AdamW keeps one step per parameter: a parameter that had no gradient
        # for a while -- blocks unfrozen mid-run, a resumed checkpoint -- simply carries a smaller count)
        entries, keep = [], []
        for p, group in todo:
            st = self.state[p]
            if len(st) == 0:
                st["step"] = torch.zeros((), dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["step"] += 1
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            keep.append(g)
            e = _lib.AdamWTensor()
            e.param, e.grad, e.exp_avg, e.exp_avg_sq = p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            e.numel, e.lr, e.weight_decay, e.step = p.numel(), float(group["lr"]), float(group["weight_decay"]), int(st["step"])
            entries.append(e)
        if not entries:
            return loss
        _lib.require_device(dev)
        table = (_lib.AdamWTensor * len(entries))(*entries)
        with torch.cuda.device(dev):
            s = _lib.stream_ptr()
            clip = float(self.max_grad_norm) if self.max_grad_norm else 0.0
            sq = None
            if clip > 0:
                sq = torch.empty(1, dtype=torch.float64, device=dev)
                _lib.check(lib.msd_grad_sqnorm_f32(table, len(entries), _lib.ptr(sq), s))
                self.last_grad_sqnorm = sq
            _lib.check(lib.msd_adamw_step_f32(table, len(entries), hyper[0], hyper[1], hyper[2], 1, clip, _lib.ptr(sq), s))
        # The kernel wrote the parameters through raw pointers: tell autograd (and every cache keyed on `_version`, e.g. the packed
        # bf16 / transposed copies of SatMAERestoration) that they changed.
        torch.autograd.graph.increment_version([p for p, _ in todo])
        return loss
