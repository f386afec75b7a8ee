"""Thin wrappers the GPU tests use to call libmsd_b200 building blocks through the C ABI."""
from __future__ import annotations

import torch

from msd_b200 import _lib
from msd_b200.data.transforms import _torch_grid


def dev():
    return torch.device("cuda", 0)


def torch_grid(numel: int) -> int:
    return _torch_grid(numel, torch.cuda.get_device_properties(0))


def linear(a, w, bias, resid=None, act=0, out_f32=False, backend=0):
    lib = _lib.load()
    M, K = a.shape
    N = w.shape[0]
    out = resid.clone() if resid is not None else torch.empty(M, N, device=a.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    _lib.check(lib.msd_linear_bf16(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(out), _lib.ptr(out if resid is not None else None),
                                   M, N, K, act, int(out_f32 or resid is not None), backend, _lib.stream_ptr()))
    return out


def conv3x3(x_halo, w_packed, bias, skip_halo, B, H, W, cin, cout, backend=0, poison=True):
    lib = _lib.load()
    y = torch.full((B, H + 2, W + 2, cout), 7.0 if poison else 0.0, device=x_halo.device, dtype=torch.bfloat16)
    _lib.check(lib.msd_conv3x3_nhwc_bf16(_lib.ptr(x_halo), _lib.ptr(w_packed), _lib.ptr(bias), _lib.ptr(skip_halo), _lib.ptr(y),
                                         B, H, W, cin, cout, backend, _lib.stream_ptr()))
    return y


def convt2x2(x, w_packed, bias, B, Hin, Win, padded, cin, cout, backend=0):
    lib = _lib.load()
    y = torch.full((B, 2 * Hin + 2, 2 * Win + 2, cout), 7.0, device=x.device, dtype=torch.bfloat16)
    _lib.check(lib.msd_convt2x2_nhwc_bf16(_lib.ptr(x), _lib.ptr(w_packed), _lib.ptr(bias), _lib.ptr(y), B, Hin, Win, int(padded),
                                          cin, cout, backend, _lib.stream_ptr()))
    return y
