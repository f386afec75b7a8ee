"""Raw-tile ingest on the GPU: the preprocessing ``EuroSATMultiSpectral.__getitem__`` applies per sample on
the CPU (reference src/data/dataset.py:160-186) -- reflectance scaling, bilinear 64 -> 192 resize, clamp, normalisation --
as one fused kernel over a batch of uint16 digital numbers.  Halves the host->device bytes of a real sweep (uint16 64x64
in instead of fp32 192x192) and removes the per-sample CPU work in front of the corruption simulator."""
from __future__ import annotations

import torch

from .. import _lib


def ingest_tiles(dn: torch.Tensor, target_size: int = 192) -> torch.Tensor:
    """dn: uint16 (or int16 / int32 holding 0..65535) CUDA tensor [B,C,h,w] of digital numbers -> fp32 [B,C,S,S] in the
    model's input normalisation ([-2, 2])."""
    if not dn.is_cuda:
        raise _lib.MsdError("ingest_tiles: CUDA tensor required (libmsd_b200 has no CPU path)")
    if dn.ndim != 4:
        raise _lib.MsdError(f"ingest_tiles: expected [B,C,h,w], got {tuple(dn.shape)}")
    if dn.dtype != torch.uint16:
        if dn.dtype not in (torch.int16, torch.int32, torch.int64, torch.uint8):
            raise _lib.MsdError(f"ingest_tiles: integer digital numbers expected, got {dn.dtype}")
        dn = dn.to(torch.int32).clamp_(0, 65535).to(torch.uint16)
    dn = dn.contiguous()
    B, C, h, w = dn.shape
    lib = _lib.load()
    _lib.require_device(dn.device)
    with torch.cuda.device(dn.device):
        out = torch.empty(B, C, target_size, target_size, dtype=torch.float32, device=dn.device)
        scale = torch.empty(B, dtype=torch.float32, device=dn.device)
        _lib.check(lib.msd_ingest_u16_f32(_lib.ptr(dn), _lib.ptr(out), _lib.ptr(scale), B, C, h, w, target_size, _lib.stream_ptr()))
    return out


def export_tiles(x: torch.Tensor) -> torch.Tensor:
    """Restored tiles in the model's normalisation ([-2, 2], fp32 CUDA) -> Sentinel-2 digital numbers (uint16, reflectance x
    10000): the inverse of `ingest_tiles`' scaling, for writing results back in the source product's units."""
    if not x.is_cuda or x.dtype != torch.float32:
        raise _lib.MsdError("export_tiles: fp32 CUDA tensor required (libmsd_b200 has no CPU path)")
    x = x.contiguous()
    if x.numel() % 4:
        raise _lib.MsdError("export_tiles: element count must be a multiple of 4")
    lib = _lib.load()
    _lib.require_device(x.device)
    with torch.cuda.device(x.device):
        out = torch.empty(x.shape, dtype=torch.uint16, device=x.device)
        _lib.check(lib.msd_export_u16(_lib.ptr(x), _lib.ptr(out), x.numel(), _lib.stream_ptr()))
    return out
