"""The export graph of SatMAERestoration: the forward pass spelled in stock PyTorch ops over the module's OWN parameters, for
tracers only.  The reference's scripts/export.py:45-66 and src/deployment/export_onnx.py:40-69 build the
model on the CPU and hand ``model.forward`` to ``torch.onnx.export`` with a CPU dummy input; a tracer can only record ATen ops,
so the drop-in needs a definition of the same function in those ops -- this file.  It is NOT an execution path of the product:
``SatMAERestoration.forward`` takes it only while a tracer is recording (``torch.jit.is_tracing()`` /
``torch.onnx.is_in_onnx_export()``) or inside an explicit ``export_mode()`` block (what ``msd_b200.run`` opens around the
reference's export script, whose ``compare_pytorch_onnx_outputs`` evaluates the module on the CPU as the ONNX ground truth);
any other non-CUDA call raises ``MsdError`` as before.

Follows the reference op for op: encoder.py:41-52 (patch embedding), :309 (+pos_embed), :154-176 / timm Block (pre-norm
blocks, softmax(q k^T / sqrt(d)) v, Linear-GELU-Linear), :322 (final norm), satmae_restoration.py:128-133 (tokens -> map),
blocks.py:74-85, 113-123 and decoder.py:76-81, 101-118 (decoder).
"""
from __future__ import annotations

import contextlib
import threading

import torch
import torch.nn.functional as F

from .blocks import ResidualBlock, UpsampleBlock

_state = threading.local()


@contextlib.contextmanager
def export_mode():
    """Inside this block a non-CUDA ``SatMAERestoration.forward`` evaluates the export graph instead of raising."""
    prev = getattr(_state, "on", False)
    _state.on = True
    try:
        yield
    finally:
        _state.on = prev


def export_graph_requested() -> bool:
    if getattr(_state, "on", False) or torch.jit.is_tracing():
        return True
    try:
        return bool(torch.onnx.is_in_onnx_export())
    except Exception:  # pragma: no cover
        return False


def _block(blk, t: torch.Tensor, num_heads: int) -> torch.Tensor:
    w = blk.tensors()
    B, N, D = t.shape
    hd = D // num_heads
    h = w["ln1"](t)
    qkv = F.linear(h, w["qkv_w"], w["qkv_b"]).reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    att = torch.softmax(torch.matmul(qkv[0], qkv[1].transpose(-2, -1)) * (hd ** -0.5), dim=-1)
    ctx = torch.matmul(att, qkv[2]).transpose(1, 2).reshape(B, N, D)
    t = t + F.linear(ctx, w["proj_w"], w["proj_b"])
    h = w["fc1"](w["ln2"](t))
    return t + w["fc2"](F.gelu(h))


def export_forward(model, x: torch.Tensor) -> torch.Tensor:
    enc, dec = model.encoder, model.decoder
    pe = enc.patch_embed
    t = pe.norm(pe.proj(x).flatten(2).transpose(1, 2)) + enc.pos_embed
    for blk in enc.blocks:
        t = _block(blk, t, enc.num_heads)
    t = enc.norm(t)
    B, _, D = t.shape
    f = t.transpose(1, 2).reshape(B, D, model.grid_size, model.grid_size)
    for mod in dec.upsample_blocks:
        if isinstance(mod, UpsampleBlock):
            f = mod.upsample(f)
        elif isinstance(mod, ResidualBlock):
            f = mod.activation(f + mod.block(f))
        else:  # pragma: no cover
            raise TypeError(f"unexpected decoder module {type(mod).__name__}")
    return dec.head(f)
