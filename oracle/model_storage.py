"""Oracle, second form: the training step of SatMAERestoration with the CUDA path's STORAGE ROUNDINGS written into the graph.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): nothing in the product imports it.

`oracle/model.py::restoration_forward_train` is the reference's training forward in fp32 (src/models/satmae_restoration.py:81-110,
encoder.py:154-176, blocks.py:66-85, 113-123, decoder.py:76-118 under autograd).  The library computes the same graph with bf16
tensor-core operands and bf16 activations / inter-layer gradients in HBM, which puts it 1 % (head) ... 10 % (encoder) relative RMS
away from fp32 autograd -- the same distance torch's own bf16 autocast shows.  To separate "bf16 storage" from "a wrong kernel", this
module states WHERE the library rounds -- every tensor it keeps in bf16, forward and backward -- as identity-like autograd nodes in
the same fp32 torch graph:

    rb(x)  value rounded to bf16 in the forward AND the gradient rounded to bf16 in the backward   (a bf16 tensor in HBM whose
           gradient is also kept in bf16: activations, raw conv outputs, qkv, ...)
    rf(x)  value rounded, gradient passed through        (bf16 operand copies of the fp32 master weights; P of the attention)
    rg(x)  value passed through, gradient rounded        (a gradient that is rounded where no forward tensor is stored: the
           pre-activation gradient inside BatchNorm+GELU, the bf16 copy of the residual-stream gradient fed to the GEMMs, dS)

All arithmetic between those points stays fp32 (the kernels accumulate in fp32; BatchNorm statistics come from the rounded tensor
in fp64).  What the emulation does not model: summation order, the tanh-form / fp16 evaluation of GELU and its derivative in the
epilogues (|error| <= 1e-3), fp32 atomics order of split-K.  tests/test_gpu_train.py compares the CUDA step against this graph at a
tolerance several times tighter than against plain fp32 autograd.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .model import _block_tensors, _bn


class _Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fwd: bool, bwd: bool):
        ctx.bwd = bwd
        return x.to(torch.bfloat16).to(x.dtype) if fwd else x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return (g.to(torch.bfloat16).to(g.dtype) if ctx.bwd else g), None, None


def rb(x):
    return _Round.apply(x, True, True)


def rf(x):
    return _Round.apply(x, True, False)


def rg(x):
    return _Round.apply(x, False, True)


def encoder_forward(sd: dict, x: torch.Tensor, num_heads: int = 12, ln_eps_block: float = 1e-5, ln_eps_outer: float = 1e-5):
    w = sd["encoder.patch_embed.proj.weight"]
    P, D = w.shape[-1], w.shape[0]
    # patch operand and weights in bf16, fp32 output; the gradient that reaches the GEMM is the bf16 copy of the stream gradient
    t = rg(F.conv2d(x.to(torch.bfloat16).to(x.dtype), rf(w), sd["encoder.patch_embed.proj.bias"], stride=P))
    t = t.flatten(2).transpose(1, 2)
    t = F.layer_norm(t, (D,), sd["encoder.patch_embed.norm.weight"], sd["encoder.patch_embed.norm.bias"], ln_eps_outer)
    t = t + sd["encoder.pos_embed"]                                  # fp32 residual stream, fp32 gradient
    B, N, _ = t.shape
    hd = D // num_heads
    i = 0
    while f"encoder.blocks.{i}.norm1.weight" in sd:
        p = f"encoder.blocks.{i}."
        qkv_w, qkv_b, proj_w, proj_b, fc1_w, fc1_b, fc2_w, fc2_b = _block_tensors(sd, i)
        h = rb(F.layer_norm(t, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], ln_eps_block))
        qkv = rb(F.linear(h, rf(qkv_w), qkv_b)).reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
        s = rg((qkv[0] @ qkv[1].transpose(-2, -1)) * hd ** -0.5)    # dS is a bf16 MMA operand in the backward
        att = rb(rf(torch.softmax(s, dim=-1)) @ qkv[2])              # P is a bf16 MMA operand; dP stays in registers
        t = t + rg(F.linear(att.transpose(1, 2).reshape(B, N, D), rf(proj_w), proj_b))
        h = rb(F.layer_norm(t, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], ln_eps_block))
        hpre = rb(F.linear(h, rf(fc1_w), fc1_b))
        t = t + rg(F.linear(rb(F.gelu(hpre)), rf(fc2_w), fc2_b))
        i += 1
    return rb(F.layer_norm(t, (D,), sd["encoder.norm.weight"], sd["encoder.norm.bias"], ln_eps_outer))


def decoder_forward(sd: dict, feat: torch.Tensor):
    x = feat
    i = 0
    while True:
        p = f"decoder.upsample_blocks.{i}."
        if p + "upsample.0.weight" in sd:
            u = rb(F.conv_transpose2d(x, rf(sd[p + "upsample.0.weight"]), sd[p + "upsample.0.bias"], stride=2))
            x = rb(F.gelu(rg(_bn(u, sd, p + "upsample.1.", training=True))))
        elif p + "block.0.weight" in sd:
            c1 = rb(F.conv2d(x, rf(sd[p + "block.0.weight"]), None, padding=1))
            y = rb(F.gelu(rg(_bn(c1, sd, p + "block.1.", training=True))))
            c2 = rb(F.conv2d(y, rf(sd[p + "block.3.weight"]), None, padding=1))
            x = rb(F.gelu(rg(x + _bn(c2, sd, p + "block.4.", training=True))))
        else:
            break
        i += 1
    c = rb(F.conv2d(x, rf(sd["decoder.head.0.weight"]), sd["decoder.head.0.bias"], padding=1))
    x = rb(F.gelu(rg(_bn(c, sd, "decoder.head.1.", training=True))))
    # the 1x1 head runs on fp32 weights; its weight / bias gradients are formed from the bf16-rounded output gradient
    return rg(F.conv2d(x, sd["decoder.head.3.weight"], sd["decoder.head.3.bias"]))


def restoration_forward_train(sd: dict, x: torch.Tensor, num_heads: int = 12, ln_eps_block: float = 1e-5, ln_eps_outer: float = 1e-5):
    """The training forward of oracle/model.py with the library's bf16 storage points in the graph (see the module docstring);
    running BatchNorm buffers in `sd` are updated in place."""
    tok = encoder_forward(sd, x, num_heads, ln_eps_block, ln_eps_outer)
    B, N, D = tok.shape
    g = int(round(N ** 0.5))
    return decoder_forward(sd, tok.transpose(1, 2).reshape(B, D, g, g))
