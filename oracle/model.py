"""Oracle: SatMAERestoration.forward (reference src/models/satmae_restoration.py:81-135, encoder.py:41-52, 154-176,
295-324, blocks.py:74-85, 113-123, decoder.py:101-118) as ONE functional torch fp32 routine over a state dict.

Accepts either transformer-block key layout.  Pinned by tests/golden/model_*.npz against the reference
run with its fallback CustomBlock (LayerNorm eps 1e-5).  The timm-Block variant (eps 1e-6) follows timm's public
Block/Attention/Mlp definition and is PARITY UNPINNED: timm is not installed and nothing in the reference pins it.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _block_tensors(sd: dict, i: int):
    p = f"encoder.blocks.{i}."
    if p + "attn.in_proj_weight" in sd:
        return (sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"], sd[p + "attn.out_proj.weight"],
                sd[p + "attn.out_proj.bias"], sd[p + "mlp.0.weight"], sd[p + "mlp.0.bias"], sd[p + "mlp.2.weight"],
                sd[p + "mlp.2.bias"])
    return (sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"], sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"],
            sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"], sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])


def _bn(x, sd, p, eps=1e-5, training=False):
    """eval: running statistics; training: batch statistics, running buffers updated in place with momentum 0.1
    (nn.BatchNorm2d defaults, reference blocks.py:69)."""
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"], training,
                        0.1 if training else 0.0, eps)


def encoder_forward(sd: dict, x: torch.Tensor, num_heads: int = 12, ln_eps_block: float = 1e-5, ln_eps_outer: float = 1e-5):
    w = sd["encoder.patch_embed.proj.weight"]
    P, D = w.shape[-1], w.shape[0]
    t = F.conv2d(x, w, sd["encoder.patch_embed.proj.bias"], stride=P)             # encoder.py:48
    t = t.flatten(2).transpose(1, 2)                                                # encoder.py:50
    t = F.layer_norm(t, (D,), sd["encoder.patch_embed.norm.weight"], sd["encoder.patch_embed.norm.bias"], ln_eps_outer)
    t = t + sd["encoder.pos_embed"]                                                 # encoder.py:309
    B, N, _ = t.shape
    hd = D // num_heads
    i = 0
    while f"encoder.blocks.{i}.norm1.weight" in sd:
        p = f"encoder.blocks.{i}."
        qkv_w, qkv_b, proj_w, proj_b, fc1_w, fc1_b, fc2_w, fc2_b = _block_tensors(sd, i)
        h = F.layer_norm(t, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], ln_eps_block)
        qkv = F.linear(h, qkv_w, qkv_b).reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
        att = torch.softmax((qkv[0] @ qkv[1].transpose(-2, -1)) * hd ** -0.5, dim=-1) @ qkv[2]
        t = t + F.linear(att.transpose(1, 2).reshape(B, N, D), proj_w, proj_b)
        h = F.layer_norm(t, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], ln_eps_block)
        t = t + F.linear(F.gelu(F.linear(h, fc1_w, fc1_b)), fc2_w, fc2_b)
        i += 1
    return F.layer_norm(t, (D,), sd["encoder.norm.weight"], sd["encoder.norm.bias"], ln_eps_outer)  # encoder.py:322


def decoder_forward(sd: dict, feat: torch.Tensor, training: bool = False):
    x = feat
    i = 0
    while True:
        p = f"decoder.upsample_blocks.{i}."
        if p + "upsample.0.weight" in sd:                                           # blocks.py:113-120
            x = F.conv_transpose2d(x, sd[p + "upsample.0.weight"], sd[p + "upsample.0.bias"], stride=2)
            x = F.gelu(_bn(x, sd, p + "upsample.1.", training=training))
        elif p + "block.0.weight" in sd:                                            # blocks.py:74-85
            y = F.gelu(_bn(F.conv2d(x, sd[p + "block.0.weight"], None, padding=1), sd, p + "block.1.", training=training))
            y = _bn(F.conv2d(y, sd[p + "block.3.weight"], None, padding=1), sd, p + "block.4.", training=training)
            x = F.gelu(x + y)
        else:
            break
        i += 1
    x = F.gelu(_bn(F.conv2d(x, sd["decoder.head.0.weight"], sd["decoder.head.0.bias"], padding=1), sd, "decoder.head.1.",
                  training=training))
    return F.conv2d(x, sd["decoder.head.3.weight"], sd["decoder.head.3.bias"])      # decoder.py:80


@torch.no_grad()
def restoration_forward(sd: dict, x: torch.Tensor, num_heads: int = 12, ln_eps_block: float = 1e-5,
                        ln_eps_outer: float = 1e-5, return_tokens: bool = False):
    tok = encoder_forward(sd, x, num_heads, ln_eps_block, ln_eps_outer)
    B, N, D = tok.shape
    g = int(round(N ** 0.5))
    feat = tok.transpose(1, 2).reshape(B, D, g, g)                                  # satmae_restoration.py:128-133
    y = decoder_forward(sd, feat)
    return (y, tok) if return_tokens else y


def restoration_forward_train(sd: dict, x: torch.Tensor, num_heads: int = 12, ln_eps_block: float = 1e-5,
                              ln_eps_outer: float = 1e-5):
    """Training-mode forward (BatchNorm batch statistics; running buffers in `sd` are updated in place), differentiable:
    what `self.model(noisy)` computes inside Trainer.train_epoch (reference src/training/trainer.py:141-143) in fp32."""
    tok = encoder_forward(sd, x, num_heads, ln_eps_block, ln_eps_outer)
    B, N, D = tok.shape
    g = int(round(N ** 0.5))
    return decoder_forward(sd, tok.transpose(1, 2).reshape(B, D, g, g), training=True)
