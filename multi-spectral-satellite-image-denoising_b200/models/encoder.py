"""Parameter tree of the SatMAE ViT-Base encoder -- same names, shapes and init order as the reference's
``src.models.encoder`` (src/models/encoder.py:21-324) so checkpoints load with ``strict=True`` and a seeded
construction draws the same random numbers.  These modules only HOLD parameters: the arithmetic lives in
libmsd_b200 and is driven by ``SatMAERestoration.forward``.

Two state-dict key layouts exist for the transformer blocks:
  "timm"     blocks.N.{norm1,norm2}.*, attn.qkv.*, attn.proj.*, mlp.fc1.*, mlp.fc2.*   LayerNorm eps 1e-6
             (what trained checkpoints and the SatMAE weights contain; encoder.py:126-134)
  "fallback" blocks.N.attn.in_proj_{weight,bias}, attn.out_proj.*, mlp.0.*, mlp.2.*    LayerNorm eps 1e-5
             (the reference's CustomBlock when timm is absent; encoder.py:154-176)
The tensors are layout-identical, so either loads into either via a key rename (``convert_block_keys``).
"""
from __future__ import annotations

import logging
import re
from typing import List, Optional

import torch
import torch.nn as nn

logger = logging.getLogger(__name__)

_TIMM_TO_FALLBACK = {
    "attn.qkv.weight": "attn.in_proj_weight", "attn.qkv.bias": "attn.in_proj_bias",
    "attn.proj.weight": "attn.out_proj.weight", "attn.proj.bias": "attn.out_proj.bias",
    "mlp.fc1.weight": "mlp.0.weight", "mlp.fc1.bias": "mlp.0.bias",
    "mlp.fc2.weight": "mlp.2.weight", "mlp.fc2.bias": "mlp.2.bias",
}
_FALLBACK_TO_TIMM = {v: k for k, v in _TIMM_TO_FALLBACK.items()}
_BLOCK_RE = re.compile(r"^(.*blocks\.\d+\.)(.+)$")


def convert_block_keys(state_dict: dict, target: str) -> dict:
    """Rename transformer-block keys of ``state_dict`` to the ``target`` layout ("timm" | "fallback")."""
    table = _TIMM_TO_FALLBACK if target == "fallback" else _FALLBACK_TO_TIMM
    out = {}
    for k, v in state_dict.items():
        m = _BLOCK_RE.match(k)
        if m and m.group(2) in table:
            k = m.group(1) + table[m.group(2)]
        out[k] = v
    return out


def default_block_layout() -> str:
    """The reference uses timm's Block when timm imports and its CustomBlock otherwise (encoder.py:120-124)."""
    try:
        import timm  # noqa: F401
        return "timm"
    except ImportError:
        return "fallback"


class PatchEmbed(nn.Module):
    """Conv2d(C, D, k=P, s=P) + LayerNorm(D) holder (encoder.py:21-52)."""

    def __init__(self, in_channels: int = 13, embed_dim: int = 768, patch_size: int = 16):
        super().__init__()
        self.patch_size = patch_size
        self.proj = nn.Conv2d(in_channels, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = nn.LayerNorm(embed_dim)


class _FallbackBlock(nn.Module):
    """Holder with the key names of the reference's CustomBlock (encoder.py:157-170)."""
    ln_eps = 1e-5

    def __init__(self, dim: int, num_heads: int, mlp_ratio: float):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = nn.MultiheadAttention(dim, num_heads, batch_first=True)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = nn.Sequential(nn.Linear(dim, int(dim * mlp_ratio)), nn.GELU(), nn.Linear(int(dim * mlp_ratio), dim))

    def tensors(self):
        return dict(ln1=self.norm1, ln2=self.norm2, qkv_w=self.attn.in_proj_weight, qkv_b=self.attn.in_proj_bias,
                    proj_w=self.attn.out_proj.weight, proj_b=self.attn.out_proj.bias, fc1=self.mlp[0], fc2=self.mlp[2])


class _TimmAttention(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _TimmMlp(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _TimmBlock(nn.Module):
    """Holder with the key names of timm.models.vision_transformer.Block(qkv_bias=True, LayerNorm eps 1e-6)."""
    ln_eps = 1e-6

    def __init__(self, dim: int, num_heads: int, mlp_ratio: float):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _TimmAttention(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _TimmMlp(dim, int(dim * mlp_ratio))

    def tensors(self):
        return dict(ln1=self.norm1, ln2=self.norm2, qkv_w=self.attn.qkv.weight, qkv_b=self.attn.qkv.bias,
                    proj_w=self.attn.proj.weight, proj_b=self.attn.proj.bias, fc1=self.mlp.fc1, fc2=self.mlp.fc2)


class SatMAEEncoder(nn.Module):
    """ViT-Base encoder parameters: patch_embed, pos_embed [1,N,D], blocks, norm (encoder.py:55-152)."""

    def __init__(self, in_channels: int = 13, image_size: int = 224, patch_size: int = 16, embed_dim: int = 768,
                 depth: int = 12, num_heads: int = 12, mlp_ratio: float = 4.0, pretrained_path: Optional[str] = None,
                 freeze_layers: Optional[List[int]] = None, use_checkpointing: bool = True,
                 block_layout: Optional[str] = None):
        super().__init__()
        self.image_size, self.patch_size, self.embed_dim, self.depth = image_size, patch_size, embed_dim, depth
        self.num_heads, self.mlp_ratio = num_heads, mlp_ratio
        self.block_layout = block_layout or default_block_layout()
        if self.block_layout not in ("timm", "fallback"):
            raise ValueError(f"block_layout must be 'timm' or 'fallback', got {self.block_layout!r}")
        self.patch_embed = PatchEmbed(in_channels=in_channels, embed_dim=embed_dim, patch_size=patch_size)
        num_patches = (image_size // patch_size) ** 2
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches, embed_dim))
        block_cls = _TimmBlock if self.block_layout == "timm" else _FallbackBlock
        self.blocks = nn.ModuleList([block_cls(embed_dim, num_heads, mlp_ratio) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        if pretrained_path:
            from .pretrained import load_pretrained_encoder
            load_pretrained_encoder(self, pretrained_path, in_channels)
        if freeze_layers:
            self._freeze_layers(freeze_layers)
        self.use_checkpointing = use_checkpointing
        if use_checkpointing:
            logger.info("Gradient checkpointing requested (recompute policy is decided by the kernel library)")
        self._register_load_state_dict_pre_hook(self._rename_block_keys)

    def _rename_block_keys(self, state_dict, prefix, *_):
        renamed = convert_block_keys({k: v for k, v in state_dict.items() if k.startswith(prefix)}, self.block_layout)
        for k in [k for k in state_dict if k.startswith(prefix)]:
            del state_dict[k]
        state_dict.update(renamed)

    @property
    def ln_eps_block(self) -> float:
        return self.blocks[0].ln_eps if len(self.blocks) else 1e-6

    def _freeze_layers(self, freeze_indices: List[int]):
        """requires_grad=False on the listed blocks only; patch_embed / pos_embed / norm stay trainable (encoder.py:270-293)."""
        frozen = 0
        for idx in freeze_indices:
            if idx < len(self.blocks):
                for p in self.blocks[idx].parameters():
                    p.requires_grad = False
                    frozen += p.numel()
        total = sum(p.numel() for p in self.parameters())
        logger.info("Frozen %d layers: %s / %s params (%.1f%%)", len(freeze_indices), f"{frozen:,}", f"{total:,}",
                    100 * frozen / total)

    def get_num_params(self) -> dict:
        total = sum(p.numel() for p in self.parameters())
        trainable = sum(p.numel() for p in self.parameters() if p.requires_grad)
        return {"total": total, "trainable": trainable, "frozen": total - trainable}

    def forward(self, x):  # pragma: no cover - holder only
        raise RuntimeError("SatMAEEncoder holds parameters only; call SatMAERestoration.forward (libmsd_b200 runs "
                           "encoder and decoder as one kernel sequence)")
