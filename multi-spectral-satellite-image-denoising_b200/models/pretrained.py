"""Checkpoint surgery for SatMAE weights (reference: src/models/encoder.py:178-268): accept a
``{'model'|'state_dict'|flat}`` checkpoint, widen a 3-channel patch embedding to the target band count, drop or
interpolate the position embedding, rename timm/fallback block keys, and load what fits (non-strict)."""
from __future__ import annotations

import logging
import math

import torch
import torch.nn.functional as F

logger = logging.getLogger(__name__)


def adapt_satmae_state(state: dict, model_state: dict, target_channels: int) -> dict:
    state = dict(state)
    w = state.get("patch_embed.proj.weight")
    if w is not None and w.shape[1] != target_channels:
        # bands 0-2 keep the RGB filters, every further band starts from the first filter
        wide = w.new_zeros(w.shape[0], target_channels, w.shape[2], w.shape[3])
        n = min(3, w.shape[1], target_channels)
        wide[:, :n] = w[:, :n]
        wide[:, n:] = w[:, :1]
        state["patch_embed.proj.weight"] = wide
    pe, want = state.get("pos_embed"), model_state.get("pos_embed")
    if pe is not None and want is not None and pe.shape[1] != want.shape[1]:
        if pe.shape[1] == want.shape[1] + 1:
            state["pos_embed"] = pe[:, 1:, :]  # drop the CLS slot
        else:
            n_src = pe.shape[1]
            has_cls = int(math.sqrt(n_src)) ** 2 != n_src
            grid_src = int(math.sqrt(n_src - 1 if has_cls else n_src))
            body = pe[:, 1:, :] if has_cls else pe
            grid_dst = int(math.sqrt(want.shape[1]))
            body = body.reshape(1, grid_src, grid_src, -1).permute(0, 3, 1, 2)
            body = F.interpolate(body, size=(grid_dst, grid_dst), mode="bicubic", align_corners=False)
            state["pos_embed"] = body.flatten(2).transpose(1, 2)
    return {k: v for k, v in state.items() if k in model_state and model_state[k].shape == v.shape}


def load_pretrained_encoder(encoder, path: str, target_channels: int) -> None:
    from .encoder import convert_block_keys
    ckpt = torch.load(str(path), map_location="cpu", weights_only=False)
    state = ckpt.get("model", ckpt.get("state_dict", ckpt)) if isinstance(ckpt, dict) else ckpt
    state = convert_block_keys(state, encoder.block_layout)
    fitted = adapt_satmae_state(state, encoder.state_dict(), target_channels)
    encoder.load_state_dict(fitted, strict=False)
    logger.info("Loaded %d / %d pretrained tensors from %s", len(fitted), len(encoder.state_dict()), path)
