"""Oracle (test infrastructure) for the data formats either side of the path:
  * ingest: EuroSATMultiSpectral.__getitem__'s preprocessing (reference src/data/dataset.py:160-186) restated with torch CPU ops
    -- pinned against the reference class itself by tests/golden/ingest.npz (oracle/make_golden.py);
  * scene tiling / blend (BASELINE config C5): the reference has no scene path, so this brute-force numpy
    scatter-accumulate IS the specification the CUDA gather kernel is checked against ("parity unpinned" by construction)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def ingest(dn: torch.Tensor, target_size: int = 192) -> torch.Tensor:
    """dn integer [B,C,h,w] -> fp32 [B,C,S,S]; one sample at a time, exactly the reference's statement order."""
    out = []
    for img in dn:
        img = img.float()
        if img.max() > 10:                       # dataset.py:160-165
            img = img / 10000.0
        elif img.max() > 1.5:
            img = img / 255.0
        if img.shape[1] != target_size or img.shape[2] != target_size:   # dataset.py:168-175
            img = F.interpolate(img.unsqueeze(0), size=(target_size, target_size), mode="bilinear", align_corners=False).squeeze(0)
        img = torch.clamp(img, 0, 1)             # dataset.py:182
        out.append((img - 0.5) / 0.25)           # dataset.py:186
    return torch.stack(out)


def grid_origins(extent: int, tile: int, overlap: int):
    stride = tile - overlap
    n = -(-(extent - tile) // stride) + 1
    return [min(k * stride, extent - tile) for k in range(n)]


def ramp(tile: int, r: int) -> np.ndarray:
    i = np.arange(tile)
    return np.minimum(np.minimum(i + 1, tile - i), r).astype(np.float32) / np.float32(r)


def blend(tiles: np.ndarray, gy, gx, H: int, W: int, overlap: int) -> np.ndarray:
    """tiles [ny*nx, C, S, S] (row-major grid) -> [C,H,W]: weighted scatter-accumulate in float64, then normalise."""
    S = tiles.shape[-1]
    w1 = ramp(S, max(overlap, 1)).astype(np.float64)
    w2 = np.outer(w1, w1)
    acc = np.zeros((tiles.shape[1], H, W), np.float64)
    ws = np.zeros((H, W), np.float64)
    t = 0
    for y0 in gy:
        for x0 in gx:
            acc[:, y0:y0 + S, x0:x0 + S] += w2 * tiles[t].astype(np.float64)
            ws[y0:y0 + S, x0:x0 + S] += w2
            t += 1
    return (acc / ws).astype(np.float32)
