"""Oracle: the Philox4x32-10 stream PyTorch's CUDA generator produces, restated in numpy.

Algorithm source (third-party, not vendored in the reference): Random123's philox4x32-10 as implemented by cuRAND
(/usr/local/cuda/include/curand_philox4x32_x.h) and consumed by ATen's distribution kernels
(torch/include/ATen/native/cuda/DistributionTemplates.h:50-90).  The reference reaches it through
``torch.rand(C, device=x.device) < p`` and ``torch.randn_like`` (src/data/transforms.py:91,96,102,116).

Pinned by: the Random123 known-answer vectors (tests/test_oracle_philox.py) on the CPU; on the GPU box the tests
compare libmsd's stream bit-for-bit with torch.rand / torch.randn on the device.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr: uint32 [...,4], key: uint32 [...,2] -> uint32 [...,4]."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0 = key[..., 0].astype(np.uint32).copy()
    k1 = key[..., 1].astype(np.uint32).copy()
    for r in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c = [(hi1 ^ c[1] ^ k0.astype(np.uint64)) & MASK32, lo1, (hi0 ^ c[3] ^ k1.astype(np.uint64)) & MASK32, lo0]
        if r < 9:
            with np.errstate(over="ignore"):
                k0 = (k0 + W0).astype(np.uint32)
                k1 = (k1 + W1).astype(np.uint32)
    return np.stack([x.astype(np.uint32) for x in c], axis=-1)


def torch_cuda_draws(seed: int, offset: int, numel: int, torch_grid: int) -> np.ndarray:
    """Raw uint32 word feeding element e of a `numel`-element CUDA distribution call.

    Thread t = e % T (T = 256*grid) in its call r = e // (4T) reads component (e // T) % 4 of
    philox(counter = (offset/4 + r, subsequence = t), key = seed).
    """
    assert offset % 4 == 0
    T = 256 * torch_grid
    e = np.arange(numel, dtype=np.uint64)
    t = e % np.uint64(T)
    r = e // np.uint64(4 * T)
    comp = ((e // np.uint64(T)) % np.uint64(4)).astype(np.int64)
    lo = np.uint64(offset // 4) + r
    ctr = np.stack([lo & MASK32, lo >> np.uint64(32), t & MASK32, t >> np.uint64(32)], axis=-1).astype(np.uint32)
    key = np.broadcast_to(np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32), (numel, 2))
    out = philox4x32_10(ctr, key)
    return out[np.arange(numel), comp]


def curand_uniform(x: np.ndarray) -> np.ndarray:
    """_curand_uniform: (float)x * 2^-32 + 2^-33 -- x is rounded to fp32 first, then one fused multiply-add -> (0, 1]."""
    return (x.astype(np.float32).astype(np.float64) * 2.0 ** -32 + 2.0 ** -33).astype(np.float32)


def torch_cuda_rand(seed: int, offset: int, numel: int, torch_grid: int) -> np.ndarray:
    """torch.rand(numel, device='cuda'): curand uniform in (0,1], with 1.0 mapped to 0.0."""
    u = curand_uniform(torch_cuda_draws(seed, offset, numel, torch_grid))
    return np.where(u == np.float32(1.0), np.float32(0.0), u)


def offset_increment(numel: int, torch_grid: int) -> int:
    return ((numel - 1) // (256 * torch_grid * 4) + 1) * 4


def dead_band_masks(seed: int, offset0: int, B: int, C: int, p: float, torch_grid: int = 1) -> np.ndarray:
    """[B,C] bool: sample b draws torch.rand(C) at offset0 + b*inc and compares with float32(p) (transforms.py:101-106)."""
    inc = offset_increment(C, torch_grid)
    return np.stack([torch_cuda_rand(seed, offset0 + b * inc, C, torch_grid) < np.float32(p) for b in range(B)])
