"""Full-scene restoration (BASELINE config C5): a Sentinel-2 scene [13, H, W] (10980 x 10980 at 10 m) is cut into 192 x 192
tiles on a regular grid with overlap, every tile goes through the restoration model, and the restored tiles are blended
back with separable linear-ramp weights.  The reference restores single tiles only: the tiling / blend defined
here (and restated in oracle/scene.py) is the semantics.

Multi-GPU: the tile grid is split into contiguous row bands, one per rank.  A rank also restores every tile row above its
band that still covers one of its scene rows (halo recompute: normally one row, two when the edge-anchored last row sits
closer than `overlap` to its neighbour), so every scene row it owns gathers only tiles it holds and NO collective is needed
on the data path; each rank returns its own band of scene rows.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import torch

from . import _lib


def grid_origins(extent: int, tile: int, overlap: int) -> List[int]:
    """Tile origins along one axis: stride tile-overlap, last tile anchored at extent-tile."""
    if extent < tile:
        raise ValueError(f"scene extent {extent} smaller than the tile {tile}")
    if not (0 <= overlap <= tile // 2):
        raise ValueError(f"overlap must be in [0, {tile // 2}], got {overlap}")
    stride = tile - overlap
    n = -(-(extent - tile) // stride) + 1
    return [min(k * stride, extent - tile) for k in range(n)]


@dataclass
class BandPlan:
    tile_rows: Tuple[int, int]     # [first, last) grid rows this rank restores (own band + one halo row above)
    own_rows: Tuple[int, int]      # [first, last) grid rows it owns
    pixel_rows: Tuple[int, int]    # [row0, row1) scene rows it writes


def band_plan(gy: List[int], scene_h: int, rank: int, world: int, tile: int = 192) -> BandPlan:
    """Contiguous split of the grid rows; pixel rows are cut at the origin of each band's first tile row.  The halo comes from the
    geometry: every grid row above the band whose tiles reach into the band's first scene row (gy[r] + tile > row0)."""
    ny = len(gy)
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(ny, world)
    first = rank * base + min(rank, extra)
    last = first + base + (1 if rank < extra else 0)
    if first == last:                                   # more ranks than tile rows: nothing to do
        return BandPlan((first, first), (first, last), (0, 0))
    row0 = 0 if first == 0 else gy[first]
    nxt = last
    row1 = scene_h if nxt >= ny else gy[nxt]
    halo = first
    while halo > 0 and gy[halo - 1] + tile > row0:
        halo -= 1
    return BandPlan((halo, last), (first, last), (row0, row1))


@torch.no_grad()
def restore_scene(model, scene: torch.Tensor, overlap: int = 32, batch: int = 256, rank: int = 0, world: int = 1,
                  tile: int = 192, transform=None):
    """scene: fp32 CUDA [C,H,W] (already in the model's normalisation).  Returns (band, (row0, row1)): the restored scene rows
    this rank owns, fp32 [C, row1-row0, W].  `transform` (e.g. AddSensorNoise) is applied to each batch of tiles before the
    model, as in the evaluation path."""
    if not scene.is_cuda or scene.ndim != 3 or scene.dtype != torch.float32:
        raise _lib.MsdError("restore_scene: fp32 CUDA tensor [C,H,W] required")
    scene = scene.contiguous()
    C_, H, W = scene.shape
    lib = _lib.load()
    _lib.require_device(scene.device)
    gy, gx = grid_origins(H, tile, overlap), grid_origins(W, tile, overlap)
    ny, nx = len(gy), len(gx)
    plan = band_plan(gy, H, rank, world, tile)
    r0, r1 = plan.tile_rows
    row0, row1 = plan.pixel_rows
    if row1 <= row0:
        return scene.new_empty(C_, 0, W), (0, 0)
    dev = scene.device
    n_tiles = (r1 - r0) * nx
    ty0 = torch.tensor([gy[r] for r in range(r0, r1) for _ in range(nx)], dtype=torch.int32, device=dev)
    tx0 = torch.tensor([gx[k] for _ in range(r0, r1) for k in range(nx)], dtype=torch.int32, device=dev)
    index = torch.full((ny * nx,), -1, dtype=torch.int32)
    index[r0 * nx:r1 * nx] = torch.arange(n_tiles, dtype=torch.int32)
    index = index.to(dev)
    gy_d, gx_d = torch.tensor(gy, dtype=torch.int32, device=dev), torch.tensor(gx, dtype=torch.int32, device=dev)
    restored = torch.empty(n_tiles, C_, tile, tile, dtype=torch.float32, device=dev)
    model.eval()
    with torch.cuda.device(dev):
        s = _lib.stream_ptr()
        for t0 in range(0, n_tiles, batch):
            n = min(batch, n_tiles - t0)
            tiles = torch.empty(n, C_, tile, tile, dtype=torch.float32, device=dev)
            _lib.check(lib.msd_tile_extract_f32(_lib.ptr(scene), _lib.ptr(ty0[t0:t0 + n]), _lib.ptr(tx0[t0:t0 + n]), _lib.ptr(tiles), n,
                                                C_, H, W, tile, s))
            if transform is not None:
                tiles = transform(tiles)
            restored[t0:t0 + n] = model(tiles)
        out = torch.empty(C_, row1 - row0, W, dtype=torch.float32, device=dev)
        # the blend kernel addresses the output as a full [C,H,W] scene: hand it a pointer shifted so that row `row0` is ours
        ramp = max(overlap, 1)
        full = torch.empty(C_, H, W, dtype=torch.float32, device=dev) if (row0, row1) != (0, H) else None
        target = full if full is not None else out
        _lib.check(lib.msd_tile_blend_f32(_lib.ptr(restored), _lib.ptr(gy_d), _lib.ptr(gx_d), _lib.ptr(index), ny, nx, _lib.ptr(target),
                                          C_, H, W, tile, ramp, row0, row1, s))
        if full is not None:
            out.copy_(full[:, row0:row1])
    return out, (row0, row1)


def scene_to_dn(band: torch.Tensor, rows: Tuple[int, int], scene_h: int, rank: int = 0, world: int = 1, host_out: torch.Tensor | None = None):
    """Writer side of C5: a rank's restored band (fp32, model normalisation) -> Sentinel-2 digital numbers (uint16,
    `msd_export_u16`), gathered to rank 0 over NCCL (one padded gather; the bands are the only data that moves) and copied into
    `host_out` (pinned uint16 [C, scene_h, W]) when given.  Returns the full uint16 scene on rank 0 (device tensor), None elsewhere."""
    from .data import export_tiles
    C_, _, W = band.shape
    dev = band.device
    row0, row1 = rows
    pad = (-band.numel()) % 4
    flat = band.reshape(-1)
    if pad:
        flat = torch.cat([flat, flat.new_zeros(pad)])
    dn = export_tiles(flat)[:band.numel()].view(C_, row1 - row0, W) if band.numel() else torch.empty(C_, 0, W, dtype=torch.uint16, device=dev)
    if world == 1:
        full = dn
    else:
        import torch.distributed as dist
        spans = [None] * world
        dist.all_gather_object(spans, (row0, row1))
        hmax = max(b - a for a, b in spans)
        # NCCL has no 16-bit integer type: the bands travel as bytes
        mine = torch.zeros(C_, hmax, W * 2, dtype=torch.uint8, device=dev)
        mine[:, :row1 - row0] = dn.view(torch.uint8).view(C_, row1 - row0, W * 2)
        parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, parts, dst=0)
        if rank != 0:
            return None
        full = torch.empty(C_, scene_h, W * 2, dtype=torch.uint8, device=dev)
        for (a, b), part in zip(spans, parts):
            if b > a:
                full[:, a:b] = part[:, :b - a]
        full = full.view(torch.uint16)
    if host_out is not None:
        host_out.copy_(full, non_blocking=True)
    return full
