"""GPU parity of the data formats either side of the path: raw-tile ingest against the reference-generated
golden vectors and the oracle, and full-scene tiling / blending (BASELINE config C5) against the oracle's brute-force blend."""
from pathlib import Path

import numpy as np
import pytest
import torch

from msd_b200 import _lib
from msd_b200.data import ingest_tiles
from msd_b200.scene import band_plan, grid_origins, restore_scene
from oracle import scene as oscene

pytestmark = pytest.mark.gpu


def test_ingest_matches_reference_golden_and_oracle():
    g = np.load(Path(__file__).parent / "golden" / "ingest.npz")
    dn = torch.from_numpy(g["dn"].astype(np.int32)).repeat(1, 13, 1, 1)
    want = torch.from_numpy(g["clean"]).repeat(1, 13, 1, 1)
    got = ingest_tiles(dn.cuda(), 192).cpu()
    assert (got - want).abs().max().item() <= 2e-6          # fp32 bilinear: same weights, association may differ by an ulp
    # EuroSAT-like 16-bit digital numbers, all three scaling branches, a non-square source
    gen = torch.Generator().manual_seed(3)
    cases = [torch.randint(0, 5300, (5, 13, 64, 64), generator=gen), torch.randint(0, 200, (2, 13, 64, 64), generator=gen),
             torch.randint(0, 2, (2, 13, 64, 64), generator=gen), torch.randint(0, 12000, (3, 4, 50, 70), generator=gen)]
    cases[1][:, :, 0, 0] = 9       # max 9 -> not above 10 -> the /255 branch needs max > 1.5
    cases[1] = cases[1].clamp(max=9)
    for dn in cases:
        got = ingest_tiles(dn.to(torch.int32).cuda(), 192).cpu()
        assert (got - oscene.ingest(dn, 192)).abs().max().item() <= 2e-6
    with pytest.raises(_lib.MsdError):
        ingest_tiles(torch.zeros(1, 13, 8, 8, dtype=torch.int32))           # CPU tensor: no fallback


class _Affine(torch.nn.Module):
    """stand-in for the restoration model: a per-pixel affine map, so the blend result is known in closed form"""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.ones(1))

    def forward(self, x):
        return 0.5 * x + 1.0


@pytest.mark.parametrize("H,W,overlap,world", [(384, 384, 0, 1), (500, 700, 32, 1), (500, 700, 32, 3), (1000, 433, 96, 2), (192, 192, 0, 1),
                                                  (500, 433, 96, 5)])   # last: one rank per tile row, edge row 20 px below its neighbour
def test_scene_tiling_and_blend_vs_oracle(H, W, overlap, world):
    gen = torch.Generator().manual_seed(H + W)
    scene = torch.randn(5, H, W, generator=gen)
    model = _Affine().cuda()
    gy, gx = grid_origins(H, 192, overlap), grid_origins(W, 192, overlap)
    tiles = np.stack([scene[:, y:y + 192, x:x + 192].numpy() for y in gy for x in gx])
    want = oscene.blend(0.5 * tiles + 1.0, gy, gx, H, W, overlap)
    out = np.zeros_like(want)
    seen = np.zeros(H, bool)
    for rank in range(world):
        band, (r0, r1) = restore_scene(model, scene.cuda(), overlap=overlap, batch=7, rank=rank, world=world)
        out[:, r0:r1] = band.cpu().numpy()
        assert not seen[r0:r1].any()
        seen[r0:r1] = True
    assert seen.all()
    assert np.abs(out - want).max() <= 2e-6
    # an affine map commutes with the (normalised) blend: the scene comes back exactly transformed
    assert np.abs(out - (0.5 * scene.numpy() + 1.0)).max() <= 2e-6


def test_scene_restoration_with_the_real_model_is_shard_invariant():
    import recipes
    from msd_b200.models import SatMAERestoration
    torch.manual_seed(42)
    model = SatMAERestoration(recipes.base_config(), block_layout="fallback").eval().cuda()
    scene = torch.randn(13, 400, 520, generator=torch.Generator().manual_seed(1)).cuda()
    full, rows = restore_scene(model, scene, overlap=32, batch=4)
    assert rows == (0, 400) and full.shape == (13, 400, 520) and torch.isfinite(full).all()
    parts = [restore_scene(model, scene, overlap=32, batch=4, rank=r, world=2) for r in range(2)]
    glued = torch.cat([p[0] for p in parts], dim=1)
    assert parts[0][1][1] == parts[1][1][0] and torch.equal(glued, full)     # same tiles, same gather order: bit-identical


def test_export_is_the_inverse_of_the_ingest_scaling():
    from msd_b200.data import export_tiles
    gen = torch.Generator().manual_seed(4)
    x = torch.rand(3, 13, 192, 192, generator=gen) * 5 - 2.5            # beyond [-2, 2] on both sides: clamped
    got = export_tiles(x.cuda()).cpu().to(torch.int32)
    want = torch.round((x * 0.25 + 0.5).clamp(0, 1) * 10000).to(torch.int32)
    assert (got - want).abs().max().item() <= 1                          # rounding ties at fp32 resolution
    assert (got != want).float().mean().item() < 1e-3
    dn = torch.randint(0, 10001, (2, 13, 192, 192), generator=gen)
    back = export_tiles(ingest_tiles(dn.to(torch.int32).cuda(), 192)).cpu().to(torch.int32)   # same size: no resampling
    assert (back - dn).abs().max().item() <= 1


def test_scene_writer_exports_digital_numbers_into_a_host_buffer():
    """scene_to_dn (the C5 writer): restored band -> msd_export_u16 -> pinned host buffer; single rank = the whole scene, equal to
    the export formula of SURVEY 8f N4 applied to the blended band."""
    from msd_b200.scene import scene_to_dn
    gen = torch.Generator().manual_seed(9)
    scene = (torch.rand(13, 230, 260, generator=gen) * 5 - 2.5).cuda()
    model = _Affine().cuda()
    band, rows = restore_scene(model, scene, overlap=32, batch=5)
    host = torch.zeros(13, 230, 260, dtype=torch.uint16).pin_memory()
    full = scene_to_dn(band, rows, 230, host_out=host)
    torch.cuda.synchronize()
    want = torch.round((band.cpu() * 0.25 + 0.5).clamp(0, 1) * 10000).to(torch.int32)
    got = host.to(torch.int32)
    assert torch.equal(full.cpu().to(torch.int32), got)
    assert (got - want).abs().max().item() <= 1 and (got != want).float().mean().item() < 1e-3
