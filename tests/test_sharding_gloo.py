"""Multi-rank host logic on the CPU: shard partition of the evaluation sweep and the metric-sum all-reduce
(world_size 2, gloo, 127.0.0.1).  The CUDA arithmetic is covered by the -m gpu tests; here the tracker's device
accumulator is filled by hand so that the collective path itself runs without a GPU."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from msd_b200.sweep import batch_sizes, shard_batches
from msd_b200.training.metrics import MetricsTracker


@pytest.mark.parametrize("n_tiles,batch,world", [(27000, 250, 1), (27000, 250, 8), (27000, 256, 8), (1000, 256, 3), (5, 8, 4), (0, 8, 2)])
def test_shards_cover_every_batch_exactly_once(n_tiles, batch, world):
    sizes = batch_sizes(n_tiles, batch)
    assert sum(sizes) == n_tiles and all(0 < s <= batch for s in sizes)
    seen = []
    counts = []
    for r in range(world):
        lo, hi = shard_batches(len(sizes), r, world)
        assert 0 <= lo <= hi <= len(sizes)
        seen.extend(range(lo, hi))
        counts.append(hi - lo)
    assert seen == list(range(len(sizes)))              # contiguous, ordered, no gaps, no overlap
    assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        shard_batches(len(sizes), world, world)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t = MetricsTracker()
    # what fused_metrics would have accumulated on this rank: sum psnr, sum ssim, sum sam, sum rmse, count
    t._acc = torch.tensor([30.0 * (rank + 1), 0.9 * (rank + 1), 2.0 * (rank + 1), 0.05 * (rank + 1), float(rank + 1)], dtype=torch.float64)
    t._dirty = True
    t.all_reduce()
    out[rank] = (t.compute(), t.count)
    dist.destroy_process_group()


def test_metric_sums_allreduce_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29517, out), nprocs=world, join=True)
    for rank in range(world):
        m, count = out[rank]
        assert count == 3                                   # 1 + 2 images
        assert m["psnr"] == pytest.approx(90.0 / 3) and m["ssim"] == pytest.approx(2.7 / 3)
        assert m["sam"] == pytest.approx(6.0 / 3) and m["rmse"] == pytest.approx(0.15 / 3)


def test_launcher_rebinds_reference_names():
    from pathlib import Path
    ref = Path("/root/reference")
    if not (ref / "src").exists():
        pytest.skip("reference tree not present on this machine")
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); from pathlib import Path; from msd_b200.run import rebind; "
            "done = rebind(Path('/root/reference')); import src.training.eval as e, src.data.dataloader as d, src.utils; "
            "import msd_b200.training.metrics as M, msd_b200.data as D; "
            "assert e.MetricsTracker is M.MetricsTracker and d.AddSensorNoise is D.AddSensorNoise and len(done) >= 20") % str(Path(__file__).resolve().parent.parent)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


def _grad_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from msd_b200.training.distributed import allreduce_gradients
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(s)) for s in ((7, 5), (1000,), (3, 3, 3, 3), (2,))]
    params[1].requires_grad_(False)                           # a frozen parameter: no gradient on any rank
    for i, p in enumerate(params):
        if p.requires_grad:
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    n = allreduce_gradients(params, bucket_mb=1)
    out[rank] = (n, [None if p.grad is None else p.grad.clone() for p in params])
    dist.destroy_process_group()


def test_gradient_allreduce_world2():
    """C4 data-parallel step: gradients are averaged in place, frozen parameters are skipped, shapes survive the bucketing."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_grad_worker, args=(world, 29519, out), nprocs=world, join=True)
    for rank in range(world):
        n, grads = out[rank]
        assert n >= 1 and grads[1] is None
        for i in (0, 2, 3):
            assert grads[i].shape == [(7, 5), (1000,), (3, 3, 3, 3), (2,)][i]
            assert torch.allclose(grads[i], torch.full_like(grads[i], 1.5 * (i + 1)))


def _group_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from msd_b200.training.distributed import reduce_gradient_groups
    # an arena of three gradient groups with an alignment gap and an absent (empty) group in between
    arena = torch.arange(40, dtype=torch.float32) * (rank + 1)
    bounds, ids = [(0, 13), (16, 16), (16, 40)], [0, 3, 7]
    n = reduce_gradient_groups(arena, bounds, ids, None, None)
    out[rank] = (n, arena.clone())
    dist.destroy_process_group()


def test_gradient_group_reduction_world2():
    """the in-place per-group averaging SatMAERestoration.enable_data_parallel() runs on its gradient arena (on CUDA: on a
    communication stream behind each group's event; here the same code on CPU tensors over gloo)"""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_group_worker, args=(world, 29523, out), nprocs=world, join=True)
    base = torch.arange(40, dtype=torch.float32)
    for rank in range(world):
        n, arena = out[rank]
        assert n == 2                                           # the empty group issues no collective
        assert torch.allclose(arena[:13], base[:13] * 1.5) and torch.allclose(arena[16:], base[16:] * 1.5)
        assert torch.equal(arena[13:16], base[13:16] * (rank + 1))          # the alignment gap is nobody's gradient: untouched
