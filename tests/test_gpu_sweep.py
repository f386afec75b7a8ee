"""GPU tests of the callers either side of the hot path: evaluate_model, the batch-sharded
evaluation sweep (world-size invariance) and the NCCL all-reduce of the tracker's running sums."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import recipes
from msd_b200.data import AddSensorNoise
from msd_b200.models import SatMAERestoration
from msd_b200.sweep import evaluation_sweep
from msd_b200.training import MetricsTracker, evaluate_model
from oracle import metrics as ometrics

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    torch.manual_seed(42)
    return SatMAERestoration(recipes.base_config(), block_layout="fallback").eval().cuda()


def test_evaluate_model_vs_oracle_tracker(model):
    """evaluate_model (reference src/training/eval.py:12-70) over a loader with a ragged last batch: the means it returns are the
    oracle's per-image PSNR / SSIM / SAM / RMSE means of the SAME restored tiles (BASELINE bounds 0.05 dB / 1e-4 / 0.01 deg), the
    tracker is reset on reuse, and the numbers do not depend on how the tiles are batched."""
    clean, noisy = recipes.eurosat_tiles(8, seed=21, noise=0.05)
    loader = [(clean[0:3], noisy[0:3]), (clean[3:6], noisy[3:6]), (clean[6:8], noisy[6:8])]      # CPU tensors, as a DataLoader yields
    got = evaluate_model(model, loader, torch.device("cuda"), verbose=False)
    with torch.no_grad():
        restored = torch.cat([model(n.cuda()).cpu() for _, n in loader])
    otr = ometrics.MetricsTracker()
    otr.update(restored.numpy(), clean.numpy())
    want = otr.compute()
    assert set(got) == {"psnr", "ssim", "sam", "rmse"}
    assert abs(got["psnr"] - want["psnr"]) < 0.05 and abs(got["ssim"] - want["ssim"]) < 1e-4
    assert abs(got["sam"] - want["sam"]) < 0.01 and abs(got["rmse"] - want["rmse"]) < 1e-5 * max(1.0, want["rmse"])
    tracker = MetricsTracker()
    tracker.update(torch.zeros(2, 13, 192, 192, device="cuda"), torch.ones(2, 13, 192, 192, device="cuda"))   # stale content
    again = evaluate_model(model, [(clean, noisy)], "cuda", metrics_tracker=tracker, verbose=False)          # one batch of 8
    assert tracker.count == 8
    for k in got:       # per-image metrics of per-image-independent restorations: batching changes nothing but fp64 sum order
        assert abs(again[k] - got[k]) <= 1e-9 * max(1.0, abs(got[k])), k


def test_evaluation_sweep_is_world_size_invariant(model):
    """Config C3 in miniature: 5 global batches (the last ragged).  Batch k is generated and corrupted from seed + k whoever runs
    it, so the per-rank running sums of a simulated 8-rank (and 2-, 3-rank) sweep add up to the single-rank sums -- the metric
    all-reduce is then a plain fp64 sum of five numbers."""
    noise = AddSensorNoise(*recipes.NOISE_CONFIGS["base"])
    n_tiles, batch = 37, 8
    one = MetricsTracker()
    res1 = evaluation_sweep(model, noise, n_tiles, batch, rank=0, world=1, tracker=one)
    acc1 = one._acc.cpu().numpy().copy()
    assert int(acc1[4]) == n_tiles and np.isfinite(acc1).all()
    for world in (2, 3, 8):
        total = np.zeros(5)
        for r in range(world):
            t = MetricsTracker()
            evaluation_sweep(model, noise, n_tiles, batch, rank=r, world=world, tracker=t, all_reduce=False)
            if t._acc is not None:
                total += t._acc.cpu().numpy()
        assert int(total[4]) == n_tiles
        assert np.allclose(total, acc1, rtol=1e-13, atol=0), (world, total, acc1)
    # and a second single-rank run reproduces the first bit for bit (seeded tiles, seeded corruption, deterministic kernels)
    two = MetricsTracker()
    res2 = evaluation_sweep(model, noise, n_tiles, batch, rank=0, world=1, tracker=two)
    assert res1 == res2 and np.array_equal(two._acc.cpu().numpy(), acc1)


def test_full_27000_tile_sweep_is_world_size_invariant(model):
    """BASELINE config C3 at its full size: 27,000 EuroSAT-shaped tiles (106 global batches of 256, the last ragged) on one rank
    vs the eight shards of an 8-rank run executed one after the other -- the per-shard fp64 sums add up to the single-rank sums
    (the NCCL all-reduce of the real run is then a sum of five numbers per rank), every metric finite and inside its range."""
    noise = AddSensorNoise(*recipes.NOISE_CONFIGS["base"])
    n_tiles, batch = 27000, 256
    one = MetricsTracker()
    res = evaluation_sweep(model, noise, n_tiles, batch, rank=0, world=1, tracker=one)
    acc1 = one._acc.cpu().numpy().copy()
    assert int(acc1[4]) == n_tiles and np.isfinite(acc1).all()
    assert 0 < res["psnr"] < 60 and -1 <= res["ssim"] <= 1 and 0 <= res["sam"] <= 180 and res["rmse"] > 0
    total = np.zeros(5)
    counts = []
    for r in range(8):
        t = MetricsTracker()
        evaluation_sweep(model, noise, n_tiles, batch, rank=r, world=8, tracker=t, all_reduce=False)
        part = t._acc.cpu().numpy()
        counts.append(int(part[4]))
        total += part
    assert sum(counts) == n_tiles and max(counts) - min(counts) <= 256 + 256      # contiguous shards of 13-14 batches
    assert np.allclose(total, acc1, rtol=1e-12, atol=0), (total, acc1)


def test_tracker_prefetch_then_all_reduce_reads_the_reduced_sums(model):
    """prefetch() -> all_reduce() -> compute() must not return the pre-collective local sums.  Single-process group."""
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29541", rank=0, world_size=1, device_id=torch.device("cuda", 0))
        created = True
    else:
        created = False
    try:
        t = MetricsTracker()
        a, b = recipes.iid_tiles(2, seed=8)
        t.update(b.cuda(), a.cuda())
        t.prefetch()
        t.all_reduce()
        t._acc += 1.0                      # stand-in for what other ranks would have added: only a post-collective read sees it
        t._dirty = True
        assert t.count == 3
    finally:
        if created:
            dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL all-reduce across ranks)")
def test_sweep_cli_two_ranks_over_nccl_matches_one_rank(tmp_path):
    """python -m msd_b200.sweep under torchrun with 2 ranks: MetricsTracker.all_reduce over NCCL gives the single-rank metrics."""
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    args = ["--tiles", "40", "--batch", "8"]
    one = subprocess.run([sys.executable, "-m", "msd_b200.sweep", *args], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert one.returncode == 0, one.stderr[-2000:]
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29547", "-m", "msd_b200.sweep", *args], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert two.returncode == 0, two.stderr[-2000:]
    r1 = json.loads([l for l in one.stdout.splitlines() if l.startswith("{")][-1])
    r2 = json.loads([l for l in two.stdout.splitlines() if l.startswith("{")][-1])
    assert r2["world"] == 2 and r1["world"] == 1
    for k in ("psnr", "ssim", "sam", "rmse"):
        assert abs(r1[k] - r2[k]) <= 1e-12 * max(1.0, abs(r1[k])), (k, r1[k], r2[k])
