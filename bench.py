#!/usr/bin/env python
"""bench.py -- images/sec of the restoration hot path (corrupt -> restore -> metrics) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's own CPU path (baseline/_ref, staged by
                                                                     # tools/stage_ref.sh; oracle port if absent), host cores

Workload (BASELINE.json configs[1], "C2"): a step = one pass over a batch of 256 synthetic 13x192x192 tiles per GPU:
AddSensorNoise (base.yaml noise) -> SatMAERestoration forward (ViT-B encoder + CNN decoder, bf16 tensor cores, fp32
accumulate / residual) -> MetricsTracker.update (PSNR / SSIM / SAM / RMSE per image).  Random-init weights
(seed 42, the reference's own constructor order), EuroSAT-shaped synthetic tiles.

  value : whole-job img/s with the clean batch already resident in HBM (device-timed, max over ranks)
  e2e   : the same step through the public API with the clean batch in pinned HOST memory: H2D copy of the
          batch (double-buffered on a copy stream, overlapping the previous step's compute) + one D2H read of
          the 5 fp64 metric sums inside the timed region, every step
  roofline     : the tcgen05 GEMM / implicit-conv kernel family (97% of the FLOPs), timed live per launch with CUDA
                 events on the launching stream during a second pass over the same steps
  cpu_baseline : the reference's CPU path (its own src.* from baseline/_ref, else the oracle port) on a bounded sample (rank 0, N=1)
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for _p in (str(ROOT), str(ROOT / "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import recipes  # noqa: E402

BATCH = 256                       # tiles per GPU per step (config C2)
TILE_ELEMS = 13 * 192 * 192
FWD_GFLOP_PER_IMG = 53.357        # 2*M*N*K over all GEMM/conv layers
METRIC = "13x192x192 images/sec restore+metrics"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc = gpu_index, None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *exc):
        self.rows = []
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 8:
                self.rows.append(f)

    def summary(self):
        if not getattr(self, "rows", None):
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[4 + i].lower().startswith("active") for r in self.rows)]
        power = [float(r[3]) for r in self.rows if r[3].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][2]) if self.rows else None,
                "reasons": reasons, "power_w_max": max(power) if power else None, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------ reference arm
REF_ROOT = ROOT / "baseline" / "_ref"          # staged by tools/stage_ref.sh (git-ignored, travels with the gpurun snapshot)


def reference_available() -> bool:
    return (REF_ROOT / "src" / "models" / "satmae_restoration.py").exists()


def reference_cpu_run(steps: int, warmup: int, sample_b: int = 8):
    """The UNMODIFIED reference (baseline/_ref/src: AddSensorNoise -> SatMAERestoration -> MetricsTracker.update), fp32 eager on
    all host threads, `sample_b` tiles per step; per-stage and end-to-end times.  timm is not installed,
    so the encoder runs the reference's own fallback CustomBlock (src/models/encoder.py:154-176)."""
    import yaml
    from msd_b200.run import _stub_missing
    if str(REF_ROOT) not in sys.path:
        sys.path.insert(0, str(REF_ROOT))
    _stub_missing()                                  # matplotlib / IPython / onnx are imported by src.utils, never used here
    from src.data.transforms import AddSensorNoise as RefNoise
    from src.models import SatMAERestoration as RefModel
    from src.training.metrics import MetricsTracker as RefTracker
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = yaml.safe_load((REF_ROOT / "configs" / "base.yaml").read_text())
    cfg["model"]["encoder"]["pretrained_path"] = None
    torch.manual_seed(42)
    model = RefModel(cfg).eval()
    n = cfg["noise"]
    noise = RefNoise(n["gaussian_sigma"], n["speckle_sigma"], n["dead_band_prob"], n["thermal_noise_scale"])
    clean, _ = recipes.eurosat_tiles(sample_b, seed=3)
    tracker = RefTracker()
    stage = {"corrupt": 0.0, "restore": 0.0, "metrics": 0.0}
    torch.manual_seed(42)

    def one(record):
        t0 = time.perf_counter()
        noisy = noise(clean)
        t1 = time.perf_counter()
        with torch.no_grad():
            restored = model(noisy)
        t2 = time.perf_counter()
        tracker.update(restored, clean)
        t3 = time.perf_counter()
        if record:
            stage["corrupt"] += t1 - t0; stage["restore"] += t2 - t1; stage["metrics"] += t3 - t2

    for _ in range(warmup):
        one(False)
    t0 = time.perf_counter()
    for _ in range(steps):
        one(True)
    dt = time.perf_counter() - t0
    return {"value": sample_b * steps / dt, "ms_per_step": 1e3 * dt / steps, "cores": cores, "threads": torch.get_num_threads(),
            "kind": "reference",
            "stage_img_per_s": {k: sample_b * steps / v for k, v in stage.items() if v > 0},
            "sample": f"{steps} steps x {sample_b} tiles through the reference's own src.data.AddSensorNoise -> src.models.SatMAERestoration "
                      f"(fallback CustomBlock, timm absent) -> src.training.metrics.MetricsTracker.update, torch eager fp32, {warmup} warm-up"}


def cpu_port_run(steps: int, warmup: int, sample_b: int = 8):
    """Fallback when baseline/_ref is not staged: the oracle's op-for-op torch-eager restatement of the same CPU path."""
    from oracle import reference_port as rp
    from msd_b200.models import SatMAERestoration
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(42)
    sd = {k: v.clone() for k, v in SatMAERestoration(recipes.base_config(), block_layout="fallback").state_dict().items()}
    clean, _ = recipes.eurosat_tiles(sample_b, seed=3)
    tracker = rp.MetricsTracker()
    torch.manual_seed(42)
    for _ in range(warmup):
        rp.hot_path_step(sd, clean, recipes.NOISE_CONFIGS["base"], tracker)
    t0 = time.perf_counter()
    for _ in range(steps):
        rp.hot_path_step(sd, clean, recipes.NOISE_CONFIGS["base"], tracker)
    dt = time.perf_counter() - t0
    return {"value": sample_b * steps / dt, "ms_per_step": 1e3 * dt / steps, "cores": cores, "threads": torch.get_num_threads(),
            "kind": "port",
            "sample": f"{steps} steps x {sample_b} tiles (corrupt -> restore -> per-image PSNR/SSIM/SAM/RMSE), oracle port, torch eager fp32, "
                      f"{warmup} warm-up"}


def cpu_baseline_run(steps: int, warmup: int, sample_b: int = 8):
    return reference_cpu_run(steps, warmup, sample_b) if reference_available() else cpu_port_run(steps, warmup, sample_b)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warmup = max(1, min(args.steps, 10)), max(1, min(args.warmup, 3))          # 3 warm-up + 10 timed
    r = cpu_baseline_run(steps, warmup)
    base = {"value": r["value"], "unit": "img/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
    if "stage_img_per_s" in r:
        base["stage_img_per_s"] = r["stage_img_per_s"]
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "img/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2 hot path (corrupt -> SatMAE ViT-B restore -> metrics), bounded CPU sample of 8 tiles/step",
                   "tile": "13x192x192", "weights": "random-init seed 42", "noise": "configs/base.yaml"},
        "cpu_baseline": base,
        "e2e": {"value": r["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ training step (C4)
def train_step_bench(dev, world, steps=5, warmup=2, micro=64):
    """BASELINE config C4: one data-parallel training step per rank on `micro` tiles: corrupt -> training forward (batch-stat
    BatchNorm) -> MSE + 0.1 (1-SSIM) -> backward (blocks 0-5 frozen) -> gradient all-reduce -> clip 1.0 + AdamW.  All of it runs
    in libmsd_b200 (the all-reduce is NCCL through torch.distributed)."""
    import torch.distributed as dist
    from msd_b200.data import AddSensorNoise
    from msd_b200.models import SatMAERestoration
    from msd_b200.training import CombinedLoss, FusedAdamW
    torch.manual_seed(42)
    model = SatMAERestoration(recipes.base_config(), block_layout="fallback").to(dev).train()
    if world > 1:
        model.enable_data_parallel()                     # gradient all-reduce per group, on a side stream, under the backward
    noise = AddSensorNoise(*recipes.NOISE_CONFIGS["base"])
    crit = CombinedLoss(1.0, 0.1)
    opt = FusedAdamW(model.get_optimizer_param_groups(1e-4, 0.05), betas=(0.9, 0.999), max_grad_norm=1.0)
    clean, _ = recipes.eurosat_tiles(32, seed=200 + int(os.environ.get("RANK", "0")))
    clean = clean.repeat(micro // 32 + 1, 1, 1, 1)[:micro].contiguous().to(dev)

    def step():
        out = model(noise(clean))
        loss = crit(out, clean)
        loss.backward()                                  # world > 1: leaves rank-averaged gradients (overlapped NCCL all-reduces)
        opt.step()                                       # global-norm clip (1.0) + AdamW, one fused pass (csrc/optim.cu)
        opt.zero_grad(set_to_none=True)
        return loss

    for _ in range(warmup):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = ms.item() / steps
    del model, opt
    torch.cuda.empty_cache()
    return {"value": world * micro / ms_step * 1e3, "unit": "img/s", "ms_per_step": ms_step, "micro_batch_per_gpu": micro,
            "steps": steps, "algorithmic_gflop_per_img": 146.4, "tflops_per_gpu": 146.4 * micro / ms_step,
            "loss": float(loss.detach()), "frozen_blocks": [0, 1, 2, 3, 4, 5],
            "note": "corrupt + fwd + loss + bwd + clip/AdamW all in libmsd_b200 (bf16 tensor cores, fp32 accumulate / master weights)"}


def loss_bench(dev, hbm, facts, micro=64, iters=10):
    """MSE + 0.1 (1 - SSIM) forward + analytic backward (config C4's loss at micro-batch 64): reads pred and target, writes the
    gradient = 5.75 MB per tile."""
    from msd_b200.training import CombinedLoss
    clean, pred = recipes.eurosat_tiles(32, seed=7)
    clean = clean.repeat(micro // 32, 1, 1, 1).to(dev).contiguous()
    pred = pred.repeat(micro // 32, 1, 1, 1).to(dev).contiguous()
    crit = CombinedLoss(1.0, 0.1)

    def one():
        p = pred.detach().requires_grad_(True)
        crit(p, clean).backward()

    for _ in range(3):
        one()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nbytes = micro * 3 * TILE_ELEMS * 4
    return {"kernel": "loss_fwd_kernel + loss_bwd_kernel: CombinedLoss forward + backward, micro-batch 64 (config C4)", "ms": ms,
            "algorithmic_bytes": nbytes, "unit": "GB/s", "achieved": nbytes / ms / 1e6, "peak": hbm, "frac": nbytes / ms / 1e6 / hbm,
            **facts.get("loss_kernels", {})}


# ------------------------------------------------------------------------------------------------ full scene (C5)
def scene_bench(model, dev, rank, world, size=10980, overlap=32, batch=256):
    """BASELINE config C5: a synthetic 13 x 10980 x 10980 Sentinel-2 scene -> 192 x 192 tiles with overlap 32 (4,761 tiles), tile
    rows sharded over the ranks in contiguous bands (halo rows recomputed, no data-path collective) -> restore -> overlap-blend
    -> msd_export_u16 -> NCCL gather of the uint16 bands to rank 0 -> D2H into pinned host memory (the writer's buffer)."""
    import torch.distributed as dist
    from msd_b200.scene import grid_origins, restore_scene, scene_to_dn
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    scene = torch.empty(13, size, size, device=dev)
    for c in range(13):
        scene[c] = torch.rand(size, size, generator=g, device=dev) * 4 - 2
    host = torch.empty(13, size, size, dtype=torch.uint16, pin_memory=True) if rank == 0 else None
    small = scene[:, :400, :400].contiguous()
    b, r = restore_scene(model, small, overlap=overlap, batch=4, rank=rank, world=world)   # warm-up: kernels, allocator and the
    scene_to_dn(b, r, 400, rank, world)                                                    # communicator's gather channels
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e[0].record()
    band, rows = restore_scene(model, scene, overlap=overlap, batch=batch, rank=rank, world=world)
    e[1].record()
    scene_to_dn(band, rows, size, rank, world, host_out=host)
    e[2].record()
    torch.cuda.synchronize()
    ms = torch.tensor([e[0].elapsed_time(e[2]), e[0].elapsed_time(e[1])], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    n = len(grid_origins(size, 192, overlap)) ** 2
    total_s, restore_s = ms[0].item() / 1e3, ms[1].item() / 1e3
    ok = bool(torch.isfinite(band).all()) and (host is None or int(host[:, ::997, ::991].to(torch.int32).max()) <= 10000)
    del scene, band
    torch.cuda.empty_cache()
    return {"scene": f"13x{size}x{size} fp32 (6.27 GB)", "tile": 192, "overlap": overlap, "tiles": n, "seconds": total_s,
            "restore_blend_seconds": restore_s, "export_gather_d2h_seconds": total_s - restore_s, "tiles_per_s": n / total_s,
            "scene_gb_per_s": 13 * size * size * 4 / 1e9 / total_s, "d2h_bytes": 13 * size * size * 2, "ranks": world,
            "sharding": "contiguous tile-row bands, halo rows recomputed, no data-path collective; uint16 bands gathered to rank 0",
            "ok": ok}


def ncu_facts():
    """Figures that only a profiler can give (DRAM bytes per launch, tensor-pipe / issue-slot activity): read from the newest
    profiles/r*_ncu_metrics.json (written by tools/ncu_extract.py from the .ncu-rep captures), never typed into this file."""
    files = sorted((ROOT / "profiles").glob("r*_ncu_metrics.json"))
    if not files:
        return {}
    d = json.loads(files[-1].read_text())
    d["source"] = f"profiles/{files[-1].name}"
    return d


# ------------------------------------------------------------------------------------------------ CUDA arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="msd", choices=["msd", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--backend", type=int, default=0, help="0 tcgen05 (product), 1 legacy mma.sync (A/B)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the C4 training-step measurement")
    ap.add_argument("--no-scene", action="store_true", help="skip the C5 full-scene measurement")
    ap.add_argument("--profile-out", default=None, help="write the per-launch event profile (json) here")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from msd_b200 import _lib
    from msd_b200.data import AddSensorNoise
    from msd_b200.models import SatMAERestoration
    from msd_b200.training import MetricsTracker

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line.  NCCL writes its version banner (and anything NCCL_DEBUG asks for) to file descriptor 1
    # from C, so descriptor 1 is pointed at stderr for the whole run and the JSON line is written to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- libmsd_b200 has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K, B = args.steps, args.batch

    torch.manual_seed(42)
    model = SatMAERestoration(recipes.base_config(), block_layout="fallback", gemm_backend=args.backend).eval().to(dev)
    noise = AddSensorNoise(*recipes.NOISE_CONFIGS["base"])
    # each rank owns its own shard of tiles (weak scaling: B tiles per GPU), seeded by global shard index
    clean_host, _ = recipes.eurosat_tiles(32, seed=100 + rank)
    clean_host = clean_host.repeat(B // 32 + 1, 1, 1, 1)[:B].contiguous().pin_memory()
    clean_dev = clean_host.to(dev)
    patches = torch.empty(B * 144, 13 * 256, dtype=torch.bfloat16, device=dev)
    torch.manual_seed(1234 + rank)
    tracker = MetricsTracker()

    def step(clean):
        noisy = noise(clean, patch_out=patches)          # fused corruption, also emits the patch-embed operand
        restored = model(noisy, patches=patches)
        tracker.update(restored, clean)
        return restored

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    with torch.no_grad():
        for _ in range(W):
            step(clean_dev)
        # ---------------- value: inputs resident in HBM, metric sums stay on the device, one allreduce at the end
        def resident():
            step(clean_dev)
        tracker.reset()
        with ClockSampler(local) as clk:
            ms_total = timed(lambda: resident(), K)
            if world > 1:
                tracker.all_reduce()
        metrics_dev = tracker.compute()
        # ---------------- roofline_hbm: the bandwidth-bound kernels of the same step, CUDA events on the launching stream
        hev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
        for it in range(K):
            hev[it][0].record()
            noisy_ = noise(clean_dev, patch_out=patches)
            hev[it][1].record()
            restored_ = model(noisy_, patches=patches)
            hev[it][2].record()
            tracker.update(restored_, clean_dev)
            hev[it][3].record()
        torch.cuda.synchronize()
        corrupt_ms = sum(h[0].elapsed_time(h[1]) for h in hev) / K
        metrics_ms = sum(h[2].elapsed_time(h[3]) for h in hev) / K
        # ---------------- e2e: host buffers in, metric sums out, every step.  The H2D copy of step i+1 runs on a
        # copy stream while step i computes (two device buffers); every step's copy is inside the timed region.
        copy_stream = torch.cuda.Stream(device=dev)
        from msd_b200.data import ingest_tiles
        # mode "u16": the reference's real data flow (src/data/dataset.py:160-186) -- raw uint16 64x64 digital numbers in pinned
        # host memory, 27 MB per step; the ingest kernel (scale, bilinear 64 -> 192, clamp, normalise) runs on the device in front
        # of the corruption.  mode "f32": pre-processed fp32 192x192 tiles from the host, 491 MB per step.
        gen = torch.Generator().manual_seed(100 + rank)
        dn_host = torch.randint(0, 5200, (B, 13, 64, 64), generator=gen, dtype=torch.int32).to(torch.uint16).pin_memory()
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]

        def make_e2e(host_src):
            bufs = [torch.empty(host_src.shape, dtype=host_src.dtype, device=dev) for _ in range(2)]

            def prefetch(i):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(freed[i % 2])
                    bufs[i % 2].copy_(host_src, non_blocking=True)
                    ready[i % 2].record(copy_stream)

            trackers = [MetricsTracker(), MetricsTracker()]     # two, alternating: their pinned result buffers are allocated once

            def e2e_run(steps):
                for b in range(2):
                    freed[b].record()
                prefetch(0)
                out, prev = None, None
                for i in range(steps):
                    if i + 1 < steps:
                        prefetch(i + 1)
                    torch.cuda.current_stream().wait_event(ready[i % 2])
                    c = bufs[i % 2]
                    if c.dtype == torch.uint16:
                        c = ingest_tiles(c, 192)                # raw digital numbers -> the model's normalisation, on the device
                    t = trackers[i % 2]
                    t.reset()                                   # step i - 2's result was read during step i - 1
                    noisy = noise(c, patch_out=patches)
                    restored = model(noisy, patches=patches)
                    t.update(restored, c)
                    t.prefetch()                                # D2H of the 5 fp64 sums of THIS step, behind its kernels
                    freed[i % 2].record()
                    if prev is not None:
                        out = prev.compute()                    # read step i-1's result while step i runs (its launches are queued)
                    prev = t
                out = prev.compute()
                return out
            return e2e_run

        e2e_u16, e2e_f32 = make_e2e(dn_host), make_e2e(clean_host)
        e2e_u16(2)
        ms_e2e = timed(lambda: e2e_u16(K), 1)
        e2e_f32(2)
        ms_e2e_f32 = timed(lambda: e2e_f32(K), 1)
        # ---------------- roofline: per-launch CUDA events inside the library, same steps again
        lib = _lib.load()
        _lib.check(lib.msd_model_profile_enable(model._handle, 1))
        cap = 256
        ms_buf, kind_buf, fl_buf = (ctypes.c_float * cap)(), (ctypes.c_int * cap)(), (ctypes.c_double * cap)()
        agg = {}
        per_launch = None
        for it in range(K):
            step(clean_dev)
            torch.cuda.synchronize()
            n = lib.msd_model_profile_read(model._handle, ms_buf, kind_buf, fl_buf, cap)
            rows = [(ms_buf[i], kind_buf[i], fl_buf[i]) for i in range(n)]
            per_launch = rows if per_launch is None else [(a[0] + b[0], a[1], a[2]) for a, b in zip(per_launch, rows)]
        _lib.check(lib.msd_model_profile_enable(model._handle, 0))
        per_launch = [(ms / K, kind, fl) for ms, kind, fl in per_launch]
        for ms, kind, fl in per_launch:
            a = agg.setdefault(kind, [0.0, 0.0, 0])
            a[0] += ms; a[1] += fl; a[2] += 1

    ms_step = ms_total / K
    value = world * B * K / (ms_total / 1e3)
    e2e_value = world * B * K / (ms_e2e / 1e3)
    e2e_f32_value = world * B * K / (ms_e2e_f32 / 1e3)
    facts = ncu_facts()
    pk = peaks()
    hbm = pk["hbm_gbs"]
    gemm_ms, gemm_flops, gemm_n = agg.get(0, [0.0, 0.0, 0])
    family_tf = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    # dominant kernel = the encoder GEMMs (patch embedding + 12 x {qkv, proj, fc1, fc2}): launch order is pe GEMM, LN+pos, then
    # {[LN,] qkv, attention, proj, [LN,] fc1, fc2} per block, final LN: the tensor-kind launches before the final LayerNorm.  At this
    # batch every one of them runs on gemm_tc5_pair_kernel<256, 0> (the kernel with the largest share of the ncu launch list).
    # (the final LayerNorm is the last kind-2 launch whether or not the blocks' LayerNorms are folded into their GEMMs)
    enc_end = max(i for i, (_, kind, _) in enumerate(per_launch) if kind == 2)
    enc = [(ms, fl) for i, (ms, kind, fl) in enumerate(per_launch) if kind == 0 and i < enc_end]
    enc_ms, enc_flops, enc_n = sum(m for m, _ in enc), sum(f for _, f in enc), len(enc)
    achieved_tf = enc_flops / (enc_ms / 1e3) / 1e12 if enc_ms > 0 else 0.0
    if args.backend != 0:
        achieved_tf, enc_ms, enc_flops, enc_n = family_tf, gemm_ms, gemm_flops, gemm_n
    launches_per_step = 2 + (model.launches_per_forward() - 1) + 2      # corrupt(2) + forward (patchify fused away) + metrics(fused pass + finalize)
    n_forward_launches = model.launches_per_forward()

    line = {
        "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "C2: corrupt + SatMAE ViT-B restoration forward + PSNR/SSIM/SAM/RMSE, batch 256 tiles/GPU (13x192x192)",
                   "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world} (batch-sharded, metric-sum allreduce only)",
                   "weights": "random-init seed 42 (reference constructor order), eval-mode BN folded",
                   "noise": "gaussian 0.015, speckle 0.008, dead-band 0.08, thermal 0.005 (configs/base.yaml)",
                   "gemm_backend": "tcgen05" if args.backend == 0 else "mma.sync",
                   "l2": "per-step working set (491 MB clean + 6.2 GB activations) exceeds the 126 MB L2; no flush needed"},
        # headline e2e: the reference's real data flow -- raw uint16 64x64 digital numbers from pinned host memory, the dataset's
        # preprocessing (src/data/dataset.py:160-186) as the ingest kernel in front of the corruption; e2e_f32: pre-processed fp32 tiles
        "e2e": {"value": e2e_value, "unit": "img/s", "ms_per_step": ms_e2e / K, "h2d_bytes_per_step": B * 13 * 64 * 64 * 2,
                "d2h_bytes_per_step": 40, "mode": "uint16 13x64x64 DN tiles (host, pinned) -> H2D -> ingest kernel -> corrupt -> restore -> metrics -> D2H of the 5 fp64 sums"},
        "e2e_f32": {"value": e2e_f32_value, "unit": "img/s", "ms_per_step": ms_e2e_f32 / K, "h2d_bytes_per_step": B * TILE_ELEMS * 4,
                    "d2h_bytes_per_step": 40, "mode": "fp32 13x192x192 tiles (host, pinned) -> H2D -> corrupt -> restore -> metrics -> D2H"},
        "gpu_launches": launches_per_step * K,
        "roofline": {"bound": "tensor", "kernel": "gemm_tc5_pair_kernel<256,0> (cta_group::2 tcgen05 GEMM): the encoder GEMMs pe/qkv/proj/fc1/fc2, 49 of its 54 launches per step" if args.backend == 0 else "gemm_hmma_kernel<BN> (whole GEMM/conv family)",
                     "achieved": achieved_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": achieved_tf / pk["tf_sustained"],
                     "peak_source": f"{pk['src']} bf16_tflops_sustained (kernel timed inside a long step)",
                     # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the most frequent member of the family (the
                     # qkv GEMM, 36 of 70 launches are this shape class), ncu --set full, profiles/r1_v3_ncu_full_summaries.txt;
                     # its algorithmic bytes are 56.6 MB (A) + 3.5 MB (W) + 169.9 MB (out): the 126 MB L2 still holds part of the
                     # output when the capture ends, nothing is re-read
                     "traffic": facts.get("traffic_dominant_bytes") if args.backend == 0 else None,
                     "traffic_kernel": facts.get("traffic_dominant_kernel"),
                     "tensor_pipe_active_pct_ncu": facts.get("tensor_pipe_active_pct"),
                     "ncu_source": facts.get("source"),
                     "launches_per_step": enc_n, "ms_per_step": enc_ms,
                     "algorithmic_gflop_per_step": enc_flops / 1e9,
                     "share_of_step": enc_ms / ms_step if ms_step > 0 else None,
                     # every tensor-core launch of the step (GEMMs, implicit 3x3 convs, ConvTranspose, the head conv with its fused 1x1)
                     "family": {"kernels": "gemm_tc5_pair_kernel<BN,MERGED> / gemm_tc5_kernel<BN,MODE> / conv_roll_kernel<EPI> (stage-4 3x3 convs + fused head)", "achieved": family_tf,
                                "frac": family_tf / pk["tf_sustained"], "launches_per_step": gemm_n, "ms_per_step": gemm_ms,
                                "algorithmic_gflop_per_step": gemm_flops / 1e9, "share_of_step": gemm_ms / ms_step if ms_step > 0 else None},
                     "other_ms_per_step": {"attention": agg.get(1, [0])[0], "layernorm": agg.get(2, [0])[0], "other_model": agg.get(3, [0])[0]},
                     "forward_tflops_whole_step": B * FWD_GFLOP_PER_IMG / ms_step / 1e3},
        # the bandwidth-bound kernels of the same step: algorithmic bytes / CUDA-event time on the launching stream,
        # against the measured copy bandwidth; issue_active_pct / dram_pct from the ncu captures (they are instruction-issue bound:
        # the bit-exact torch Philox stream, the 11x11 blur and the SSIM rational cost more issue slots than HBM cycles)
        "roofline_hbm": [
            {"kernel": "corrupt_kernel (+ dead_mask_kernel): AddSensorNoise, also writes the bf16 patch-embed operand",
             "ms": corrupt_ms, "algorithmic_bytes": B * (2 * TILE_ELEMS * 4 + TILE_ELEMS * 2), "unit": "GB/s",
             "achieved": B * (2 * TILE_ELEMS * 4 + TILE_ELEMS * 2) / corrupt_ms / 1e6, "peak": hbm,
             "frac": B * (2 * TILE_ELEMS * 4 + TILE_ELEMS * 2) / corrupt_ms / 1e6 / hbm, **facts.get("corrupt_kernel", {})},
            {"kernel": "metrics_fused_kernel (+ finalize): PSNR / SSIM / SAM / RMSE, one pass over pred and target",
             "ms": metrics_ms, "algorithmic_bytes": B * 2 * TILE_ELEMS * 4, "unit": "GB/s",
             "achieved": B * 2 * TILE_ELEMS * 4 / metrics_ms / 1e6, "peak": hbm, "frac": B * 2 * TILE_ELEMS * 4 / metrics_ms / 1e6 / hbm,
             **facts.get("metrics_fused_kernel", {})},
            {"kernel": "layernorm_kernel: LN + pos_embed (+ bf16 copy, row sums) and the final LayerNorm; the 24 block LayerNorms ride in GEMM epilogues",
             "ms": agg.get(2, [0.0])[0], "launches": agg.get(2, [0, 0, 0])[2],
             "algorithmic_bytes": B * 144 * 768 * (4 + 4 + 2 + 4 + 2), "unit": "GB/s",
             "achieved": (B * 144 * 768 * 16 / agg[2][0] / 1e6) if agg.get(2, [0.0])[0] > 0 else None, "peak": hbm,
             "frac": (B * 144 * 768 * 16 / agg[2][0] / 1e6 / hbm) if agg.get(2, [0.0])[0] > 0 else None},
        ],
        "clocks": clk.summary(),
        "quality": {k: float(v) for k, v in metrics_dev.items()},
    }
    if not args.no_scene and args.backend == 0:
        with torch.no_grad():
            line["scene"] = scene_bench(model, dev, rank, world)
    if not args.no_train and args.backend == 0:
        del model
        torch.cuda.empty_cache()
        line["train_step"] = train_step_bench(dev, world)
        line["roofline_hbm"].append(loss_bench(dev, hbm, facts))
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_baseline_run(steps=8, warmup=2)
        line["cpu_baseline"] = {"value": r["value"], "unit": "img/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        if "stage_img_per_s" in r:
            line["cpu_baseline"]["stage_img_per_s"] = r["stage_img_per_s"]
    if rank == 0:
        if args.profile_out:
            Path(args.profile_out).parent.mkdir(parents=True, exist_ok=True)
            names = {0: "gemm/conv", 1: "attention", 2: "layernorm", 3: "other"}
            Path(args.profile_out).write_text(json.dumps(
                {"ms_per_step": ms_step, "launches": [{"i": i, "kind": names[k], "ms": ms, "gflop": fl / 1e9,
                                                       "tflops": (fl / ms / 1e9) if ms > 0 and fl > 0 else None}
                                                      for i, (ms, k, fl) in enumerate(per_launch)]}, indent=1))
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
