"""Micro-benchmark of the bandwidth-bound kernels at the C2 batch size (256 tiles): corruption, fused metrics, loss.
    python tools/elem_bench.py [--only corrupt|metrics|loss] [--iters N]"""
import argparse, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import recipes
from msd_b200.data import AddSensorNoise
from msd_b200.training import CombinedLoss
from msd_b200.training.metrics import fused_metrics

ap = argparse.ArgumentParser()
ap.add_argument("--only", default=None)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--batch", type=int, default=256)
args = ap.parse_args()
B = args.batch
a, b = recipes.eurosat_tiles(32, seed=5)
a = a.repeat(B // 32, 1, 1, 1).cuda().contiguous(); b = b.repeat(B // 32, 1, 1, 1).cuda().contiguous()
nbytes = a.numel() * 4


def timeit(fn, algo_bytes, name):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(args.iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    print(f"{name:24s} {ms:8.3f} ms  {algo_bytes / ms / 1e6:8.1f} GB/s algorithmic", flush=True)


if args.only in (None, "corrupt"):
    noise = AddSensorNoise(*recipes.NOISE_CONFIGS["base"])
    timeit(lambda: noise(a), 2 * nbytes, "corrupt")
    patches = torch.empty(B * 144, 13 * 256, dtype=torch.bfloat16, device="cuda")
    timeit(lambda: noise(a, patch_out=patches), 2 * nbytes + nbytes // 2, "corrupt+patches")
if args.only in (None, "metrics"):
    timeit(lambda: fused_metrics(b, a), 2 * nbytes, "metrics (ssim+sam)")
if args.only in (None, "loss"):
    crit = CombinedLoss()
    def fwd_bwd():
        p = b.detach().requires_grad_(True)
        crit(p, a).backward()
    timeit(fwd_bwd, 3 * nbytes, "loss fwd+bwd")
    with torch.no_grad():
        timeit(lambda: crit(b, a), 2 * nbytes, "loss fwd")
if args.only in (None, "attention"):
    from msd_b200 import _lib
    lib = _lib.load()
    qkv = torch.randn(B * 144, 2304, device="cuda").to(torch.bfloat16)
    out = torch.empty(B * 144, 768, device="cuda", dtype=torch.bfloat16)
    def att():
        _lib.check(lib.msd_attention_bf16(_lib.ptr(qkv), _lib.ptr(out), B, 144, 12, _lib.stream_ptr()))
    timeit(att, qkv.numel() * 2 + out.numel() * 2, "attention fwd")
    print(f"   attention: {4.0 * B * 12 * 144 * 144 * 64 / 1e9:.1f} GFLOP per launch")
