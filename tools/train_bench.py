"""Training-step timing (BASELINE config C4): micro-batch 64 tiles/GPU, fwd + MSE+0.1*SSIM loss + bwd (+ AdamW step),
blocks 0-5 frozen.  CUDA-event timed per phase; not the driver's bench line (bench.py reports it as `train_step`)."""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import recipes
from msd_b200.data import AddSensorNoise
from msd_b200.models import SatMAERestoration
from msd_b200.training import CombinedLoss, FusedAdamW

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--no-opt", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda")
torch.manual_seed(42)
model = SatMAERestoration(recipes.base_config(), block_layout="fallback").to(dev).train()
noise = AddSensorNoise(*recipes.NOISE_CONFIGS["base"])
crit = CombinedLoss(1.0, 0.1)
opt = FusedAdamW(model.get_optimizer_param_groups(1e-4, 0.05), betas=(0.9, 0.999), max_grad_norm=1.0)
clean, _ = recipes.eurosat_tiles(32, seed=3)
clean = clean.repeat(args.batch // 32 + 1, 1, 1, 1)[:args.batch].contiguous().to(dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
tot = {"corrupt": 0.0, "forward": 0.0, "loss": 0.0, "backward": 0.0, "clip+opt": 0.0}
for it in range(args.steps + 2):
    e = [ev() for _ in range(6)]
    e[0].record()
    noisy = noise(clean)
    e[1].record()
    out = model(noisy)
    e[2].record()
    loss = crit(out, clean)
    e[3].record()
    loss.backward()
    e[4].record()
    if not args.no_opt:
        opt.step()
    opt.zero_grad(set_to_none=True)
    e[5].record()
    torch.cuda.synchronize()
    if it >= 2:
        for k, a, b in zip(tot, e[:-1], e[1:]):
            tot[k] += a.elapsed_time(b)
ms = {k: v / args.steps for k, v in tot.items()}
step = sum(ms.values())
print({k: round(v, 2) for k, v in ms.items()}, "ms/step", round(step, 2), "img/s", round(args.batch / step * 1e3, 1), "loss", float(loss),
      "train GFLOP/img 146.4 ->", round(146.4 * args.batch / step, 1), "TFLOP/s", "peak mem GB", round(torch.cuda.max_memory_allocated() / 1e9, 1))

if "--profile" in sys.argv or True:
    import time
    from torch.profiler import profile, ProfilerActivity
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    noisy = noise(clean); out = model(noisy)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"forward host-side call {1e3*(t1-t0):.2f} ms, until GPU done {1e3*(t2-t0):.2f} ms")
    loss = crit(out, clean); loss.backward(); opt.zero_grad(set_to_none=True)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        noisy = noise(clean); out = model(noisy); loss = crit(out, clean); loss.backward()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
