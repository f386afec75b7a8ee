"""Full-scene restoration timing (BASELINE config C5): synthetic 13 x 10980 x 10980 scene, 192 x 192 tiles, overlap 32.
    python tools/scene_bench.py [--size 10980] [--overlap 32] [--batch 256]
Under torchrun every rank restores its row band (no data-path collective); the time reported is the slowest rank's."""
import argparse, json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import recipes
from msd_b200.models import SatMAERestoration
from msd_b200.scene import grid_origins, restore_scene

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=10980)
ap.add_argument("--overlap", type=int, default=32)
ap.add_argument("--batch", type=int, default=256)
args = ap.parse_args()
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(42)
model = SatMAERestoration(recipes.base_config(), block_layout="fallback").eval().cuda()
g = torch.Generator(device="cuda"); g.manual_seed(1)
scene = torch.empty(13, args.size, args.size, device="cuda")
for c in range(13):
    scene[c] = torch.rand(args.size, args.size, generator=g, device="cuda") * 4 - 2
restore_scene(model, scene[:, :400, :400].contiguous(), overlap=args.overlap, batch=4)          # warm-up
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0.record()
band, rows = restore_scene(model, scene, overlap=args.overlap, batch=args.batch, rank=rank, world=world)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
n = len(grid_origins(args.size, 192, args.overlap)) ** 2
if rank == 0:
    print(json.dumps({"scene": f"13x{args.size}x{args.size}", "overlap": args.overlap, "tiles": n, "world": world, "seconds": ms.item() / 1e3,
                      "tiles_per_s": n / (ms.item() / 1e3), "scene_GB": scene.numel() * 4 / 1e9, "rows_of_rank0": rows,
                      "finite": bool(torch.isfinite(band).all())}))
if world > 1:
    dist.destroy_process_group()
