"""Where does the end-to-end step lose time against the device-resident step?  Variants of bench.py's e2e loop, timed on one box:
    python tools/e2e_probe.py [--steps 20]
Not part of the product path."""
import argparse, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import recipes
from msd_b200.data import AddSensorNoise, ingest_tiles
from msd_b200.models import SatMAERestoration
from msd_b200.training import MetricsTracker

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
K, B = args.steps, 256
dev = torch.device("cuda")
torch.manual_seed(42)
model = SatMAERestoration(recipes.base_config(), block_layout="fallback").eval().to(dev)
noise = AddSensorNoise(*recipes.NOISE_CONFIGS["base"])
clean_host, _ = recipes.eurosat_tiles(32, seed=100)
clean_host = clean_host.repeat(B // 32, 1, 1, 1).contiguous().pin_memory()
clean_dev = clean_host.to(dev)
dn_host = torch.randint(0, 5200, (B, 13, 64, 64), dtype=torch.int32).to(torch.uint16).pin_memory()
dn_dev = dn_host.to(dev)
patches = torch.empty(B * 144, 13 * 256, dtype=torch.bfloat16, device=dev)
copy_stream = torch.cuda.Stream()


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


def resident():
    t = MetricsTracker()
    for _ in range(K):
        t.update(model(noise(clean_dev, patch_out=patches), patches=patches), clean_dev)
    t.compute()


def variant(h2d, ingest, per_step_read, new_tracker):
    src = dn_host if ingest else clean_host
    res = dn_dev if ingest else clean_dev
    bufs = [torch.empty_like(res) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    trackers = [MetricsTracker(), MetricsTracker()]

    def run():
        for b in range(2):
            freed[b].record()
        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[i % 2])
                bufs[i % 2].copy_(src, non_blocking=True)
                ready[i % 2].record(copy_stream)
        if h2d: prefetch(0)
        prev = None
        for i in range(K):
            if h2d:
                if i + 1 < K: prefetch(i + 1)
                torch.cuda.current_stream().wait_event(ready[i % 2])
                c = bufs[i % 2]
            else:
                c = res
            if ingest: c = ingest_tiles(c, 192)
            t = MetricsTracker() if new_tracker else trackers[i % 2]
            t.reset()
            t.update(model(noise(c, patch_out=patches), patches=patches), c)
            if per_step_read:
                t.prefetch()
            if h2d: freed[i % 2].record()
            if per_step_read and prev is not None:
                prev.compute()
            prev = t
        prev.compute()
    return run


with torch.no_grad():
    for _ in range(3):
        resident()
    print(f"resident loop                                  {timed(resident):7.3f} ms/step")
    for name, cfg in [("no h2d, no ingest, read at end", (0, 0, 0, 0)), ("no h2d, no ingest, per-step read", (0, 0, 1, 0)),
                      ("no h2d, ingest, per-step read", (0, 1, 1, 0)), ("h2d u16 + ingest, per-step read (bench e2e)", (1, 1, 1, 0)),
                      ("h2d f32, per-step read (bench e2e_f32)", (1, 0, 1, 0)), ("h2d u16 + ingest, read at end", (1, 1, 0, 0)),
                      ("resident again", None)]:
        if cfg is None:
            print(f"{name:46s} {timed(resident):7.3f} ms/step"); continue
        fn = variant(*cfg)
        fn()
        print(f"{name:46s} {timed(fn):7.3f} ms/step")
