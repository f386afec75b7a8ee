"""Does replaying the eval forward from a CUDA graph (no per-kernel launch latency, back-to-back scheduling) beat stream launches?
    python tools/graph_probe.py        (timing experiment, not the product path)"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import recipes
from msd_b200.data import AddSensorNoise
from msd_b200.models import SatMAERestoration

B, K = 256, 20
dev = torch.device("cuda")
torch.manual_seed(42)
model = SatMAERestoration(recipes.base_config(), block_layout="fallback").eval().to(dev)
noise = AddSensorNoise(*recipes.NOISE_CONFIGS["base"])
clean, _ = recipes.eurosat_tiles(32, seed=100)
clean = clean.repeat(B // 32, 1, 1, 1).contiguous().to(dev)
patches = torch.empty(B * 144, 13 * 256, dtype=torch.bfloat16, device=dev)
noisy = noise(clean, patch_out=patches)


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(K):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


with torch.no_grad():
    for _ in range(3):
        y = model(noisy, patches=patches)
    print(f"stream launches : {timed(lambda: model(noisy, patches=patches)):.3f} ms per forward")
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        model(noisy, patches=patches)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        yg = model(noisy, patches=patches)
    g.replay(); torch.cuda.synchronize()
    print("graph output equals stream output:", torch.equal(yg, y))
    print(f"graph replay    : {timed(g.replay):.3f} ms per forward")
    print(f"stream launches : {timed(lambda: model(noisy, patches=patches)):.3f} ms per forward")
