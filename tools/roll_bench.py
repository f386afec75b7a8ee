"""Timing of the rolling-row 48-channel conv (csrc/conv_roll.cu) on arbitrary H x W x B, CUDA-event timed.
    python tools/roll_bench.py H W B [skip]
MSD_NO_ROLL=1 times the tap-major kernel on the same shape.  Kernel iteration tool, not part of the product path."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
from msd_b200 import _lib

H, W, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
skip = len(sys.argv) > 4
dev, bf = torch.device("cuda"), torch.bfloat16
lib = _lib.load()
x = torch.randn(B, H + 2, W + 2, 48, device=dev).to(bf)
w = (torch.randn(48, (10 if skip else 9) * 64, device=dev) * 432 ** -0.5).to(bf)
bias = torch.randn(48, device=dev)
y = torch.empty_like(x)
sk = torch.randn_like(x) if skip else None
s = _lib.stream_ptr()


def fn():
    _lib.check(lib.msd_conv3x3_nhwc_bf16(_lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(sk), _lib.ptr(y), B, H, W, 48, 48, 0, s))


for _ in range(3):
    fn()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
px = B * H * W
# run-to-run and batch-composition determinism (bitwise): the same images inside a different launch must come out identical
y0 = y.clone()
fn()
torch.cuda.synchronize()
same_run = torch.equal(y0, y)
b0, b1 = B // 3, B // 3 + max(1, B // 5)
xs, ys = x[b0:b1].contiguous(), torch.empty_like(x[b0:b1])
sks = sk[b0:b1].contiguous() if skip else None
_lib.check(lib.msd_conv3x3_nhwc_bf16(_lib.ptr(xs), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(sks), _lib.ptr(ys), b1 - b0, H, W, 48, 48, 0, s))
torch.cuda.synchronize()
# stress: 30 more launches, each compared bit for bit with the first
nbad = 0
for _ in range(30):
    y.zero_()
    fn()
    torch.cuda.synchronize()
    nbad += 0 if torch.equal(y0, y) else 1
print(f"deterministic: run-to-run {same_run}, sub-batch {torch.equal(ys, y0[b0:b1])}, 30 repeats with {nbad} mismatches")
print(f"H={H} W={W} B={B} skip={int(skip)}: {ms:.3f} ms  {px / ms / 1e6:.2f} Gpx/s  {2.0 * px * 432 * 48 / ms / 1e9:.0f} TFLOP/s")
