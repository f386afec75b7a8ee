"""Time the decoder head at the bench shape (B = 256, 192 x 192, 48 -> 48 -> 13): fused conv3x3+1x1 kernel vs conv3x3 then head1x1.
    python tools/head_bench.py [B]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import torch.nn as nn
import gpu_helpers as gh
from msd_b200 import _lib
from msd_b200.models.satmae_restoration import pack_conv3x3

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H = W = 192
torch.manual_seed(0)
conv, bn = nn.Conv2d(48, 48, 3, padding=1, bias=True), nn.BatchNorm2d(48).eval()
with torch.no_grad():
    w, b = pack_conv3x3(conv, bn)
wc, bc = w.cuda(), b.cuda()
w1, b1 = (torch.randn(13, 48) * 0.3).cuda(), torch.randn(13).cuda()
xh = torch.randn(B, H + 2, W + 2, 48, device="cuda").bfloat16()
xh[:, 0] = 0; xh[:, -1] = 0; xh[:, :, 0] = 0; xh[:, :, -1] = 0
y = torch.empty(B, 13, H, W, device="cuda")
mid = torch.empty(B, H + 2, W + 2, 48, device="cuda", dtype=torch.bfloat16)
lib = _lib.load()
st = _lib.stream_ptr()

def fused():
    _lib.check(lib.msd_conv3x3_head_nchw_f32(_lib.ptr(xh), _lib.ptr(wc), _lib.ptr(bc), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(y), B, H, W, 48, 13, st))

def separate():
    _lib.check(lib.msd_conv3x3_nhwc_bf16(_lib.ptr(xh), _lib.ptr(wc), _lib.ptr(bc), None, _lib.ptr(mid), B, H, W, 48, 48, 0, st))
    _lib.check(lib.msd_head1x1_nchw_f32(_lib.ptr(mid), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(y), B, H, W, 48, 13, st))

def conv_only():
    _lib.check(lib.msd_conv3x3_nhwc_bf16(_lib.ptr(xh), _lib.ptr(wc), _lib.ptr(bc), None, _lib.ptr(mid), B, H, W, 48, 48, 0, st))

def head_only():
    _lib.check(lib.msd_head1x1_nchw_f32(_lib.ptr(mid), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(y), B, H, W, 48, 13, st))

for name, fn in (("fused conv3x3+1x1", fused), ("conv3x3 + head1x1", separate), ("conv3x3 alone", conv_only), ("head1x1 alone", head_only)):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10:.4f} ms", flush=True)
