"""Time the two attention kernels at the bench shape (B = 256 images, 12 heads, 144 tokens): python tools/attn_bench.py [B]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from msd_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
FLUSH = "--noflush" not in sys.argv      # --noflush with a small B: operands stay in L2 (compute-side time per problem)
T, H, E = 144, 12, 768
lib = _lib.load()
qkv = (torch.randn(B * T, 3 * E, device="cuda") * 1.5).bfloat16()
out = torch.empty(B * T, E, dtype=torch.bfloat16, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name in ("msd_attention_tc5_bf16", "msd_attention_bf16"):
    fn = getattr(lib, name)
    for _ in range(3):
        _lib.check(fn(_lib.ptr(qkv), _lib.ptr(out), B, T, H, _lib.stream_ptr()))
    ts = []
    for _ in range(20):
        if FLUSH:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(fn(_lib.ptr(qkv), _lib.ptr(out), B, T, H, _lib.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    gflop = 4.0 * B * H * T * T * 64 / 1e9
    gbytes = B * T * 4 * E * 2 / 1e9
    print(f"B={B} flush={FLUSH} {name}: {ms:.4f} ms  {gflop / ms:.0f} TFLOP/s  {gbytes / ms * 1e3:.0f} GB/s (q,k,v in + out)")
