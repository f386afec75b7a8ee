"""Micro-benchmark of the bandwidth-bound passes of the training step at the decoder's shapes (micro-batch 64):
BatchNorm + GELU forward / backward (msd_bn_act_*), LayerNorm backward, GELU backward.  Prints time and the fraction of the
HBM rate the algorithmic bytes of each pass correspond to.
    python tools/bn_bench.py [--batch 64] [--iters 20]"""
import argparse, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from msd_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--only", type=int, default=0, help="decoder resolution to run alone (192, 96, 48, 24); 0 = everything")
ap.add_argument("--wgrad", action="store_true", help="time the 3x3 conv weight gradients instead")
args = ap.parse_args()
lib = _lib.load()
s = _lib.stream_ptr()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, nbytes, name):
    for _ in range(2):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.iters)]
    for e0, e1 in ev:
        flush.zero_()                      # the operands of a pass were written long before it runs in the real step
        e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev)[len(ev) // 2]
    print(f"{name:44s} {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:8.1f} GB/s algorithmic", flush=True)


B = args.batch
if args.wgrad:
    for (H, C) in ((192, 48), (96, 96), (48, 192), (24, 384)):
        if args.only and H != args.only:
            continue
        x = torch.randn(B, H + 2, H + 2, C, device="cuda").to(torch.bfloat16)
        dc = torch.randn(B, H + 2, H + 2, C, device="cuda").to(torch.bfloat16)
        kpad = (C + 63) // 64 * 64
        dw = torch.zeros(9, kpad, C, device="cuda")
        T = x.numel() * 2
        gflop = 2.0 * 9 * C * C * B * H * H / 1e9
        timeit(lambda: _lib.check(lib.msd_wgrad_conv3x3_bf16(_lib.ptr(x), _lib.ptr(dc), _lib.ptr(dw), B, H, H, C, C, kpad, C, s)), 2 * T,
               f"wgrad conv3x3 {H}x{H}x{C} ({gflop:.0f} GFLOP)")
    sys.exit(0)
for (H, C) in ((192, 48), (96, 96), (48, 192), (24, 384)):
    if args.only and H != args.only:
        continue
    shape = (B, H + 2, H + 2, C)
    T = B * (H + 2) * (H + 2) * C * 2
    c = torch.randn(shape, device="cuda").to(torch.bfloat16)
    skip = torch.randn(shape, device="cuda").to(torch.bfloat16)
    dy = torch.randn(shape, device="cuda").to(torch.bfloat16)
    y, dz, dc = torch.empty_like(c), torch.empty_like(c), torch.empty_like(c)
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    stats, coef = torch.empty(4, C, device="cuda"), torch.empty(2, C, device="cuda")
    sums = torch.empty(2, C, dtype=torch.float64, device="cuda")
    dgam, dbet = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    _lib.check(lib.msd_bn_stats_bf16(_lib.ptr(c), B, H, H, C, _lib.ptr(gamma), _lib.ptr(beta), 1e-5, 0.1, _lib.ptr(rm), _lib.ptr(rv),
                                     _lib.ptr(stats), _lib.ptr(sums), s))
    for sk in (None, skip):
        tag = f"{H}x{H}x{C}{' +skip' if sk is not None else ''}"
        n_in = 2 if sk is not None else 1
        timeit(lambda: _lib.check(lib.msd_bn_act_fwd_bf16(_lib.ptr(c), _lib.ptr(sk), _lib.ptr(stats), _lib.ptr(y), B, H, H, C, s)),
               (n_in + 1) * T, f"bn_act_fwd {tag}")
        timeit(lambda: _lib.check(lib.msd_bn_act_bwd_bf16(_lib.ptr(dy), _lib.ptr(c), _lib.ptr(sk), _lib.ptr(stats), _lib.ptr(dz), _lib.ptr(dc),
                                                          _lib.ptr(dgam), _lib.ptr(dbet), _lib.ptr(coef), _lib.ptr(sums), B, H, H, C, s)),
               (n_in + 2 + 3) * T, f"bn_act_bwd (both passes) {tag}")
    timeit(lambda: _lib.check(lib.msd_bn_sums_bf16(_lib.ptr(c), B * (H + 2) * (H + 2), C, _lib.ptr(sums), s)), T, f"bn_sums {H}x{H}x{C}")

if args.only:
    sys.exit(0)
M, D, Hd = B * 144, 768, 3072
x = torch.randn(M, D, device="cuda")
dyb = torch.randn(M, D, device="cuda").to(torch.bfloat16)
g = torch.randn(M, D, device="cuda")
gbf = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
gam = torch.ones(D, device="cuda")
dg, db, cs = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
timeit(lambda: _lib.check(lib.msd_layernorm_bwd_colsum(_lib.ptr(dyb), 0, _lib.ptr(x), _lib.ptr(gam), 1e-6, _lib.ptr(g), 1, _lib.ptr(gbf),
                                                       _lib.ptr(dg), _lib.ptr(db), _lib.ptr(cs), M, D, s)),
       M * D * (4 + 2 + 8 + 2), "layernorm_bwd (+colsum) [9216 x 768]")
dh = torch.randn(M, Hd, device="cuda").to(torch.bfloat16)
hp = torch.randn(M, Hd, device="cuda").to(torch.bfloat16)
csh = torch.zeros(Hd, device="cuda")
timeit(lambda: _lib.check(lib.msd_gelu_bwd_bf16(_lib.ptr(dh), _lib.ptr(hp), _lib.ptr(dh), M * Hd, s)), M * Hd * 6, "gelu_bwd [9216 x 3072]")
timeit(lambda: _lib.check(lib.msd_gelu_bwd_colsum_bf16(_lib.ptr(dh), _lib.ptr(hp), _lib.ptr(dh), _lib.ptr(csh), M, Hd, s)), M * Hd * 6,
       "gelu_bwd + colsum [9216 x 3072]")
