"""Micro-benchmark of the GEMM / implicit-conv kernel on the model's own shapes (B=256), CUDA-event timed.
    python tools/gemm_bench.py [--backend 0|1] [--only NAME] [--iters N]
Used for kernel iteration and as the command ncu wraps (profiles/); not part of the product path."""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import gpu_helpers as gh
from msd_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--backend", type=int, default=0)
ap.add_argument("--only", default=None)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--batch", type=int, default=256)
args = ap.parse_args()
B = args.batch
M = B * 144
dev = torch.device("cuda")
lib = _lib.load()
bf = torch.bfloat16


def timeit(fn, flops, name):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    print(f"{name:28s} {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)


def lin(name, N, K, act=0, resid=False, out_f32=False):
    if args.only and args.only != name:
        return
    a = torch.randn(M, K, device=dev).to(bf)
    w = (torch.randn(N, K, device=dev) * K ** -0.5).to(bf)
    bias = torch.randn(N, device=dev)
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if (out_f32 or resid) else bf)
    s = _lib.stream_ptr()
    def fn():
        _lib.check(lib.msd_linear_bf16(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(out), _lib.ptr(out) if resid else None,
                                       M, N, K, act, int(out_f32 or resid), args.backend, s))
    timeit(fn, 2.0 * M * N * K, name)


def conv(name, C, H, skip):
    if args.only and args.only != name:
        return
    x = torch.randn(B, H + 2, H + 2, C, device=dev).to(bf)
    kpad = (C + 63) // 64 * 64
    w = (torch.randn(C, (10 if skip else 9) * kpad, device=dev) * (9 * C) ** -0.5).to(bf)
    bias = torch.randn(C, device=dev)
    y = torch.empty_like(x)
    sk = torch.randn_like(x) if skip else None
    s = _lib.stream_ptr()
    def fn():
        _lib.check(lib.msd_conv3x3_nhwc_bf16(_lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(sk), _lib.ptr(y), B, H, H, C, C, args.backend, s))
    timeit(fn, 2.0 * B * H * H * C * 9 * C, name)


def convt(name, Cin, Cout, Hin, padded):
    if args.only and args.only != name:
        return
    ph = Hin + 2 if padded else Hin
    x = torch.randn(B, ph, ph, Cin, device=dev).to(bf)
    kpad = (Cin + 63) // 64 * 64
    w = (torch.randn(4 * Cout, kpad, device=dev) * Cin ** -0.5).to(bf)
    bias = torch.randn(4 * Cout, device=dev)
    y = torch.empty(B, 2 * Hin + 2, 2 * Hin + 2, Cout, device=dev, dtype=bf)
    s = _lib.stream_ptr()
    def fn():
        _lib.check(lib.msd_convt2x2_nhwc_bf16(_lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(y), B, Hin, Hin, int(padded), Cin, Cout, args.backend, s))
    timeit(fn, 2.0 * B * Hin * Hin * 4 * Cout * Cin, name)


def lin_ln(name, N, K, act=0, producer=False):
    """the folded-LayerNorm variants: producer = fp32 residual epilogue + bf16 copy + partial row sums; consumer = rstd/mean epilogue"""
    if args.only and args.only != name:
        return
    parts = lib.msd_ln_stats_parts(768, M)
    a = torch.randn(M, K, device=dev).to(bf)
    w = (torch.randn(N, K, device=dev) * K ** -0.5).to(bf)
    bias, lns = torch.randn(N, device=dev), torch.randn(N, device=dev)
    stats = torch.rand(M, parts, 2, device=dev)
    s = _lib.stream_ptr()
    if producer:
        x, xb = torch.zeros(M, N, device=dev), torch.empty(M, N, device=dev, dtype=bf)
        def fn():
            _lib.check(lib.msd_linear_resid_stats_f32(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(x), _lib.ptr(xb), _lib.ptr(stats), parts, M, N, K, s))
    else:
        out = torch.empty(M, N, device=dev, dtype=bf)
        def fn():
            _lib.check(lib.msd_linear_lnfold_bf16(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(lns), _lib.ptr(stats), parts, 1e-5, _lib.ptr(out), M, N, K, act, s))
    timeit(fn, 2.0 * M * N * K, name)


lin("pe", 768, 3328, out_f32=True)
lin("qkv", 2304, 768)
lin("proj", 768, 768, resid=True)
lin("fc1", 3072, 768, act=1)
lin("fc2", 768, 3072, resid=True)
if args.backend == 0:
    lin_ln("qkv_ln", 2304, 768)
    lin_ln("proj_stats", 768, 768, producer=True)
    lin_ln("fc1_ln", 3072, 768, act=1)
    lin_ln("fc2_stats", 768, 3072, producer=True)
convt("convt1", 768, 384, 12, False)
conv("conv1", 384, 24, False)
conv("conv1_skip", 384, 24, True)
convt("convt2", 384, 192, 24, True)
conv("conv2", 192, 48, False)
conv("conv2_skip", 192, 48, True)
convt("convt3", 192, 96, 48, True)
conv("conv3", 96, 96, False)
conv("conv3_skip", 96, 96, True)
convt("convt4", 96, 48, 96, True)
conv("conv4", 48, 192, False)
conv("conv4_skip", 48, 192, True)
