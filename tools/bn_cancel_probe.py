"""How much does bf16 storage of the incoming gradient cost in BatchNorm backward when dy has a per-channel DC part?"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch, torch.nn.functional as F
from msd_b200 import _lib
lib = _lib.load()
B, H, W, C = 2, 96, 96, 96
g = torch.Generator().manual_seed(0)
bf = lambda t: t.to(torch.bfloat16)
halo = lambda t: bf(F.pad(t.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1))).contiguous()
for dc_off in (0.0, 1.0, 3.0, 10.0):
    c = torch.randn(B, C, H, W, generator=g)
    dy32 = torch.randn(B, C, H, W, generator=g) + dc_off * torch.randn(1, C, 1, 1, generator=g)
    for name, dy in (("fp32 dy ref vs bf16-rounded dy ref", bf(dy32).float()),):
        cr = bf(c).float().requires_grad_(True)
        gam, bet = torch.ones(C, requires_grad=True), torch.zeros(C, requires_grad=True)
        F.gelu(F.batch_norm(cr, None, None, gam, bet, True, 0.1, 1e-5)).backward(dy32)
        ref32 = cr.grad.clone(); cr.grad = None
        F.gelu(F.batch_norm(cr, None, None, gam, bet, True, 0.1, 1e-5)).backward(dy)
        refbf = cr.grad.clone()
    ch, dyh = halo(bf(c).float()).cuda(), halo(dy32).cuda()
    stats = torch.empty(4, C, device="cuda"); sums = torch.empty(2, C, dtype=torch.float64, device="cuda")
    gd, bd = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    s = _lib.stream_ptr()
    _lib.check(lib.msd_bn_stats_bf16(_lib.ptr(ch), B, H, W, C, _lib.ptr(gd), _lib.ptr(bd), 1e-5, 0.1, None, None, _lib.ptr(stats), _lib.ptr(sums), s))
    dz = torch.empty_like(ch); dcb = torch.empty_like(ch); coef = torch.empty(2, C, device="cuda")
    _lib.check(lib.msd_bn_act_bwd_bf16(_lib.ptr(dyh), _lib.ptr(ch), None, _lib.ptr(stats), _lib.ptr(dz), _lib.ptr(dcb), None, None, _lib.ptr(coef), _lib.ptr(sums), B, H, W, C, s))
    got = dcb[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float().cpu()
    rel = lambda a, b: ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()
    print(f"dc_off {dc_off:5.1f}: |dy| rms {dy32.pow(2).mean().sqrt():.3f} |dc| rms {ref32.pow(2).mean().sqrt():.3f}  kernel vs fp32-dy ref {rel(got, ref32):.4f}  bf16-dy torch ref vs fp32-dy ref {rel(refbf, ref32):.4f}")
