"""GPU parity of the training-step kernels (weight-gradient contractions, BatchNorm / LayerNorm / GELU / attention
backward, the whole SatMAERestoration forward+backward in training mode) against torch fp32 autograd of the oracle.

Tolerance: bf16 operands with fp32 accumulation -- errors are judged relative to the largest reference magnitude of the
tensor (the same budget the forward GEMM tests use)."""
import pytest
import torch
import torch.nn.functional as F

from msd_b200 import _lib

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16)


def _close(got, ref, rel):
    scale = ref.abs().max().item() + 1e-12
    err = (got.float() - ref.float()).abs().max().item()
    assert err <= rel * scale, f"max err {err:.4g} vs scale {scale:.4g} (rel {err / scale:.3g})"


@pytest.mark.parametrize("rows,out_f,in_f", [(256, 128, 64), (1000, 768, 768), (4608, 2304, 768), (777, 192, 96),
                                             (2000, 48, 200), (64, 13 * 16, 3328), (5000, 3072, 768)])
def test_wgrad_linear(rows, out_f, in_f):
    g = torch.Generator().manual_seed(rows + out_f)
    dy = _bf(torch.randn(rows, out_f, generator=g)).cuda()
    x = _bf(torch.randn(rows, in_f, generator=g)).cuda()
    dw = torch.zeros(out_f, in_f, device="cuda")
    lib = _lib.load()
    _lib.check(lib.msd_wgrad_linear_bf16(_lib.ptr(dy), out_f, _lib.ptr(x), in_f, _lib.ptr(dw), in_f, rows, out_f, in_f, _lib.stream_ptr()))
    ref = dy.float().t() @ x.float()
    _close(dw, ref, 2e-3)
    # accumulates: a second call doubles the result
    _lib.check(lib.msd_wgrad_linear_bf16(_lib.ptr(dy), out_f, _lib.ptr(x), in_f, _lib.ptr(dw), in_f, rows, out_f, in_f, _lib.stream_ptr()))
    _close(dw, 2 * ref, 2e-3)


def test_wgrad_linear_exact_on_small_integers():
    g = torch.Generator().manual_seed(3)
    rows, out_f, in_f = 300, 256, 320
    dy = torch.randint(-3, 4, (rows, out_f), generator=g).float()
    x = torch.randint(-3, 4, (rows, in_f), generator=g).float()
    dw = torch.zeros(out_f, in_f, device="cuda")
    lib = _lib.load()
    a, b = _bf(dy).cuda(), _bf(x).cuda()
    _lib.check(lib.msd_wgrad_linear_bf16(_lib.ptr(a), out_f, _lib.ptr(b), in_f, _lib.ptr(dw), in_f, rows, out_f, in_f, _lib.stream_ptr()))
    assert torch.equal(dw.cpu(), dy.t() @ x)


def _halo(t_nchw):
    """[B,C,H,W] fp32 -> haloed NHWC bf16 [B,H+2,W+2,C] with a zero border."""
    return _bf(F.pad(t_nchw.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1))).contiguous()


@pytest.mark.parametrize("B,H,W,cin,cout", [(2, 8, 8, 48, 48), (3, 12, 10, 96, 96), (2, 6, 6, 192, 192), (1, 24, 24, 384, 384),
                                            (2, 16, 16, 48, 96), (5, 40, 40, 48, 48)])
def test_wgrad_conv3x3(B, H, W, cin, cout):
    g = torch.Generator().manual_seed(B * H + cin)
    x = _bf(torch.randn(B, cin, H, W, generator=g)).float()
    dc = _bf(torch.randn(B, cout, H, W, generator=g)).float()
    w = torch.zeros(cout, cin, 3, 3, requires_grad=True)
    F.conv2d(x, w, padding=1).backward(dc)
    ref = w.grad                                                     # [cout, cin, 3, 3]
    cinpad = (cin + 63) // 64 * 64
    dw = torch.zeros(9, cinpad, cout, device="cuda")
    lib = _lib.load()
    xh, dch = _halo(x).cuda(), _halo(dc).cuda()
    _lib.check(lib.msd_wgrad_conv3x3_bf16(_lib.ptr(xh), _lib.ptr(dch), _lib.ptr(dw), B, H, W, cin, cout, cinpad, cout, _lib.stream_ptr()))
    got = dw.cpu()[:, :cin, :].permute(2, 1, 0).reshape(cout, cin, 3, 3)
    _close(got, ref, 2e-3)
    assert dw[:, cin:, :].abs().max().item() == 0 if cinpad > cin else True
