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


# ----------------------------------------------------------------------------------------------- elementwise / reductions
def _interior(t_halo):
    return t_halo[:, 1:-1, 1:-1, :]


def test_colsum_and_batch_sum():
    lib = _lib.load()
    g = torch.Generator().manual_seed(1)
    for rows, C in [(1000, 48), (5000, 768), (300, 3072), (77, 2304), (4000, 16)]:
        a = _bf(torch.randn(rows, C, generator=g)).cuda()
        sums = torch.empty(C, dtype=torch.float64, device="cuda")
        _lib.check(lib.msd_colsum_bf16(_lib.ptr(a), rows, C, _lib.ptr(sums), _lib.stream_ptr()))
        ref = a.double().sum(0)
        assert (sums - ref).abs().max().item() <= 1e-3 * ref.abs().max().item() + 1e-4
        out = torch.empty(C, device="cuda")
        _lib.check(lib.msd_f64_to_f32(_lib.ptr(sums), _lib.ptr(out), C, _lib.stream_ptr()))
        assert torch.equal(out, sums.float())
    x = torch.randn(5, 144 * 768, generator=g).cuda()
    out = torch.empty(144 * 768, device="cuda")
    _lib.check(lib.msd_batch_sum_f32(_lib.ptr(x), _lib.ptr(out), 5, 144 * 768, _lib.stream_ptr()))
    assert torch.allclose(out, x.sum(0), atol=1e-5)


@pytest.mark.parametrize("B,H,W,C,with_skip", [(2, 8, 8, 48, False), (3, 10, 6, 96, True), (2, 24, 24, 384, True), (4, 20, 20, 192, False)])
def test_batchnorm_train_fwd_bwd(B, H, W, C, with_skip):
    lib = _lib.load()
    g = torch.Generator().manual_seed(B + C)
    c = _bf(torch.randn(B, C, H, W, generator=g) * 1.5 + 0.3).float()
    skip = _bf(torch.randn(B, C, H, W, generator=g)).float() if with_skip else None
    gamma = (1 + 0.2 * torch.randn(C, generator=g))
    beta = 0.1 * torch.randn(C, generator=g)
    dy = _bf(torch.randn(B, C, H, W, generator=g)).float()
    # torch reference (fp32 autograd, training-mode BN)
    cr = c.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    sr = skip.clone().requires_grad_(True) if with_skip else None
    rm, rv = torch.zeros(C), torch.ones(C)
    z = F.batch_norm(cr, rm, rv, gr, br, True, 0.1, 1e-5)
    if with_skip:
        z = z + sr
    y_ref = F.gelu(z)
    y_ref.backward(dy)
    # CUDA path
    ch = _halo(c).cuda()
    sh = _halo(skip).cuda() if with_skip else None
    stats = torch.empty(4, C, device="cuda")
    sums = torch.empty(2, C, dtype=torch.float64, device="cuda")
    rm_d, rv_d = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    s = _lib.stream_ptr()
    gamma_d, beta_d = gamma.cuda(), beta.cuda()          # keep the device copies alive across the asynchronous calls
    _lib.check(lib.msd_bn_stats_bf16(_lib.ptr(ch), B, H, W, C, _lib.ptr(gamma_d), _lib.ptr(beta_d), 1e-5, 0.1, _lib.ptr(rm_d),
                                     _lib.ptr(rv_d), _lib.ptr(stats), _lib.ptr(sums), s))
    assert torch.allclose(rm_d.cpu(), rm, atol=1e-5) and torch.allclose(rv_d.cpu(), rv, atol=1e-4)
    y = torch.full((B, H + 2, W + 2, C), 7.0, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.msd_bn_act_fwd_bf16(_lib.ptr(ch), _lib.ptr(sh), _lib.ptr(stats), _lib.ptr(y), B, H, W, C, s))
    _close(_interior(y).cpu().permute(0, 3, 1, 2), y_ref.detach(), 1e-2)
    assert y[:, 0].abs().max().item() == 0 and y[:, :, 0].abs().max().item() == 0 and y[:, -1].abs().max().item() == 0
    dyh = _halo(dy).cuda()
    dz = torch.full_like(y, 7.0)
    dc = torch.full_like(y, 7.0)
    dgam, dbet = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    coef = torch.empty(2, C, device="cuda")
    _lib.check(lib.msd_bn_act_bwd_bf16(_lib.ptr(dyh), _lib.ptr(ch), _lib.ptr(sh), _lib.ptr(stats), _lib.ptr(dz), _lib.ptr(dc),
                                       _lib.ptr(dgam), _lib.ptr(dbet), _lib.ptr(coef), _lib.ptr(sums), B, H, W, C, s))
    _close(_interior(dc).cpu().permute(0, 3, 1, 2), cr.grad, 1.5e-2)
    _close(dgam.cpu(), gr.grad, 1e-2)
    _close(dbet.cpu(), br.grad, 1e-2)
    if with_skip:
        _close(_interior(dz).cpu().permute(0, 3, 1, 2), sr.grad, 1e-2)
    assert dc[:, 0].abs().max().item() == 0 and dz[:, :, -1].abs().max().item() == 0


@pytest.mark.parametrize("B,H,W,C,bias,fused", [(2, 20, 12, 48, False, 1), (5, 40, 41, 48, True, 1), (3, 30, 31, 96, False, 1),
                                                (1, 10, 14, 96, False, 0), (2, 8, 8, 192, False, 0), (64, 64, 64, 48, False, 1),
                                                # geometries of the rolling-row kernel (csrc/conv_roll.cu, E_CONV_RAW): pair strips only, full + pair, full only
                                                (3, 10, 192, 48, True, 1), (2, 5, 256, 48, False, 1), (9, 192, 192, 48, False, 1)])
def test_conv3x3_with_batchnorm_sums_in_the_epilogue(B, H, W, C, bias, fused):
    """Training forward of nn.Conv2d + nn.BatchNorm2d (blocks.py:66-72): the 48- / 96-channel convs leave sum c, sum c^2 of the tensor
    they store from their own epilogues; the sums must be those of the stored bf16 values (what the separate pass computes), halo rows
    and ragged last tiles included, and identical from run to run."""
    from msd_b200.models.satmae_restoration import pack_conv3x3_raw
    lib = _lib.load()
    g = torch.Generator().manual_seed(B * H + C)
    w = torch.randn(C, C, 3, 3, generator=g) / (3 * C ** 0.5)
    b = torch.randn(C, generator=g).cuda() if bias else None
    x = _bf(torch.randn(B, C, H, W, generator=g)).float()
    xh, wp = _halo(x).cuda(), pack_conv3x3_raw(w).cuda()
    s = _lib.stream_ptr()
    import ctypes

    def run():
        y = torch.full((B, H + 2, W + 2, C), 7.0, device="cuda", dtype=torch.bfloat16)
        sums = torch.full((2, C), -1.0, dtype=torch.float64, device="cuda")
        flag = ctypes.c_int(-1)
        _lib.check(lib.msd_conv3x3_stats_nhwc_bf16(_lib.ptr(xh), _lib.ptr(wp), _lib.ptr(b), _lib.ptr(y), _lib.ptr(sums), B, H, W, C, C, 0,
                                                   ctypes.byref(flag), s))
        return y, sums, flag.value

    y, sums, flag = run()
    assert flag == fused
    assert y[:, 0].abs().max().item() == 0 and y[:, :, 0].abs().max().item() == 0 and y[:, -1].abs().max().item() == 0 and y[:, :, -1].abs().max().item() == 0
    ref = F.conv2d(x.cuda(), _bf(w).float().cuda(), b, padding=1)
    _close(_interior(y).permute(0, 3, 1, 2), ref, 1e-2)
    yd = y.double().reshape(-1, C)
    want = torch.stack([yd.sum(0), (yd * yd).sum(0)])
    scale = torch.stack([yd.abs().sum(0), (yd * yd).sum(0)]) + 1e-12
    assert ((sums - want).abs() / scale).max().item() < 2e-6          # fp32 partial sums over <= a few thousand rows per warp
    y2, sums2, _ = run()
    assert torch.equal(y, y2) and (sums - sums2).abs().max().item() <= 1e-9 * scale.max().item()
    # and the statistics BatchNorm derives from them
    gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    stats = torch.empty(4, C, device="cuda")
    _lib.check(lib.msd_bn_finalize(_lib.ptr(sums), float(B * H * W), C, _lib.ptr(gamma), _lib.ptr(beta), 1e-5, 0.1, _lib.ptr(rm), _lib.ptr(rv),
                                   _lib.ptr(stats), s))
    yi = _interior(y).float().reshape(-1, C)
    mean, var = yi.mean(0), yi.var(0, unbiased=False)
    assert torch.allclose(stats[2], mean, atol=1e-5, rtol=1e-4) and torch.allclose(stats[3], (var + 1e-5).rsqrt(), rtol=2e-4)
    assert torch.allclose(rm, 0.1 * mean, atol=1e-5, rtol=1e-4)


@pytest.mark.parametrize("M,N,K", [(300, 256, 128), (4608, 3072, 768), (9216, 3072, 768), (1000, 768, 3072), (77, 96, 64), (2500, 192, 768)])
def test_mlp_gelu_folded_into_the_gemms(M, N, K):
    """fc1 with the pre-activation copy and the fc2 data gradient times gelu'(pre) (encoder.py:160-166 under autograd): both fused
    epilogues reproduce the GEMM + elementwise-kernel pairs they replace bit for bit (single-CTA, pair and narrow tiles)."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + N)
    a = _bf(torch.randn(M, K, generator=g)).cuda()
    w = _bf(torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    s = _lib.stream_ptr()
    pre0 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    h0 = torch.empty_like(pre0)
    _lib.check(lib.msd_linear_bf16(_lib.ptr(a), _lib.ptr(w), _lib.ptr(b), _lib.ptr(pre0), None, M, N, K, 0, 0, 0, s))
    _lib.check(lib.msd_gelu_fwd_bf16(_lib.ptr(pre0), _lib.ptr(h0), M * N, s))
    pre1, h1 = torch.full_like(pre0, 7.0), torch.full_like(pre0, 7.0)
    _lib.check(lib.msd_linear_gelu_pre_bf16(_lib.ptr(a), _lib.ptr(w), _lib.ptr(b), _lib.ptr(h1), _lib.ptr(pre1), M, N, K, s))
    assert torch.equal(pre0, pre1) and torch.equal(h0, h1)
    _close(h1, F.gelu(a.float() @ w.float().t() + b), 1e-2)
    # data gradient: dh = (dy . W2) * gelu'(pre); here `a` plays dy, `w` W2^T, and the pre-activation is an independent tensor
    z = _bf(torch.randn(M, N, generator=g) * 1.5).cuda()
    d0 = torch.empty_like(pre0)
    _lib.check(lib.msd_linear_bf16(_lib.ptr(a), _lib.ptr(w), None, _lib.ptr(d0), None, M, N, K, 0, 0, 0, s))
    _lib.check(lib.msd_gelu_bwd_bf16(_lib.ptr(d0), _lib.ptr(z), _lib.ptr(d0), M * N, s))
    d1 = torch.full_like(pre0, 7.0)
    _lib.check(lib.msd_linear_dgelu_bf16(_lib.ptr(a), _lib.ptr(w), None, _lib.ptr(z), _lib.ptr(d1), M, N, K, s))
    assert torch.equal(d0, d1)
    zr = z.float().requires_grad_(True)
    F.gelu(zr).backward(a.float() @ w.float().t())
    _close(d1, zr.grad, 1.5e-2)


def test_gelu_bwd():
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    h = _bf(torch.randn(1000, 3072, generator=g) * 2).float().requires_grad_(True)
    dh = _bf(torch.randn(1000, 3072, generator=g)).float()
    F.gelu(h).backward(dh)
    out = torch.empty(1000, 3072, device="cuda", dtype=torch.bfloat16)
    dh_d, h_d = _bf(dh).cuda(), _bf(h.detach()).cuda()
    _lib.check(lib.msd_gelu_bwd_bf16(_lib.ptr(dh_d), _lib.ptr(h_d), _lib.ptr(out), 1000 * 3072, _lib.stream_ptr()))
    _close(out.cpu(), h.grad, 1e-2)
    # the variant that also leaves the column sums of its result (fc1's bias gradient), in place, odd and even row counts
    for rows, N in ((1000, 3072), (777, 3072), (33, 256)):
        a, z = dh_d[:rows, :N].contiguous(), h_d[:rows, :N].contiguous()
        want = torch.empty_like(a)
        _lib.check(lib.msd_gelu_bwd_bf16(_lib.ptr(a), _lib.ptr(z), _lib.ptr(want), rows * N, _lib.stream_ptr()))
        cs = torch.full((N,), 2.0, device="cuda")
        _lib.check(lib.msd_gelu_bwd_colsum_bf16(_lib.ptr(a), _lib.ptr(z), _lib.ptr(a), _lib.ptr(cs), rows, N, _lib.stream_ptr()))
        assert torch.equal(a, want)
        ref = want.double().sum(0)
        assert ((cs.double() - 2.0 - ref).abs().max() <= 1e-5 * want.double().abs().sum(0).max() + 1e-4).item()


@pytest.mark.parametrize("M,f32_grad,accumulate,frozen", [(300, False, True, False), (1000, True, False, False), (77, False, False, True)])
def test_layernorm_bwd(M, f32_grad, accumulate, frozen):
    lib = _lib.load()
    D = 768
    g = torch.Generator().manual_seed(M)
    x = (torch.randn(M, D, generator=g) * 2 + 0.5).requires_grad_(True)
    gamma = (1 + 0.3 * torch.randn(D, generator=g)).requires_grad_(True)
    beta = torch.zeros(D, requires_grad=True)
    dy = torch.randn(M, D, generator=g)
    if not f32_grad:
        dy = _bf(dy).float()
    F.layer_norm(x, (D,), gamma, beta, 1e-6).backward(dy)
    g0 = torch.randn(M, D, generator=g)
    g_d = g0.clone().cuda()
    gbf = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    dgam, dbet = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    dy_d = dy.cuda() if f32_grad else _bf(dy).cuda()
    x_d, gamma_d = x.detach().cuda(), gamma.detach().cuda()
    _lib.check(lib.msd_layernorm_bwd(_lib.ptr(dy_d), int(f32_grad), _lib.ptr(x_d), _lib.ptr(gamma_d), 1e-6,
                                     _lib.ptr(g_d), int(accumulate), _lib.ptr(gbf), None if frozen else _lib.ptr(dgam),
                                     None if frozen else _lib.ptr(dbet), M, D, _lib.stream_ptr()))
    want = x.grad + (g0 if accumulate else 0)
    _close(g_d.cpu(), want, 1e-4)
    _close(gbf.cpu(), want, 1e-2)
    if not frozen:
        _close(dgam.cpu(), gamma.grad, 1e-4)
        _close(dbet.cpu(), beta.grad, 1e-4)
    # the variant that also accumulates the column sums of the bf16 copy (the bias gradient of the Linear in front of the LayerNorm)
    g2 = g0.clone().cuda()
    gbf2 = torch.empty_like(gbf)
    dgam2, dbet2 = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    cs = torch.full((D,), 3.0, device="cuda")
    _lib.check(lib.msd_layernorm_bwd_colsum(_lib.ptr(dy_d), int(f32_grad), _lib.ptr(x_d), _lib.ptr(gamma_d), 1e-6, _lib.ptr(g2), int(accumulate),
                                            _lib.ptr(gbf2), None if frozen else _lib.ptr(dgam2), None if frozen else _lib.ptr(dbet2),
                                            _lib.ptr(cs), M, D, _lib.stream_ptr()))
    assert torch.equal(g2, g_d) and torch.equal(gbf2, gbf)
    ref_cs = gbf.double().sum(0)
    assert ((cs.double() - 3.0 - ref_cs).abs().max() <= 1e-4 * gbf.double().abs().sum(0).max() + 1e-5).item()
    if not frozen:
        _close(dgam2.cpu(), gamma.grad, 1e-4)


@pytest.mark.parametrize("B,Hin,padded,C", [(2, 6, True, 48), (3, 12, False, 384), (2, 5, True, 96)])
def test_convt_unshuffle(B, Hin, padded, C):
    lib = _lib.load()
    g = torch.Generator().manual_seed(C)
    dy = _bf(torch.randn(B, 2 * Hin + 2, 2 * Hin + 2, C, generator=g)).cuda()
    ph = Hin + 2 if padded else Hin
    dys = torch.full((B * ph * ph, 4 * C), 7.0, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.msd_convt_unshuffle_bf16(_lib.ptr(dy), _lib.ptr(dys), B, Hin, Hin, int(padded), C, _lib.stream_ptr()))
    got = dys.view(B, ph, ph, 4, C)
    inner = got[:, 1:-1, 1:-1] if padded else got
    for q in range(4):
        di, dj = q >> 1, q & 1
        ref = dy[:, 1 + di:1 + di + 2 * Hin:2, 1 + dj:1 + dj + 2 * Hin:2, :]
        assert torch.equal(inner[:, :, :, q, :], ref)
    if padded:
        assert got[:, 0].abs().max().item() == 0 and got[:, :, -1].abs().max().item() == 0


def test_head1x1_bwd():
    lib = _lib.load()
    B, H, W, Cin, Cout = 2, 16, 12, 48, 13
    g = torch.Generator().manual_seed(9)
    dy = torch.randn(B, Cout, H, W, generator=g)
    w = torch.randn(Cout, Cin, generator=g) * 0.2
    dh = torch.full((B, H + 2, W + 2, Cin), 7.0, device="cuda", dtype=torch.bfloat16)
    dyt = torch.full((B * (H + 2) * (W + 2), 16), 7.0, device="cuda", dtype=torch.bfloat16)
    dy_d, w_d = dy.cuda(), w.cuda()
    _lib.check(lib.msd_head1x1_bwd(_lib.ptr(dy_d), _lib.ptr(w_d), _lib.ptr(dh), _lib.ptr(dyt), B, H, W, Cin, Cout, _lib.stream_ptr()))
    ref = torch.einsum("bohw,oc->bhwc", dy, w)
    _close(_interior(dh).cpu(), ref, 1e-2)
    assert dh[:, 0].abs().max().item() == 0
    t = dyt.view(B, H + 2, W + 2, 16)
    _close(t[:, 1:-1, 1:-1, :Cout].cpu(), dy.permute(0, 2, 3, 1), 1e-2)
    assert t[..., Cout:].abs().max().item() == 0 and t[:, 0].abs().max().item() == 0


def test_attention_bwd():
    lib = _lib.load()
    B, T, heads = 3, 144, 12
    E = heads * 64
    g = torch.Generator().manual_seed(11)
    qkv = _bf(torch.randn(B * T, 3 * E, generator=g)).float().requires_grad_(True)
    dout = _bf(torch.randn(B * T, E, generator=g)).float()
    t = qkv.reshape(B, T, 3, heads, 64).permute(2, 0, 3, 1, 4)
    att = torch.softmax((t[0] @ t[1].transpose(-2, -1)) * 0.125, dim=-1) @ t[2]
    att.transpose(1, 2).reshape(B * T, E).backward(dout)
    dqkv = torch.full((B * T, 3 * E), 7.0, device="cuda", dtype=torch.bfloat16)
    qkv_d, dout_d = _bf(qkv.detach()).cuda(), _bf(dout).cuda()
    _lib.check(lib.msd_attention_bwd_bf16(_lib.ptr(qkv_d), _lib.ptr(dout_d), _lib.ptr(dqkv), B, T, heads, _lib.stream_ptr()))
    for i, name in enumerate(("dq", "dk", "dv")):
        _close(dqkv[:, i * E:(i + 1) * E].cpu(), qkv.grad[:, i * E:(i + 1) * E], 2e-2)


# ----------------------------------------------------------------------------------------------- whole training step
def _close_rms(got, ref, rel_rms=6e-2, rel_max=0.1):
    """Training-mode activations are O(1) after every BatchNorm, so bf16 rounding (two per layer: raw conv output and
    activation) accumulates to a few percent over the 40 decoder tensors (measured 3.5% RMS at random init; layer-by-layer
    growth is smooth, tools/train_debug.py).  Judged relative to the reference's own RMS / max."""
    e = (got.float() - ref.float())
    rms, ref_rms = e.pow(2).mean().sqrt().item(), ref.float().pow(2).mean().sqrt().item()
    assert rms <= rel_rms * ref_rms, f"rms err {rms:.4g} vs ref rms {ref_rms:.4g}"
    assert e.abs().max().item() <= rel_max * ref.abs().max().item(), f"max err {e.abs().max().item():.4g} vs ref max {ref.abs().max().item():.4g}"


def _train_case(B=2, seed=42):
    import recipes
    from msd_b200.models import SatMAERestoration
    torch.manual_seed(seed)
    model = SatMAERestoration(recipes.base_config(), block_layout="fallback")      # blocks 0-5 frozen (configs/base.yaml:31)
    sd = recipes.perturb_decoder_stats({k: v.clone() for k, v in model.state_dict().items()})
    model.load_state_dict(sd)
    x = recipes.model_input(B, seed=77)
    target = torch.randn(B, 13, 192, 192, generator=torch.Generator().manual_seed(5)) * 0.5
    return model, sd, x, target


def test_train_forward_and_backward_vs_autograd():
    from oracle import model as omodel
    model, sd, x, target = _train_case()
    frozen = {n for n, p in model.named_parameters() if not p.requires_grad}
    # ---- oracle: torch fp32 autograd, training-mode BatchNorm
    ref = {k: v.clone() for k, v in sd.items()}
    for k, v in ref.items():
        if v.is_floating_point() and "running_" not in k and k not in frozen:
            v.requires_grad_(True)
    y_ref = omodel.restoration_forward_train(ref, x)
    loss_ref = ((y_ref - target) ** 2).mean()
    loss_ref.backward()
    # ---- CUDA path
    model.cuda().train()
    y = model(x.cuda())
    assert y.requires_grad
    loss = ((y - target.cuda()) ** 2).mean()
    loss.backward()
    _close_rms(y.detach().cpu(), y_ref.detach())
    # running statistics follow nn.BatchNorm2d (momentum 0.1, unbiased variance)
    for k, v in model.state_dict().items():
        if "running_" in k:
            assert torch.allclose(v.cpu(), ref[k], rtol=2e-2, atol=2e-3), k
        if k.endswith("num_batches_tracked"):
            assert int(v) == 1
    worst = []
    for n, p in model.named_parameters():
        if n in frozen:
            assert p.grad is None, n
            continue
        assert p.grad is not None, n
        gref = ref[n].grad
        scale = gref.abs().max().item()
        e = (p.grad.cpu() - gref).abs().max().item()
        if n.endswith(("upsample.0.bias", "head.0.bias")):      # a bias in front of BatchNorm: exactly zero gradient
            assert p.grad.abs().max().item() == 0 and scale < 1e-6
            continue
        rms = (p.grad.cpu() - gref).pow(2).mean().sqrt().item() / (gref.pow(2).mean().sqrt().item() + 1e-20)
        worst.append((rms, e / (scale + 1e-20), n, scale))
    worst.sort(reverse=True)
    print("worst gradient errors (rel rms, rel max, name, |grad|max):", *worst[:12], sep="\n  ")
    import os
    if os.environ.get("MSD_TRAIN_VERBOSE"):
        for w_ in sorted(worst, key=lambda t: t[2]):
            print(f"   {w_[2]:60s} rms {w_[0]:.4f} max {w_[1]:.4f} scale {w_[3]:.3g}")
    # bf16 activations/gradients vs fp32 autograd: torch's own bf16 autocast on the same model, input and loss deviates from
    # fp32 by 0.7% (head) ... 5.6% (stage 2) ... 11% (stage 0 and the whole encoder) relative RMS -- the same profile this path
    # shows (0.6% ... 4.7% ... 9-10%); recorded in DESIGN.md.  The bound below is that mixed-precision envelope.
    assert worst[0][0] <= 0.15 and max(w[1] for w in worst) <= 0.25, worst[:8]


def test_train_mode_no_grad_forward_matches_and_eval_still_works():
    from oracle import model as omodel
    model, sd, x, _ = _train_case(B=2, seed=7)
    model.cuda().train()
    with torch.no_grad():
        y = model(x.cuda())
    ref = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        y_ref = omodel.restoration_forward_train(ref, x)
    _close_rms(y.cpu(), y_ref)
    model.eval()
    with torch.no_grad():
        y2 = model(x.cuda())
    y2_ref = omodel.restoration_forward(ref, x)                 # eval forward with the UPDATED running statistics
    assert (y2.cpu() - y2_ref).abs().max().item() <= 2e-2


def test_train_step_at_larger_batch_vs_oracle_on_device():
    """Same comparison at a batch size whose tensors span many CTAs / grid-stride iterations (the CPU oracle is too slow there):
    the fp32 oracle is evaluated with torch on the GPU (TF32 off)."""
    import recipes
    from msd_b200.models import SatMAERestoration
    from oracle import model as omodel
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    B = 24
    torch.manual_seed(42)
    model = SatMAERestoration(recipes.base_config(), block_layout="fallback")
    sd = {k: v.clone().cuda() for k, v in model.state_dict().items()}
    clean, noisy = recipes.eurosat_tiles(B, seed=3)
    clean, noisy = clean.cuda(), noisy.cuda()
    model.cuda().train()
    frozen = {n for n, p in model.named_parameters() if not p.requires_grad}
    ref = {k: v.clone() for k, v in sd.items()}
    for k, v in ref.items():
        if v.is_floating_point() and "running_" not in k and k not in frozen:
            v.requires_grad_(True)
    y_ref = omodel.restoration_forward_train(ref, noisy)
    ((y_ref - clean) ** 2).mean().backward()
    y = model(noisy)
    ((y - clean) ** 2).mean().backward()
    _close_rms(y.detach(), y_ref.detach(), rel_rms=0.1, rel_max=0.2)
    errs = []
    for n, p in model.named_parameters():
        if n in frozen or n.endswith(("upsample.0.bias", "head.0.bias")):
            continue
        g, r = p.grad, ref[n].grad
        errs.append(((g - r).pow(2).mean().sqrt() / (r.pow(2).mean().sqrt() + 1e-30)).item())
    assert max(errs) <= 0.2 and sorted(errs)[len(errs) // 2] <= 0.15, (max(errs), sorted(errs)[len(errs) // 2])


def test_train_step_error_is_the_bf16_storage_error_tensor_by_tensor():
    """The fp32-autograd comparisons above carry a flat mixed-precision envelope (up to 15-20 % relative RMS in the encoder).  This test
    calibrates it: oracle/model_storage.py is the SAME fp32 torch graph with the library's bf16 storage points written in (every
    activation / raw conv output / inter-layer gradient kept in bf16, bf16 operand copies of the weights), evaluated on the GPU with
    TF32 off.  Its distance from plain fp32 autograd is what bf16 storage costs, per tensor, for this model and input -- two
    independently rounded runs do not agree with EACH OTHER any better (rounding decisions decorrelate), so the meaningful statement
    is: for every gradient tensor the library is no further from fp32 autograd than the storage emulation is, within 35 %
    (measured: ratio 0.90 ... 1.29, median 1.04; profiles/r2_train_storage_check.txt).  A wrong kernel, a missing term or a
    mis-scaled gradient shows up as a ratio far above 1 in the tensors behind it."""
    import recipes
    from msd_b200.models import SatMAERestoration
    from oracle import model as omodel, model_storage
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    B = 8
    torch.manual_seed(11)
    model = SatMAERestoration(recipes.base_config(), block_layout="fallback")
    sd = recipes.perturb_decoder_stats({k: v.clone() for k, v in model.state_dict().items()})
    model.load_state_dict(sd)
    sd = {k: v.cuda() for k, v in sd.items()}
    clean, noisy = recipes.eurosat_tiles(B, seed=9)
    clean, noisy = clean.cuda(), noisy.cuda()
    model.cuda().train()
    frozen = {n for n, p in model.named_parameters() if not p.requires_grad}

    def oracle(fwd):
        ref = {k: v.clone() for k, v in sd.items()}
        for k, v in ref.items():
            if v.is_floating_point() and "running_" not in k and k not in frozen:
                v.requires_grad_(True)
        y_ref = fwd(ref, noisy)
        ((y_ref - clean) ** 2).mean().backward()
        return y_ref.detach(), ref

    y_st, ref_st = oracle(model_storage.restoration_forward_train)
    y_32, ref_32 = oracle(omodel.restoration_forward_train)
    y = model(noisy)
    ((y - clean) ** 2).mean().backward()

    def rel(a, b):
        return ((a - b).pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-30)).item()

    assert rel(y.detach(), y_32) <= 1.35 * rel(y_st, y_32) + 2e-3, (rel(y.detach(), y_32), rel(y_st, y_32))
    rows = []
    for n, p in model.named_parameters():
        if n in frozen or n.endswith(("upsample.0.bias", "head.0.bias")):
            continue
        rows.append((rel(p.grad, ref_32[n].grad), rel(ref_st[n].grad, ref_32[n].grad), n))
    bad = [r for r in rows if r[0] > 1.35 * r[1] + 2e-3]
    assert not bad, bad[:8]
    ratios = sorted(a / (b + 1e-9) for a, b, _ in rows)
    assert ratios[len(ratios) // 2] <= 1.15, ratios[len(ratios) // 2]
    # and the emulation itself shows the profile the flat envelope was derived from: ~1 % at the head, ~10 % and more in the encoder
    assert max(b for _, b, _ in rows) >= 0.05


def test_fused_adamw_with_clip_matches_torch():
    from msd_b200.training import FusedAdamW
    g = torch.Generator().manual_seed(0)
    shapes = [(768, 3328), (13, 48, 1, 1), (7,), (384, 384, 3, 3), (1, 144, 768), (5, 3)]
    ref = [torch.nn.Parameter(torch.randn(*s, generator=g).cuda()) for s in shapes]
    mine = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    groups = lambda ps: [{"params": ps[:3], "lr": 1e-3, "weight_decay": 0.05}, {"params": ps[3:], "lr": 1e-2, "weight_decay": 0.0}]
    o_ref = torch.optim.AdamW(groups(ref), betas=(0.9, 0.999), eps=1e-8)
    o_mine = FusedAdamW(groups(mine), betas=(0.9, 0.999), eps=1e-8, max_grad_norm=1.0)
    for step in range(4):
        for a, b in zip(ref, mine):
            gr = torch.randn(a.shape, generator=g).cuda() * (0.01 if step == 3 else 1.0)    # last step: norm < 1, clip inactive
            a.grad, b.grad = gr.clone(), gr.clone()
        norm = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        o_ref.step(); o_mine.step()
        assert abs(o_mine.last_grad_sqnorm.item() ** 0.5 - norm.item()) <= 1e-4 * norm.item()
        for a, b in zip(ref, mine):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    sd = o_mine.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}


def test_fused_adamw_per_parameter_step_and_skipped_nonfinite_step():
    """torch.optim.AdamW keeps a step count per parameter: one that had no gradient for a while (blocks unfrozen mid-run) carries
    its own bias correction.  A non-finite gradient norm skips the update instead of poisoning the weights."""
    from msd_b200.training import FusedAdamW
    g = torch.Generator().manual_seed(1)
    ref = [torch.nn.Parameter(torch.randn(40, 9, generator=g).cuda()) for _ in range(2)]
    mine = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    o_ref = torch.optim.AdamW(ref, lr=1e-2, weight_decay=0.05)
    o_mine = FusedAdamW(mine, lr=1e-2, weight_decay=0.05)
    for step in range(5):
        for i, (a, b) in enumerate(zip(ref, mine)):
            if i == 1 and step < 2:              # the second parameter joins at step 2
                a.grad = b.grad = None
                continue
            gr = torch.randn(a.shape, generator=g).cuda()
            a.grad, b.grad = gr.clone(), gr.clone()
        o_ref.step(); o_mine.step()
        for a, b in zip(ref, mine):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    assert int(o_mine.state[mine[0]]["step"]) == 5 and int(o_mine.state[mine[1]]["step"]) == 3
    v0 = mine[0]._version
    clipped = FusedAdamW(mine, lr=1e-2, max_grad_norm=1.0)
    before = [p.detach().clone() for p in mine]
    for p in mine:
        p.grad = torch.full_like(p, float("inf"))
    clipped.step()
    assert all(torch.equal(a, b) for a, b in zip(before, mine))
    assert mine[0]._version > v0                 # the step still tells autograd the storage was handed to a kernel


def test_packed_weights_follow_the_fused_optimizer():
    """FusedAdamW writes parameters through raw pointers; the model's packed copies (bf16 weights, transposes, tap-flipped dgrad
    weights, replayed from a CUDA graph) are keyed on the parameters' versions, so the optimizer has to bump them.  After a few
    steps the training model must compute exactly what a FRESH model loaded from its state dict computes, and must track the
    same run driven by torch.optim.AdamW + clip_grad_norm_."""
    import recipes
    from msd_b200.models import SatMAERestoration
    from msd_b200.training import CombinedLoss, FusedAdamW
    torch.manual_seed(42)
    a = SatMAERestoration(recipes.base_config(), block_layout="fallback").cuda().train()
    b = SatMAERestoration(recipes.base_config(), block_layout="fallback").cuda().train()
    b.load_state_dict(a.state_dict())
    p0 = {n: p.detach().clone() for n, p in a.named_parameters()}
    clean, noisy = recipes.eurosat_tiles(4, seed=5)
    clean, noisy = clean.cuda(), noisy.cuda()
    crit = CombinedLoss(1.0, 0.1)
    oa = FusedAdamW(a.get_optimizer_param_groups(1e-3, 0.05), max_grad_norm=1.0)
    ob = torch.optim.AdamW(b.get_optimizer_param_groups(1e-3, 0.05))
    la, lb = [], []
    for _ in range(4):
        la.append(crit(a(noisy), clean)); la[-1].backward(); oa.step(); oa.zero_grad(set_to_none=True)
        lb.append(crit(b(noisy), clean)); lb[-1].backward()
        torch.nn.utils.clip_grad_norm_(b.parameters(), 1.0)
        ob.step(); ob.zero_grad(set_to_none=True)
    la, lb = [float(v.detach()) for v in la], [float(v.detach()) for v in lb]
    assert la[-1] < la[0], la                                    # lr 1e-3 on a random-init decoder: the loss moves
    # same run under torch.optim.AdamW: identical first step, then Adam's sign-like early updates amplify the fp32-atomic noise
    # of the weight gradients a little more every step
    assert abs(la[0] - lb[0]) <= 1e-5 * abs(lb[0]) and abs(la[1] - lb[1]) <= 2e-3 * abs(lb[1]), (la, lb)
    assert all(abs(x - y) <= 3e-2 * abs(y) for x, y in zip(la, lb)), (la, lb)
    # The UPDATES agree overall.  Not tightly: the two runs differ by fp32-atomic summation order in the weight gradients, a
    # 1e-7 difference in an fp32 weight can flip its bf16 rounding (2e-4 on a 0.05 weight, a fifth of one lr = 1e-3 Adam step),
    # and four large early steps through BatchNorm amplify that -- 15 % of the update norm is the measured run-to-run level.
    num = den = 0.0
    for (n, p), q in zip(a.named_parameters(), b.parameters()):
        if p.requires_grad:
            da, db = p.detach() - p0[n], q.detach() - p0[n]
            num += (da - db).pow(2).sum().item()
            den += db.pow(2).sum().item()
    assert (num / den) ** 0.5 <= 0.35, (num / den) ** 0.5
    moved = sum(int((q.detach() - p0[n]).abs().max().item() > 0) for n, q in b.named_parameters() if q.requires_grad)
    assert moved > 100          # (conv / ConvTranspose biases in front of BatchNorm have exactly zero gradient and stay put)
    fresh = SatMAERestoration(recipes.base_config(), block_layout="fallback").cuda().train()
    fresh.load_state_dict(a.state_dict())
    with torch.no_grad():
        ya, yf = a(noisy), fresh(noisy)
    assert torch.equal(ya, yf), "the training forward ran on stale packed weights"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL gradient all-reduce)")
def test_data_parallel_backward_averages_gradients(tmp_path):
    """config C4 on two ranks: with enable_data_parallel() the backward all-reduces its gradient arena group by group on a
    communication stream (behind msd_model_grad_group_wait) -- the result is the mean of the ranks' own gradients."""
    import json, os, subprocess, sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    out = tmp_path / "dp.json"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29551", str(root / "tests" / "dp_driver.py"), str(out)], cwd=root, capture_output=True, text=True,
                       timeout=900, env=dict(os.environ, PYTHONPATH=str(root)))
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    res = json.loads(out.read_text())
    assert res["world"] == 2 and res["tensors"] > 100
    assert res["worst_rel"] <= 2e-3, res


def test_repack_kernel_matches_the_torch_packing_functions():
    """msd_repack_weights (one launch after each optimizer step) against the torch restatements of every packed layout:
    bf16 cast, Linear transpose, 3x3 conv tap-major (+ zero K padding), its tap-flipped data-gradient form (+ identity tap),
    ConvTranspose rows (di, dj, o), its transpose, the per-sub-pixel bias -- bit for bit."""
    import ctypes
    from msd_b200.models.satmae_restoration import pack_conv3x3_dgrad, pack_conv3x3_raw, pack_convt2x2_raw
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    dev = "cuda"
    ops, want, got = [], [], []

    def add(src, kind, ref, dtype=torch.bfloat16, **kw):
        dst = torch.full(ref.shape, float("nan"), dtype=dtype, device=dev)
        op = _lib.RepackOp()
        op.src, op.dst, op.kind = src.data_ptr(), dst.data_ptr(), kind
        for k, v in kw.items():
            setattr(op, k, v)
        ops.append(op); want.append(ref); got.append(dst)

    lin = torch.randn(2304, 768, generator=g).to(dev)
    add(lin, _lib.REPACK_CAST, lin.to(torch.bfloat16), d0=2304, d1=768)
    add(lin, _lib.REPACK_TRANSPOSE, lin.t().contiguous().to(torch.bfloat16), d0=2304, d1=768)
    keep = [lin]
    for c in (48, 96, 384):
        w = torch.randn(c, c, 3, 3, generator=g).to(dev)
        kpad = (c + 63) // 64 * 64
        keep.append(w)
        add(w, _lib.REPACK_CONV3, pack_conv3x3_raw(w), d0=c, d1=c, kpad=kpad)
        add(w, _lib.REPACK_CONV3_DGRAD, pack_conv3x3_dgrad(w, with_identity=True), d0=c, d1=c, kpad=kpad, flag=1)
        add(w, _lib.REPACK_CONV3_DGRAD, pack_conv3x3_dgrad(w), d0=c, d1=c, kpad=kpad, flag=0)
    for cin, cout in ((768, 384), (96, 48)):
        ct, cb = torch.randn(cin, cout, 2, 2, generator=g).to(dev), torch.randn(cout, generator=g).to(dev)
        keep.extend([ct, cb])
        w_up, b_up, wT_up = pack_convt2x2_raw(ct, cb)
        add(ct, _lib.REPACK_CONVT, w_up, d0=cin, d1=cout, kpad=(cin + 63) // 64 * 64)
        add(ct, _lib.REPACK_CONVT_T, wT_up, d0=cin, d1=cout)
        add(cb, _lib.REPACK_REPEAT_F32, b_up, dtype=torch.float32, d0=cout, flag=4)
    table = (_lib.RepackOp * len(ops))(*ops)
    _lib.check(lib.msd_repack_weights(table, len(ops), _lib.stream_ptr()))
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(got, want)):
        assert a.shape == b.shape and torch.equal(a, b.to(a.device)), f"op {i} (kind {ops[i].kind})"
