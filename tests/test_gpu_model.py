"""GPU parity of the tensor-core path: GEMM / implicit conv (tcgen05 and the legacy mma.sync kernel), LayerNorm,
attention, and the whole SatMAERestoration forward vs the fp32 oracle and the reference-generated golden vectors.

Tolerances (BASELINE.json): restored pixels <= 2e-2 max-abs and <= 1e-3 RMS vs the FP32 reference under BF16;
patch indexing exact (checked through the bit-exact patchify test and structured-input GEMM tests)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import emulate
import gpu_helpers as gh
import recipes
from msd_b200 import _lib
from msd_b200.models import SatMAERestoration
from msd_b200.models.satmae_restoration import pack_conv3x3, pack_convt2x2, pack_ln_folded_linear
from oracle import model as omodel

pytestmark = pytest.mark.gpu
BACKENDS = [pytest.param(0, id="tcgen05"), pytest.param(1, id="mma_sync")]


def _bf(t):
    return t.to(torch.bfloat16)


def _close(got, ref, rel=2e-2):
    """bf16 operands, fp32 accumulate: error budget relative to the largest reference magnitude."""
    scale = ref.abs().max().item() + 1e-6
    err = (got.float() - ref.float()).abs().max().item()
    assert err <= rel * scale, f"max err {err:.4g} vs scale {scale:.4g}"


# ----------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1152, 768, 768), (288, 2304, 768), (1296, 3072, 768), (200, 768, 3072),
                                   (144, 768, 3328), (130, 48, 128), (576, 16, 64), (1000, 96, 192), (384, 192, 320),
                                   # >= 2048 rows with a 256- or 192-wide tile: the cta_group::2 pair kernel (ragged last pair tile too)
                                   (2304, 768, 768), (4000, 2304, 768), (2177, 3072, 768), (2560, 768, 3072), (2500, 192, 320)])
def test_linear_vs_fp32(backend, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    a = _bf(torch.randn(M, K, generator=g)).cuda()
    w = _bf(torch.randn(N, K, generator=g) * K ** -0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    ref = a.float() @ w.float().t() + bias
    _close(gh.linear(a, w, bias, out_f32=True, backend=backend), ref, 2e-3)
    _close(gh.linear(a, w, bias, backend=backend), ref)
    _close(gh.linear(a, w, bias, act=1, backend=backend), F.gelu(ref))
    resid = torch.randn(M, N, generator=g).cuda()
    _close(gh.linear(a, w, bias, resid=resid, backend=backend), ref + resid, 2e-3)
    _close(gh.linear(a, w, None, out_f32=True, backend=backend), ref - bias, 2e-3)


@pytest.mark.parametrize("backend", BACKENDS)
def test_linear_exact_on_small_integers(backend):
    """integer-valued operands: every product and partial sum is exact in bf16 x bf16 -> fp32, so any indexing or
    descriptor mistake (swizzle, K advance, tile origin) shows up as an exact mismatch."""
    g = torch.Generator().manual_seed(0)
    for M, N, K in ((400, 320, 256), (2400, 768, 320)):          # single-CTA tiles; CTA-pair tiles (M >= 2048, BN = 256)
        a = torch.randint(-3, 4, (M, K), generator=g).float()
        w = torch.randint(-3, 4, (N, K), generator=g).float()
        out = gh.linear(_bf(a).cuda(), _bf(w).cuda(), None, out_f32=True, backend=backend)
        assert torch.equal(out.cpu(), a @ w.t())


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("cin,cout,H,W,B", [(48, 48, 20, 12, 2), (96, 96, 10, 14, 3), (192, 192, 8, 8, 1), (384, 384, 24, 24, 1),
                                            (64, 16, 5, 7, 2), (192, 192, 30, 33, 2), (384, 384, 24, 24, 4), (48, 48, 40, 41, 2), (96, 96, 30, 31, 3)])   # last four: CTA-pair tiles
def test_conv3x3_vs_torch(backend, cin, cout, H, W, B):
    torch.manual_seed(cin + H)
    conv = nn.Conv2d(cin, cout, 3, padding=1, bias=False)
    bn = nn.BatchNorm2d(cout).eval()
    bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 1.5); bn.weight.data.normal_(1, 0.1); bn.bias.data.normal_(0, 0.1)
    x = _bf(torch.randn(B, cin, H, W)).float()
    skip = _bf(torch.randn(B, cout, H, W)).float()
    with torch.no_grad():
        w, b = pack_conv3x3(conv, bn)
        xh, sh = _bf(emulate.halo_nhwc(x)).cuda(), _bf(emulate.halo_nhwc(skip)).cuda()
        ref_plain = emulate.conv3x3_rows(emulate.halo_nhwc(x), w.float(), b, H, W)
    if cin == cout:   # ResidualBlock second conv: skip connection packed as the identity tenth tap
        with torch.no_grad():
            ws, _ = pack_conv3x3(conv, bn, with_skip=True)
            ref_skip = emulate.conv3x3_rows(emulate.halo_nhwc(x), ws.float(), b, H, W, skip=emulate.halo_nhwc(skip))
        y = gh.conv3x3(xh, ws.cuda(), b.cuda(), sh, B, H, W, cin, cout, backend)
        _close(y.cpu(), ref_skip)
        assert torch.all(y[:, 0] == 0) and torch.all(y[:, -1] == 0) and torch.all(y[:, :, 0] == 0) and torch.all(y[:, :, -1] == 0)
    y = gh.conv3x3(xh, w.cuda(), b.cuda(), None, B, H, W, cin, cout, backend)
    _close(y.cpu(), ref_plain)
    with torch.no_grad():
        _close(emulate.unhalo_nchw(y.cpu().float()), F.gelu(bn(conv(x))), 3e-2)     # and against the real conv


@pytest.mark.parametrize("H,W,B", [(192, 192, 2), (6, 64, 1), (5, 128, 2), (8, 320, 1), (10, 192, 3), (2, 192, 1), (1, 256, 2), (64, 192, 9)])
def test_conv48_rolling_rows_vs_tap_major_kernel_and_torch(H, W, B, monkeypatch):
    """csrc/conv_roll.cu (the 48-channel 3x3 conv with the vertical taps stacked along N and output rows in a TMEM ring;
    blocks.py:66-85) against the tap-major tcgen05 kernel it replaces: integer-valued operands make every partial sum exact,
    so the two formulations must agree BIT FOR BIT (any slot / window / pair-strip indexing slip is an exact mismatch), with and
    without the ResidualBlock skip; then random operands against torch's conv.  Shapes: full strips, pair strips, both, H = 1 / 2
    (segments shorter than the ring), more steps than CTAs (ranges that cut strips in the middle)."""
    g = torch.Generator().manual_seed(H * 1000 + W + B)
    xi = torch.randint(-1, 2, (B, 48, H, W), generator=g).float()
    si = torch.randint(-3, 4, (B, 48, H, W), generator=g).float()
    wi = torch.randint(-1, 2, (48, 48, 3, 3), generator=g).float() * (torch.rand(48, 48, 3, 3, generator=g) < 0.3)
    bi = torch.randint(-2, 3, (48,), generator=g).float()
    kpad = 64
    wp = torch.zeros(48, 3, 3, kpad); wp[..., :48] = wi.permute(0, 2, 3, 1)
    w9 = _bf(wp.reshape(48, 9 * kpad)).cuda()
    eye = torch.zeros(48, kpad); eye[:, :48] = torch.eye(48)
    w10 = _bf(torch.cat([wp.reshape(48, 9 * kpad), eye], dim=1)).cuda()
    xh, sh, bc = _bf(emulate.halo_nhwc(xi)).cuda(), _bf(emulate.halo_nhwc(si)).cuda(), bi.cuda()
    outs = {}
    for no_roll in ("0", "1"):
        monkeypatch.setenv("MSD_NO_ROLL", no_roll)
        outs[no_roll] = (gh.conv3x3(xh, w9, bc, None, B, H, W, 48, 48, 0).clone(), gh.conv3x3(xh, w10, bc, sh, B, H, W, 48, 48, 0).clone())
    monkeypatch.setenv("MSD_NO_ROLL", "0")
    for a, b in zip(outs["0"], outs["1"]):
        assert torch.equal(a, b)
        assert torch.all(a[:, 0] == 0) and torch.all(a[:, -1] == 0) and torch.all(a[:, :, 0] == 0) and torch.all(a[:, :, -1] == 0)
    pre = F.conv2d(xi, wi, bi, padding=1)
    _close(emulate.unhalo_nchw(outs["0"][0].cpu().float()), F.gelu(pre), 1e-2)
    _close(emulate.unhalo_nchw(outs["0"][1].cpu().float()), F.gelu(pre + si), 1e-2)
    # random operands, BatchNorm folded, against the real conv
    torch.manual_seed(H + W)
    conv = nn.Conv2d(48, 48, 3, padding=1, bias=False)
    bn = nn.BatchNorm2d(48).eval()
    bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 1.5); bn.weight.data.normal_(1, 0.1); bn.bias.data.normal_(0, 0.1)
    x = _bf(torch.randn(B, 48, H, W)).float()
    skip = _bf(torch.randn(B, 48, H, W)).float()
    with torch.no_grad():
        w, b = pack_conv3x3(conv, bn)
        ws, _ = pack_conv3x3(conv, bn, with_skip=True)
        y = gh.conv3x3(_bf(emulate.halo_nhwc(x)).cuda(), w.cuda(), b.cuda(), None, B, H, W, 48, 48, 0)
        ys = gh.conv3x3(_bf(emulate.halo_nhwc(x)).cuda(), ws.cuda(), b.cuda(), _bf(emulate.halo_nhwc(skip)).cuda(), B, H, W, 48, 48, 0)
        _close(emulate.unhalo_nchw(y.cpu().float()), F.gelu(bn(conv(x))), 3e-2)
        _close(emulate.unhalo_nchw(ys.cpu().float()), F.gelu(bn(conv(x)) + skip), 3e-2)


@pytest.mark.parametrize("H,W,B,cout", [(20, 12, 2, 13), (40, 41, 3, 13), (9, 7, 1, 16), (33, 30, 2, 5),
                                        # geometries the rolling-row kernel takes (csrc/conv_roll.cu, E_CONV_HEAD): full + pair strips
                                        (192, 192, 2, 13), (6, 64, 1, 13), (8, 320, 1, 16), (10, 128, 2, 5), (64, 192, 5, 13)])
def test_head_conv_with_fused_1x1_vs_separate_kernels(H, W, B, cout):
    """decoder head (decoder.py:76-81) in one kernel: conv3x3 + BN + GELU stays in the epilogue's registers (rounded to bf16 as
    the two-kernel path stores it) and is contracted with the fp32 1x1 weights there.  Same activation as the two-kernel
    path, so the two agree to fp32 rounding of a 48-term dot product; ragged last tiles and every halo row are covered."""
    torch.manual_seed(H * W)
    conv = nn.Conv2d(48, 48, 3, padding=1, bias=True)
    bn = nn.BatchNorm2d(48).eval()
    bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 1.5); bn.weight.data.normal_(1, 0.1); bn.bias.data.normal_(0, 0.1)
    w1 = (torch.randn(cout, 48) * 0.3).cuda()
    b1 = torch.randn(cout).cuda()
    x = _bf(torch.randn(B, 48, H, W)).float()
    with torch.no_grad():
        w, b = pack_conv3x3(conv, bn)
    xh = _bf(emulate.halo_nhwc(x)).cuda()
    lib = _lib.load()
    wc, bc = w.cuda(), b.cuda()                       # named: the library only sees raw pointers
    y = torch.full((B, cout, H, W), float("nan"), device="cuda")
    _lib.check(lib.msd_conv3x3_head_nchw_f32(_lib.ptr(xh), _lib.ptr(wc), _lib.ptr(bc), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(y),
                                             B, H, W, 48, cout, _lib.stream_ptr()))
    mid = gh.conv3x3(xh, wc, bc, None, B, H, W, 48, 48, 0)              # haloed NHWC bf16, the tile the fused kernel keeps on chip
    y2 = torch.empty(B, cout, H, W, device="cuda")
    _lib.check(lib.msd_head1x1_nchw_f32(_lib.ptr(mid), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(y2), B, H, W, 48, cout, _lib.stream_ptr()))
    assert torch.isfinite(y).all()
    ref = torch.einsum("bhwc,oc->bohw", mid[:, 1:-1, 1:-1].float(), w1) + b1.view(1, -1, 1, 1)
    scale = ref.abs().max().item()
    assert (y - ref).abs().max().item() <= 2e-5 * scale
    assert (y - y2).abs().max().item() <= 2e-5 * scale
    with torch.no_grad():
        full = F.conv2d(F.gelu(bn(conv(x))), w1.cpu().view(cout, 48, 1, 1), b1.cpu())
    _close(y.cpu(), full, 3e-2)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("cin,cout,Hin,Win,padded,B", [(768, 384, 12, 12, False, 2), (96, 48, 9, 7, True, 2), (192, 96, 6, 6, True, 1)])
def test_convt_vs_torch(backend, cin, cout, Hin, Win, padded, B):
    torch.manual_seed(cout)
    ct = nn.ConvTranspose2d(cin, cout, 2, stride=2)
    bn = nn.BatchNorm2d(cout).eval()
    bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 1.5); bn.weight.data.normal_(1, 0.1); bn.bias.data.normal_(0, 0.1)
    x = _bf(torch.randn(B, cin, Hin, Win)).float()
    with torch.no_grad():
        w, b = pack_convt2x2(ct, bn)
        ref = F.gelu(bn(ct(x)))
    xin = emulate.halo_nhwc(x) if padded else x.permute(0, 2, 3, 1).contiguous()
    y = gh.convt2x2(_bf(xin).cuda(), w.cuda(), b.cuda(), B, Hin, Win, padded, cin, cout, backend)
    _close(emulate.unhalo_nchw(y.cpu().float()), ref, 3e-2)
    assert torch.all(y[:, 0] == 0) and torch.all(y[:, -1] == 0) and torch.all(y[:, :, 0] == 0) and torch.all(y[:, :, -1] == 0)


# ----------------------------------------------------------------------------------------------- LN / attention
def test_layernorm_and_pos():
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    M, D = 300, 768
    x = (torch.randn(M, D, generator=g) * 3 + 0.5).cuda()
    gam, bet = torch.randn(D, generator=g).cuda(), torch.randn(D, generator=g).cuda()
    for eps in (1e-5, 1e-6):
        out = torch.empty(M, D, dtype=torch.bfloat16, device="cuda")
        _lib.check(lib.msd_layernorm_f32_bf16(_lib.ptr(x), _lib.ptr(gam), _lib.ptr(bet), eps, _lib.ptr(out), M, D, _lib.stream_ptr()))
        ref = F.layer_norm(x, (D,), gam, bet, eps)
        assert (out.float() - ref).abs().max().item() <= 2 ** -8 * ref.abs().max().item()
    pos = torch.randn(144, D, generator=g).cuda()
    x2 = x.clone()
    _lib.check(lib.msd_layernorm_pos_f32(_lib.ptr(x2), _lib.ptr(gam), _lib.ptr(bet), 1e-5, _lib.ptr(pos), 144, M, D, _lib.stream_ptr()))
    ref = F.layer_norm(x, (D,), gam, bet, 1e-5) + pos[torch.arange(M) % 144]
    torch.testing.assert_close(x2, ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("M", [300, 1152, 2304, 2500, 4736])      # single-CTA tiles of several widths; >= 2048 rows: CTA pairs, ragged tail
def test_layernorm_folded_into_the_gemms_around_it(M):
    """encoder.py:154-176 without a LayerNorm kernel: the GEMM in front (proj / fc2 shape) leaves x fp32, bf16(x) and partial row
    sums, the GEMM behind (qkv / fc1 shape) normalises in its epilogue.  Rows get a large common offset so that a wrong mean or a
    missing `- mean * s` term cannot hide."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(M)
    D, K1 = 768, 320
    parts = lib.msd_ln_stats_parts(D, M)
    assert 0 < parts <= 16
    x0 = (torch.randn(M, D, generator=g) * 2 + torch.randn(M, 1, generator=g) * 3).cuda()
    a = _bf(torch.randn(M, K1, generator=g)).cuda()
    w = _bf(torch.randn(D, K1, generator=g) * K1 ** -0.5).cuda()
    b = torch.randn(D, generator=g).cuda()
    x = x0.clone()
    xb = torch.full((M, D), float("nan"), dtype=torch.bfloat16, device="cuda")
    stats = torch.full((M, parts, 2), float("nan"), device="cuda")
    _lib.check(lib.msd_linear_resid_stats_f32(_lib.ptr(a), _lib.ptr(w), _lib.ptr(b), _lib.ptr(x), _lib.ptr(xb), _lib.ptr(stats), parts,
                                              M, D, K1, _lib.stream_ptr()))
    ref_x = x0 + a.float() @ w.float().t() + b
    _close(x, ref_x, 2e-3)
    assert torch.equal(xb, _bf(x))                                     # the bf16 copy is the rounding of the fp32 row it wrote
    tot = stats.double().sum(1)
    assert torch.allclose(tot[:, 0], x.double().sum(1), rtol=0, atol=1e-3 * D ** 0.5)
    assert torch.allclose(tot[:, 1], x.double().pow(2).sum(1), rtol=2e-5, atol=1e-3)
    # consumer: both epilogue variants (plain, GELU), both eps values, two output widths (N = 2304 / 3072 in the model)
    gam, bet = (torch.randn(D, generator=g) * 0.3 + 1).cuda(), (torch.randn(D, generator=g) * 0.2).cuda()
    for N, act, eps in ((2304, 0, 1e-5), (3072, 1, 1e-6), (256, 0, 1e-5)):
        w2 = (torch.randn(N, D, generator=g) * D ** -0.5).cuda()
        b2 = torch.randn(N, generator=g).cuda()
        wg, s, bl = pack_ln_folded_linear(w2, b2, gam, bet)
        out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
        _lib.check(lib.msd_linear_lnfold_bf16(_lib.ptr(xb), _lib.ptr(wg), _lib.ptr(bl), _lib.ptr(s), _lib.ptr(stats), parts, eps,
                                              _lib.ptr(out), M, N, D, act, _lib.stream_ptr()))
        ref = F.linear(F.layer_norm(x, (D,), gam, bet, eps), w2, b2)
        _close(out, F.gelu(ref) if act else ref)
        # and against the un-fused kernels on the same operands (LayerNorm kernel -> bf16 -> GEMM): same error class
        xn = torch.empty(M, D, dtype=torch.bfloat16, device="cuda")
        _lib.check(lib.msd_layernorm_f32_bf16(_lib.ptr(x), _lib.ptr(gam), _lib.ptr(bet), eps, _lib.ptr(xn), M, D, _lib.stream_ptr()))
        two = gh.linear(xn, _bf(w2), b2, act=act)
        assert (out.float() - two.float()).abs().max().item() <= 3e-2 * ref.abs().max().item()


def test_layernorm_pos_with_stats():
    lib = _lib.load()
    g = torch.Generator().manual_seed(9)
    M, D, parts = 300, 768, 6
    x = (torch.randn(M, D, generator=g) * 3 + 0.5).cuda()
    gam, bet, pos = torch.randn(D, generator=g).cuda(), torch.randn(D, generator=g).cuda(), torch.randn(144, D, generator=g).cuda()
    x1, x2 = x.clone(), x.clone()
    _lib.check(lib.msd_layernorm_pos_f32(_lib.ptr(x1), _lib.ptr(gam), _lib.ptr(bet), 1e-5, _lib.ptr(pos), 144, M, D, _lib.stream_ptr()))
    xb = torch.empty(M, D, dtype=torch.bfloat16, device="cuda")
    stats = torch.full((M, parts, 2), float("nan"), device="cuda")
    _lib.check(lib.msd_layernorm_pos_stats_f32(_lib.ptr(x2), _lib.ptr(gam), _lib.ptr(bet), 1e-5, _lib.ptr(pos), 144, _lib.ptr(xb),
                                               _lib.ptr(stats), parts, M, D, _lib.stream_ptr()))
    assert torch.equal(x1, x2) and torch.equal(xb, _bf(x2))
    assert torch.all(stats[:, 1:] == 0)
    assert torch.allclose(stats[:, 0, 0].double(), x2.double().sum(1), atol=1e-2)
    assert torch.allclose(stats[:, 0, 1].double(), x2.double().pow(2).sum(1), rtol=1e-5)


@pytest.mark.parametrize("B", [1, 5, 40])
@pytest.mark.parametrize("entry", ["msd_attention_tc5_bf16", "msd_attention_bf16"])
def test_attention_vs_sdpa(entry, B):
    """tcgen05 attention (product path) and the mma.sync baseline vs fp32 SDPA; B = 40 -> 480 problems on 148 persistent
    CTAs (3-4 problems each: the stage ring, both TMEM tiles and every barrier phase wrap), rows 128..143 come from the
    shifted second tile."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(B)
    T, H, E = 144, 12, 768
    qkv = _bf(torch.randn(B * T, 3 * E, generator=g) * 1.5).cuda()
    out = torch.full((B * T, E), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(getattr(lib, entry)(_lib.ptr(qkv), _lib.ptr(out), B, T, H, _lib.stream_ptr()))
    q, k, v = qkv.float().reshape(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, E)
    _close(out, ref, 1.5e-2)


def test_attention_tc5_matches_the_baseline_kernel_on_peaked_rows():
    """Large logits (near one-hot softmax rows) and a zero head: both kernels round P to bf16 after the same max-subtracted
    exp2, so they agree to a bf16 ulp of the output."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(7)
    B, T, H, E = 3, 144, 12, 768
    qkv = torch.randn(B * T, 3 * E, generator=g) * 4.0
    qkv[:, :64] = 0.0
    qkv = _bf(qkv).cuda()
    a = torch.empty(B * T, E, dtype=torch.bfloat16, device="cuda")
    b = torch.empty_like(a)
    _lib.check(lib.msd_attention_tc5_bf16(_lib.ptr(qkv), _lib.ptr(a), B, T, H, _lib.stream_ptr()))
    _lib.check(lib.msd_attention_bf16(_lib.ptr(qkv), _lib.ptr(b), B, T, H, _lib.stream_ptr()))
    assert torch.isfinite(a.float()).all()
    assert (a.float() - b.float()).abs().max().item() <= 2 ** -7 * b.float().abs().max().item()


# ----------------------------------------------------------------------------------------------- whole model
@pytest.fixture(scope="module")
def seeded_model():
    torch.manual_seed(42)
    return SatMAERestoration(recipes.base_config(), block_layout="fallback").eval()


def _check_pixels(y, ref):
    err = (y - ref).abs()
    rms = err.pow(2).mean().sqrt().item()
    assert err.max().item() <= 2e-2 and rms <= 1e-3, f"max-abs {err.max().item():.3e}, rms {rms:.3e}"
    return err.max().item(), rms


@pytest.mark.parametrize("backend", BACKENDS)
def test_model_forward_vs_oracle_and_reference(seeded_model, golden_model, backend):
    m = seeded_model
    x = recipes.model_input(2)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    ref = omodel.restoration_forward(sd, x)
    m.gemm_backend = backend
    m._handle = None
    m.cuda()
    with torch.no_grad():
        y = m(x.cuda()).cpu()
    assert y.shape == (2, 13, 192, 192) and y.dtype == torch.float32
    mx, rms = _check_pixels(y, ref)
    # the reference's own output (golden, CPU fp32), same bounds
    _check_pixels(y[:, :, ::4, ::4], torch.from_numpy(golden_model["y_sub"]))
    # relative accuracy is what matters for trained weights: random-init outputs are tiny (std 0.015)
    assert rms <= 0.02 * ref.std().item(), f"relative rms {rms / ref.std().item():.3%}"
    m.cpu()


def test_model_with_folded_batchnorm_stats_and_timm_layout(seeded_model, golden_model):
    m = seeded_model
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    recipes.perturb_decoder_stats(sd)
    x = recipes.model_input(2)
    m2 = SatMAERestoration(recipes.base_config(), block_layout="fallback").eval()
    m2.load_state_dict(sd)
    m2.cuda()
    with torch.no_grad():
        y = m2(x.cuda()).cpu()
    ref = omodel.restoration_forward(sd, x)
    err = (y - ref).abs()
    assert err.pow(2).mean().sqrt().item() <= 0.02 * ref.std().item()
    assert (y[:, :, ::4, ::4] - torch.from_numpy(golden_model["y2_sub"])).abs().max().item() <= 0.03 * ref.abs().max().item()
    # timm key layout: same tensors, LayerNorm eps 1e-6 inside the blocks
    m3 = SatMAERestoration(recipes.base_config(), block_layout="timm").eval()
    m3.load_state_dict(sd)
    m3.cuda()
    with torch.no_grad():
        y3 = m3(x.cuda()).cpu()
    ref3 = omodel.restoration_forward(m3.cpu().state_dict(), x, ln_eps_block=1e-6)
    assert (y3 - ref3).abs().pow(2).mean().sqrt().item() <= 0.02 * ref3.std().item()


def test_rolling_conv_is_deterministic_and_pdl_changes_nothing(seeded_model, monkeypatch):
    """Two properties the bit-exact sharding tests rest on.  (1) csrc/conv_roll.cu issues its MMAs from two warps into the same
    TMEM accumulators; they sit on one SM sub-partition and hand a token over after each input row, which orders the fp32
    accumulation: repeated launches (all 148 CTAs busy) must agree bit for bit.  (2) Programmatic dependent launch
    (common.cuh msd_launch_pdl) only moves kernel prologues under the previous kernel's tail: the whole forward with MSD_NO_PDL=1
    is bit-identical."""
    torch.manual_seed(3)
    B, H, W = 40, 192, 192
    x = _bf(torch.randn(B, H + 2, W + 2, 48, device="cuda"))
    sk = _bf(torch.randn(B, H + 2, W + 2, 48, device="cuda"))
    w = _bf(torch.randn(48, 10 * 64, device="cuda") * 432 ** -0.5)
    b = torch.randn(48, device="cuda")
    for skip in (None, sk):
        y0 = gh.conv3x3(x, w, b, skip, B, H, W, 48, 48, 0).clone()
        for _ in range(6):
            assert torch.equal(gh.conv3x3(x, w, b, skip, B, H, W, 48, 48, 0), y0)
        y1 = gh.conv3x3(x[7:19].contiguous(), w, b, None if skip is None else skip[7:19].contiguous(), 12, H, W, 48, 48, 0)
        assert torch.equal(y1, y0[7:19])
    m = seeded_model
    m.gemm_backend = 0
    m._handle = None
    m.cuda()
    xin = recipes.model_input(9, seed=5).cuda()
    with torch.no_grad():
        y_pdl = m(xin).clone()
        monkeypatch.setenv("MSD_NO_PDL", "1")
        y_ser = m(xin).clone()
        monkeypatch.setenv("MSD_NO_PDL", "0")
    assert torch.equal(y_pdl, y_ser)
    m.cpu()


def test_model_batch_independence_and_ragged_tiles(seeded_model):
    """B=9 -> 1296 token rows (not a multiple of the 128-row tile) and ragged conv tiles; every image must come
    out exactly as it does in a batch of 2 (no cross-image leakage through halos / tile tails)."""
    m = seeded_model
    m.gemm_backend = 0
    m._handle = None
    m.cuda()
    x = recipes.model_input(9, seed=77).cuda()
    with torch.no_grad():
        y9 = m(x)
        y2 = m(x[3:5].contiguous())
    assert torch.equal(y9[3:5], y2)
    c = m.count_parameters()
    assert c["total"] == 96_371_197
    prof = m.profile_memory(batch_size=2)
    assert prof["batch_size"] == 2 and prof["peak_vram_gb"] > 0
    assert m.launches_per_forward() == 85       # the 1x1 head rides in the head conv, the blocks' LayerNorms in their GEMMs
    m.cpu()


def _oracle_on_gpu(sd, x, chunk=16, ln_eps_block=1e-5):
    """the fp32 oracle evaluated by torch on the device (TF32 off: true fp32 cuBLAS / cuDNN), in chunks"""
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        sd = {k: v.cuda() for k, v in sd.items()}
        return torch.cat([omodel.restoration_forward(sd, x[i:i + chunk], ln_eps_block=ln_eps_block) for i in range(0, x.shape[0], chunk)])
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("B,perturbed,layout", [(32, False, "fallback"), (32, True, "fallback"), (256, False, "fallback"), (32, True, "timm")])
def test_model_forward_at_benchmark_batch_vs_fp32_oracle(seeded_model, B, perturbed, layout):
    """The benchmark's own batch: >= 2,048 token rows put every encoder GEMM on the cta_group::2 pair kernels, stage 3 on the
    resident-weight pair conv and the ResidualBlock skips on E_CONV_SKIPADD -- all under one whole-model check against the fp32
    oracle evaluated by torch on the device (TF32 off).  Seeded reference weights: the BASELINE bounds (2e-2 max-abs, 1e-3 RMS).
    Perturbed BatchNorm statistics / decoder biases (folding counts, outputs ~10x larger): the same accuracy relative to the output.
    "timm": the timm key layout with LayerNorm eps 1e-6 inside the blocks (the folded-LayerNorm epilogues take eps from the model)."""
    sd = {k: v.clone() for k, v in seeded_model.state_dict().items()}
    if perturbed:
        recipes.perturb_decoder_stats(sd)
        g = torch.Generator().manual_seed(11)
        for k, v in sd.items():      # and non-trivial LayerNorm affine terms (identity at init would hide a folding mistake)
            if ".norm" in k and k.endswith("weight") and v.ndim == 1:
                v.copy_(1 + 0.2 * torch.randn(v.shape, generator=g))
            elif ".norm" in k and k.endswith("bias"):
                v.copy_(0.1 * torch.randn(v.shape, generator=g))
    m = SatMAERestoration(recipes.base_config(), block_layout=layout).eval()
    m.load_state_dict(sd)
    m.cuda()
    x = recipes.model_input(B, seed=500 + B).cuda()
    with torch.no_grad():
        y = m(x)
    ref = _oracle_on_gpu(m.cpu().state_dict() if layout == "timm" else sd, x, ln_eps_block=1e-6 if layout == "timm" else 1e-5)
    m.cuda()
    err = (y - ref).abs()
    rms, std = err.pow(2).mean().sqrt().item(), ref.std().item()
    if not perturbed:
        assert err.max().item() <= 2e-2 and rms <= 1e-3, f"max-abs {err.max().item():.3e}, rms {rms:.3e}"
    assert rms <= 0.02 * std, f"relative rms {rms / std:.3%}"
    # every image on its own, too (a tile-tail or halo bug would hit single images, not the batch RMS)
    per_img = err.flatten(1).pow(2).mean(1).sqrt()
    assert per_img.max().item() <= 0.03 * std, f"worst image rms {per_img.max().item():.3e} vs output std {std:.3e}"
    assert m.launches_per_forward() == 85


def test_model_forward_with_and_without_folded_layernorm(seeded_model, monkeypatch):
    """A/B: the same weights through the stand-alone LayerNorm kernels (MSD_NO_LN_FUSION) -- both must sit inside the bounds
    against the oracle and within bf16 noise of each other."""
    sd = {k: v.clone() for k, v in seeded_model.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    for k, v in sd.items():      # non-trivial LayerNorm affine terms (identity at init would hide a folding mistake)
        if ".norm" in k and k.endswith("weight") and v.ndim == 1:
            v.copy_(1 + 0.2 * torch.randn(v.shape, generator=g))
        elif ".norm" in k and k.endswith("bias"):
            v.copy_(0.1 * torch.randn(v.shape, generator=g))
    x = recipes.model_input(3, seed=91)
    ref = omodel.restoration_forward(sd, x)
    outs = []
    for no_fuse in (False, True):
        if no_fuse:
            monkeypatch.setenv("MSD_NO_LN_FUSION", "1")
        m = SatMAERestoration(recipes.base_config(), block_layout="fallback").eval()
        m.load_state_dict(sd)
        m.cuda()
        with torch.no_grad():
            outs.append(m(x.cuda()).cpu())
        assert m.launches_per_forward() == (109 if no_fuse else 85)
        _check_pixels(outs[-1], ref)
    assert (outs[0] - outs[1]).abs().max().item() <= 5e-3


def test_forward_is_unchanged_after_moving_the_model_off_and_back_onto_the_device():
    """base.yaml freezes blocks 0..5, whose packed weights are cached: after .cpu() / .cuda() the parameters sit at other
    addresses (often the ones a sibling block had) and the cached copies must not be matched by address."""
    torch.manual_seed(42)
    m = SatMAERestoration(recipes.base_config(), block_layout="fallback").eval()
    x = recipes.model_input(2).cuda()
    outs = []
    for backend in (0, 0, 1, 0):
        m.gemm_backend = backend
        m.cuda()
        with torch.no_grad():
            outs.append(m(x).cpu())
        m.cpu()
        junk = [torch.empty(n, device="cuda") for n in (768 * 3072, 768 * 768, 2304 * 768)]   # shuffle the allocator's free lists
        del junk
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[3])
    assert (outs[2] - outs[0]).abs().max().item() <= 2e-2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_in_the_same_process_gives_the_same_result():
    """One process driving two GPUs: kernel attributes (opt-in shared memory) are per device, tensor maps and workspaces
    per call -- the second device must produce exactly what the first one did."""
    from msd_b200.training import MetricsTracker
    outs = []
    x = recipes.model_input(2)
    for dev in ("cuda:0", "cuda:1"):
        torch.manual_seed(42)
        m = SatMAERestoration(recipes.base_config(), block_layout="fallback").eval().to(dev)
        with torch.no_grad():
            y = m(x.to(dev))
        t = MetricsTracker()
        t.update(y, x.to(dev))
        outs.append((y.cpu(), t.compute()))
    assert torch.equal(outs[0][0], outs[1][0])
    assert outs[0][1] == outs[1][1]
