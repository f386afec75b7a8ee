"""Host-side logic (runs on the CPU): parameter tree / state-dict compatibility with the reference, key-layout
adapter, and the weight packing + implicit-GEMM indexing the CUDA kernels rely on (emulated in fp32)."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import emulate
import recipes
from msd_b200.models import SatMAERestoration, convert_block_keys
from msd_b200.models.satmae_restoration import pack_conv3x3, pack_convt2x2


@pytest.fixture(scope="module")
def model():
    torch.manual_seed(42)
    return SatMAERestoration(recipes.base_config(), block_layout="fallback")


def test_state_dict_matches_reference_init(model, golden_model):
    """seed-42 construction reproduces the reference's 284 tensors (checksums recorded from the reference)."""
    sd = model.state_dict()
    assert list(sd.keys()) == list(golden_model["state/keys"])
    np.testing.assert_allclose([v.double().sum().item() for v in sd.values()], golden_model["state/sum"], rtol=0, atol=1e-9)
    np.testing.assert_allclose([v.double().abs().sum().item() for v in sd.values()], golden_model["state/abssum"], rtol=1e-12)


def test_parameter_counts(model):
    c = model.count_parameters()
    assert c["encoder_total"] == 87_724_800 and c["decoder_total"] == 8_646_397 and c["total"] == 96_371_197
    assert c["frozen"] == 42_527_232 and c["trainable"] == 53_843_965
    assert model.encoder.patch_embed.proj.weight.requires_grad and model.encoder.pos_embed.requires_grad
    assert not any(p.requires_grad for p in model.encoder.blocks[0].parameters())
    assert all(p.requires_grad for p in model.encoder.blocks[6].parameters())


def test_optimizer_groups(model):
    groups = model.get_optimizer_param_groups("1e-4", weight_decay=0.05)
    assert len(groups) == 2 and groups[0]["lr"] == pytest.approx(1e-5) and groups[1]["lr"] == pytest.approx(1e-4)
    assert sum(p.numel() for p in groups[0]["params"]) == 45_197_568
    assert sum(p.numel() for p in groups[1]["params"]) == 8_646_397


def test_timm_and_fallback_layouts_interchange(model):
    sd = model.state_dict()
    timm_sd = convert_block_keys(sd, "timm")
    assert "encoder.blocks.3.attn.qkv.weight" in timm_sd and "encoder.blocks.3.mlp.fc2.bias" in timm_sd
    assert not any("in_proj" in k for k in timm_sd)
    m2 = SatMAERestoration(recipes.base_config(), block_layout="timm")
    assert m2.encoder.ln_eps_block == 1e-6 and model.encoder.ln_eps_block == 1e-5
    m2.load_state_dict(sd, strict=True)          # fallback keys into a timm-layout module
    m3 = SatMAERestoration(recipes.base_config(), block_layout="fallback")
    m3.load_state_dict(m2.state_dict(), strict=True)   # and back
    for k, v in sd.items():
        assert torch.equal(m3.state_dict()[k], v), k


def test_forward_rejects_cpu_and_training(model):
    from msd_b200 import MsdError
    model.eval()
    with pytest.raises(MsdError):
        model(torch.zeros(1, 13, 192, 192))
    with pytest.raises(AssertionError):
        model(torch.zeros(1, 12, 192, 192))


@pytest.mark.parametrize("cin,cout,H,W", [(48, 48, 6, 5), (96, 32, 4, 7), (64, 16, 3, 3), (96, 96, 5, 4)])
def test_conv3x3_packing_and_tap_table(cin, cout, H, W):
    torch.manual_seed(0)
    conv = nn.Conv2d(cin, cout, 3, padding=1, bias=False)
    bn = nn.BatchNorm2d(cout).eval()
    bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 1.5); bn.weight.data.normal_(1, 0.1); bn.bias.data.normal_(0, 0.1)
    x = torch.randn(2, cin, H, W)
    skip = torch.randn(2, cout, H, W)
    with torch.no_grad():
        ref = F.gelu(skip + bn(conv(x)))
        w, b = pack_conv3x3(conv, bn, with_skip=(cin == cout))     # identity tenth tap when the shapes allow it
        got = emulate.conv3x3_rows(emulate.halo_nhwc(x), w.float(), b, H, W, skip=emulate.halo_nhwc(skip))
    assert w.shape == (cout, (10 if cin == cout else 9) * ((cin + 63) // 64 * 64)) and w.dtype == torch.bfloat16
    assert torch.all(got[:, 0] == 0) and torch.all(got[:, :, 0] == 0) and torch.all(got[:, -1] == 0) and torch.all(got[:, :, -1] == 0)
    # bf16 weights -> ~3 significant digits
    torch.testing.assert_close(emulate.unhalo_nchw(got), ref, rtol=0, atol=0.05)
    with torch.no_grad():
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        w32 = (conv.weight * scale.view(-1, 1, 1, 1)).permute(0, 2, 3, 1)
        kpad = (cin + 63) // 64 * 64
        w_exact = torch.zeros(cout, 3, 3, kpad); w_exact[..., :cin] = w32
        got32 = emulate.conv3x3_rows(emulate.halo_nhwc(x), w_exact.reshape(cout, -1), b, H, W, skip=emulate.halo_nhwc(skip))
    torch.testing.assert_close(emulate.unhalo_nchw(got32), ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("cin,cout,Hin,Win,padded", [(96, 48, 3, 4, True), (64, 32, 2, 2, False), (128, 16, 3, 3, True)])
def test_convt_packing_and_pixel_shuffle(cin, cout, Hin, Win, padded):
    torch.manual_seed(1)
    ct = nn.ConvTranspose2d(cin, cout, 2, stride=2)
    bn = nn.BatchNorm2d(cout).eval()
    bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 1.5); bn.weight.data.normal_(1, 0.1); bn.bias.data.normal_(0, 0.1)
    x = torch.randn(2, cin, Hin, Win)
    with torch.no_grad():
        ref = F.gelu(bn(ct(x)))
        w, b = pack_convt2x2(ct, bn)
        xin = emulate.halo_nhwc(x) if padded else x.permute(0, 2, 3, 1).contiguous()
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        w32 = (ct.weight * scale.view(1, -1, 1, 1)).permute(2, 3, 1, 0).reshape(4 * cout, cin)
        got = emulate.convt2x2_rows(xin, w32, b, Hin, Win, padded)
    assert w.shape == (4 * cout, (cin + 63) // 64 * 64)
    torch.testing.assert_close(w.float()[:, :cin], w32, rtol=0.01, atol=1e-3)
    torch.testing.assert_close(emulate.unhalo_nchw(got), ref, rtol=1e-4, atol=1e-4)


def test_patch_embed_weight_is_im2col_order(model):
    """conv k16 s16 == (c,ky,kx)-ordered patch rows x proj.weight.reshape(768,-1)^T (encoder.py:48-50)."""
    x = recipes.model_input(1)
    pe = model.encoder.patch_embed.proj
    with torch.no_grad():
        ref = pe(x).flatten(2).transpose(1, 2)[0]
        patches = x[0].reshape(13, 12, 16, 12, 16).permute(1, 3, 0, 2, 4).reshape(144, 13 * 256)   # n = 12*py+px
        got = patches @ pe.weight.reshape(768, -1).t() + pe.bias
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)


def test_pretrained_checkpoint_surgery_matches_reference(tmp_path):
    """SatMAE checkpoint adaptation: 3 -> 13 channel patch embedding, CLS-token drop / bicubic pos-embed
    interpolation, timm -> fallback key renaming.  Compared tensor-for-tensor with the reference's own loader when the
    reference tree is present (this container); on the GPU box the structural checks alone run."""
    import sys
    from pathlib import Path
    from msd_b200.models.encoder import SatMAEEncoder
    g = torch.Generator().manual_seed(0)
    D, depth = 64, 2
    ck = {"patch_embed.proj.weight": torch.randn(D, 3, 16, 16, generator=g), "patch_embed.proj.bias": torch.randn(D, generator=g),
          "pos_embed": torch.randn(1, 1 + 14 * 14, D, generator=g), "cls_token": torch.randn(1, 1, D, generator=g),
          "norm.weight": torch.randn(D, generator=g), "norm.bias": torch.randn(D, generator=g)}
    for i in range(depth):
        ck.update({f"blocks.{i}.attn.qkv.weight": torch.randn(3 * D, D, generator=g), f"blocks.{i}.attn.qkv.bias": torch.randn(3 * D, generator=g),
                   f"blocks.{i}.attn.proj.weight": torch.randn(D, D, generator=g), f"blocks.{i}.mlp.fc1.weight": torch.randn(4 * D, D, generator=g),
                   f"blocks.{i}.norm1.weight": torch.randn(D, generator=g)})
    path = tmp_path / "satmae.pth"
    torch.save({"model": ck}, path)
    enc = SatMAEEncoder(in_channels=13, image_size=192, patch_size=16, embed_dim=D, depth=depth, num_heads=1,
                        pretrained_path=str(path), block_layout="fallback")
    sd = enc.state_dict()
    w = sd["patch_embed.proj.weight"]
    assert torch.equal(w[:, :3], ck["patch_embed.proj.weight"]) and all(torch.equal(w[:, c], ck["patch_embed.proj.weight"][:, 0]) for c in range(3, 13))
    assert sd["pos_embed"].shape == (1, 144, D)                                   # 14x14 (+CLS) -> 12x12, bicubic
    assert torch.equal(sd["blocks.1.attn.in_proj_weight"], ck["blocks.1.attn.qkv.weight"])      # timm keys land in the fallback layout
    assert torch.equal(sd["blocks.0.mlp.0.weight"], ck["blocks.0.mlp.fc1.weight"]) and torch.equal(sd["norm.bias"], ck["norm.bias"])
    ref_root = Path("/root/reference")
    if not (ref_root / "src").exists():
        return
    sys.path.insert(0, str(ref_root))
    try:
        from src.models.encoder import SatMAEEncoder as RefEncoder
        ref = RefEncoder(in_channels=13, image_size=192, patch_size=16, embed_dim=D, depth=depth, num_heads=1, pretrained_path=str(path))
    finally:
        sys.path.remove(str(ref_root))
    rsd = ref.state_dict()
    for k in ("patch_embed.proj.weight", "patch_embed.proj.bias", "pos_embed", "norm.weight", "norm.bias"):
        assert torch.equal(sd[k], rsd[k]), k


def test_frozen_parameter_cache_follows_the_parameter_not_the_address():
    """The packed copies of frozen parameters are cached; twelve blocks share every shape, so a cache keyed by address
    alone hands block i the copy of block j once `.cpu()` / `.cuda()` has re-homed the parameters."""
    m = SatMAERestoration(recipes.base_config(), block_layout="fallback")
    a = m.encoder.blocks[0].mlp[0].weight
    b = m.encoder.blocks[1].mlp[0].weight
    assert not a.requires_grad and not b.requires_grad and a.shape == b.shape      # base.yaml freezes blocks 0..5
    calls = []
    cast = lambda t: (calls.append(1), t.detach().clone())[1]
    ca = m._cached_cast(a, "b16", cast)
    assert m._cached_cast(a, "b16", cast) is ca and len(calls) == 1                # same parameter, same version: reused
    with torch.no_grad():
        b.data = a.data                                                           # b now lives at a's address
    cb = m._cached_cast(b, "b16", cast)
    assert cb is not ca and len(calls) == 2                                        # another parameter object: not a's copy
    with torch.no_grad():
        a.add_(1.0)                                                                # in-place update bumps the version
    assert m._cached_cast(a, "b16", cast) is not ca
    v = a.reshape(-1)                                                              # views are cached under their base parameter
    cv = m._cached_cast(v, "flat", cast)
    assert m._cached_cast(a.reshape(-1), "flat", cast) is cv
    m.double()                                                                     # any _apply drops the packed copies
    assert "_frozen_cache" not in m.__dict__ and m._packed_key is None


def test_layernorm_fold_algebra():
    """pack_ln_folded_linear: Linear(LayerNorm(x)) == rstd * (x . (gamma (.) W)^T - mean * s) + b' -- the identity the GEMM epilogue
    evaluates (csrc/gemm.cuh, Problem; reference encoder.py:154-176).  fp64 so that only the algebra is tested; the packed weight is
    bf16-rounded, so the comparison uses the rounded weight on both sides."""
    import torch.nn.functional as F
    from msd_b200.models.satmae_restoration import pack_ln_folded_linear
    g = torch.Generator().manual_seed(0)
    D, N, M = 768, 96, 17
    x = torch.randn(M, D, generator=g, dtype=torch.float64) * 2 + 3
    w, b = torch.randn(N, D, generator=g) * D ** -0.5, torch.randn(N, generator=g)
    gam, bet = 1 + 0.3 * torch.randn(D, generator=g), 0.2 * torch.randn(D, generator=g)
    wg, s, bl = pack_ln_folded_linear(w, b, gam, bet)
    assert wg.dtype == torch.bfloat16 and s.dtype == torch.float32 and bl.dtype == torch.float32
    for eps in (1e-5, 1e-6):
        mean, var = x.mean(1, keepdim=True), x.var(1, unbiased=False, keepdim=True)
        rstd = (var + eps).rsqrt()
        fused = rstd * (x @ wg.double().t() - mean * s.double()) + bl.double()
        w_eff = wg.double() / gam.double()             # the weight the fold actually represents after bf16 rounding
        ref = F.linear(F.layer_norm(x, (D,), gam.double(), None, eps), w_eff) + (b.double() + w.double() @ bet.double())
        assert (fused - ref).abs().max().item() < 1e-5     # fp32 column sums are the only non-fp64 term


def test_export_graph_is_the_reference_forward_and_only_runs_for_tracers(golden_model):
    """scripts/export.py traces model.forward on the CPU.  Under a tracer (or export_mode()) the drop-in
    evaluates its torch-op graph definition -- equal to the oracle / the reference's golden output to fp32 rounding, for both
    block layouts -- and everywhere else a CPU tensor still raises (no CPU execution path)."""
    import pytest
    import recipes
    from msd_b200 import MsdError
    from msd_b200.models import SatMAERestoration, export_mode
    from oracle import model as omodel
    torch.manual_seed(42)
    m = SatMAERestoration(recipes.base_config(), block_layout="fallback").eval()
    x = recipes.model_input(2)[:1]
    with pytest.raises(MsdError):
        m(x)
    with torch.no_grad(), export_mode():
        y = m(x)
    with pytest.raises(MsdError):                       # the switch does not leak out of the block
        m(x)
    ref = omodel.restoration_forward({k: v.clone() for k, v in m.state_dict().items()}, x)
    assert (y - ref).abs().max().item() <= 2e-6
    assert (y[:, :, ::4, ::4] - torch.from_numpy(golden_model["y_sub"])[:1]).abs().max().item() <= 2e-5      # the reference itself
    with torch.no_grad():
        traced = torch.jit.trace(m, x, check_trace=False)           # what torch.onnx.export does first (export_onnx.py:40-69)
        assert torch.equal(traced(x), y)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    recipes.perturb_decoder_stats(sd)
    t = SatMAERestoration(recipes.base_config(), block_layout="timm").eval()
    t.load_state_dict(sd)
    with torch.no_grad(), export_mode():
        yt = t(x)
    reft = omodel.restoration_forward(t.state_dict(), x, ln_eps_block=1e-6)
    assert (yt - reft).abs().max().item() <= 2e-5 * max(1.0, reft.abs().max().item())
