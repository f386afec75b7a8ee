"""GPU parity: Philox stream, corruption simulator, fused metrics, loss fwd/bwd -- CUDA path (through the C ABI) vs the
oracle and the reference-generated golden vectors.  Integer / mask / RNG results are compared bit for bit."""
import ctypes

import numpy as np
import pytest
import torch

import gpu_helpers as gh
import recipes
from msd_b200 import _lib
from msd_b200.data import AddSensorNoise
from msd_b200.training import (CombinedLoss, MetricsTracker, SSIMLoss, calculate_mae, calculate_psnr, calculate_rmse,
                               calculate_sam, calculate_ssim)
from msd_b200.training.metrics import fused_metrics
from oracle import corrupt as ocorrupt
from oracle import loss as oloss
from oracle import metrics as ometrics
from oracle import philox

pytestmark = pytest.mark.gpu


# ----------------------------------------------------------------------------------------------- philox
@pytest.mark.parametrize("numel", [1, 13, 255, 4099, 13 * 192 * 192, 1184 * 256 * 4 + 17, 3 * 13 * 192 * 192])
def test_philox_stream_is_torchs(lib, numel):
    torch.cuda.init()
    gen = torch.cuda.default_generators[0]
    for seed, skip in ((123, 0), (2 ** 40 + 7, 3)):
        torch.manual_seed(seed)
        for _ in range(skip):
            torch.rand(5, device="cuda")
        off = gen.get_offset()
        expect_n = torch.randn(numel, device="cuda")
        assert gen.get_offset() - off == lib.msd_philox_offset_increment(numel, gh.torch_grid(numel))
        off_u = gen.get_offset()
        expect_u = torch.rand(numel, device="cuda")
        out = torch.empty(numel, device="cuda")
        _lib.check(lib.msd_philox_normal_f32(_lib.ptr(out), numel, seed, off, gh.torch_grid(numel), _lib.stream_ptr()))
        assert torch.equal(out, expect_n)
        _lib.check(lib.msd_philox_uniform_f32(_lib.ptr(out), numel, seed, off_u, gh.torch_grid(numel), _lib.stream_ptr()))
        assert torch.equal(out, expect_u)
        if numel <= 5000:   # and the numpy restatement of the same stream (uniforms are exactly reproducible on the CPU)
            assert np.array_equal(philox.torch_cuda_rand(seed, off_u, numel, gh.torch_grid(numel)), expect_u.cpu().numpy())


# ----------------------------------------------------------------------------------------------- corruption
@pytest.mark.parametrize("name", list(recipes.NOISE_CONFIGS))
@pytest.mark.parametrize("shape", [(2, 13, 8, 8), (3, 13, 192, 192), (13, 64, 48), (5, 7, 33, 21)])
def test_corruption_bit_exact_vs_reference_semantics_on_cuda(name, shape):
    prm = recipes.NOISE_CONFIGS[name]
    g = torch.Generator().manual_seed(11)
    x = (torch.rand(*shape, generator=g) * 4 - 2).cuda()
    gen = torch.cuda.default_generators[0]
    torch.manual_seed(42)
    torch.rand(3, device="cuda")                 # start from a non-zero offset
    ref, ref_mask = ocorrupt.add_sensor_noise_torch(x, *prm, return_mask=True)
    ref_off = gen.get_offset()
    torch.manual_seed(42)
    torch.rand(3, device="cuda")
    mod = AddSensorNoise(*prm)
    out = mod(x)
    assert gen.get_offset() == ref_off           # generator left where the reference leaves it
    assert out.shape == x.shape and out.dtype == x.dtype and out.device == x.device
    if prm[2] > 0:
        assert torch.equal(mod.last_dead_mask, ref_mask)     # dead-band masks: bit exact
        # and equal to the CPU restatement of the Philox stream
        C = x.shape[-3]
        n_full = x.numel()
        off = 4 + sum(philox.offset_increment(n_full, gh.torch_grid(n_full)) for s in prm[:2] if s > 0)
        B = x.shape[0] if x.ndim == 4 else 1
        assert np.array_equal(philox.dead_band_masks(42, off, B, C, prm[2]), ref_mask.cpu().numpy())
    assert torch.equal(out, ref), f"max diff {(out - ref).abs().max().item()}"
    assert out.min() >= -3 and out.max() <= 3


def test_corruption_striping_and_patch_operand():
    x = (torch.rand(4, 13, 192, 192, generator=torch.Generator().manual_seed(3)) * 4 - 2).cuda()
    prm = recipes.NOISE_CONFIGS["base"]
    torch.manual_seed(7)
    ref = ocorrupt.add_sensor_noise_torch(x, *prm, enable_striping=True, stripe_prob=1.0)
    torch.manual_seed(7)
    out = AddSensorNoise(*prm, enable_striping=True, stripe_prob=1.0)(x)
    assert torch.equal(out, ref)
    # fused patch-embed operand == patchify(noisy) in bf16, (c,ky,kx) order, token n = 12*py+px
    torch.manual_seed(7)
    patches = torch.empty(4 * 144, 13 * 256, dtype=torch.bfloat16, device="cuda")
    out2 = AddSensorNoise(*prm)(x, patch_out=patches)
    expect = out2.reshape(4, 13, 12, 16, 12, 16).permute(0, 2, 4, 1, 3, 5).reshape(4 * 144, 13 * 256).to(torch.bfloat16)
    assert torch.equal(patches, expect)
    lib = _lib.load()
    p2 = torch.empty_like(patches)
    _lib.check(lib.msd_patchify_bf16(_lib.ptr(out2), _lib.ptr(p2), 4, 13, 192, 16, _lib.stream_ptr()))
    assert torch.equal(p2, expect)


def test_corruption_full_batch_statistics():
    """B=256 (config C2 size): size-independent properties -- mask density, per-band thermal std on dead bands, clamp."""
    prm = recipes.NOISE_CONFIGS["base"]
    x = torch.zeros(256, 13, 192, 192, device="cuda")
    torch.manual_seed(42)
    mod = AddSensorNoise(*prm)
    y = mod(x)
    m = mod.last_dead_mask
    assert 0.05 < m.float().mean().item() < 0.11
    std = y.flatten(2).std(dim=2)                       # [256,13]
    w = torch.linspace(1, 2, 13, device="cuda")
    dead_std = (std / (0.005 * w))[m]
    assert (dead_std - 1).abs().max() < 0.03            # thermal only on dead bands (D5)
    live_std = std[~m]
    assert (live_std > 0.0149).all() and (live_std < 0.0185).all()
    # a CPU sample (the reference's call site inside Dataset.__getitem__, dataset.py:188-190) is staged through the device: same
    # kernel, same stream of random numbers, result returned on the CPU -- not a CPU implementation
    xs = torch.rand(13, 24, 40) * 4 - 2
    torch.manual_seed(7)
    on_dev = mod(xs.cuda())
    torch.manual_seed(7)
    via_host = mod(xs)
    assert via_host.device.type == "cpu" and torch.equal(via_host, on_dev.cpu())


# ----------------------------------------------------------------------------------------------- metrics
@pytest.mark.parametrize("name", ["iid_full", "eurosat_full", "small_ragged", "small_c4"])
def test_metrics_vs_oracle_and_reference(golden_metrics, name):
    a, b = recipes.metric_cases()[name]
    g = golden_metrics
    per, sums, raw = fused_metrics(b.cuda(), a.cuda())
    per = per.cpu().numpy()
    ref = g[f"{name}/per_image"]
    assert np.abs(per[:, 0] - ref[:, 0]).max() < 0.05      # PSNR dB   (BASELINE bounds)
    assert np.abs(per[:, 1] - ref[:, 1]).max() < 1e-4      # SSIM
    assert np.abs(per[:, 2] - ref[:, 2]).max() < 0.01      # SAM degrees
    np.testing.assert_allclose(per[:, 3], ref[:, 3], rtol=1e-5)
    # much tighter against the numpy oracle on the same tensors
    orc = np.array([[ometrics.calculate_psnr(b[i:i + 1], a[i:i + 1]), ometrics.calculate_ssim(b[i:i + 1], a[i:i + 1]),
                     ometrics.calculate_sam(b[i:i + 1], a[i:i + 1]), ometrics.calculate_rmse(b[i:i + 1], a[i:i + 1])]
                    for i in range(a.shape[0])])
    np.testing.assert_allclose(per, orc, rtol=2e-5, atol=2e-5)
    t = MetricsTracker()
    t.update(b.cuda(), a.cuda())
    c = t.compute()
    np.testing.assert_allclose([c["psnr"], c["ssim"], c["sam"], c["rmse"]], g[f"{name}/tracker"], rtol=2e-5, atol=2e-5)
    assert t.count == a.shape[0] and "PSNR" in repr(t)
    bg = g[f"{name}/batch"]
    assert calculate_psnr(b.cuda(), a.cuda()) == pytest.approx(bg[0], abs=1e-3)
    assert calculate_ssim(b.cuda(), a.cuda()) == pytest.approx(bg[1], abs=2e-5)
    assert calculate_sam(b.cuda(), a.cuda()) == pytest.approx(bg[2], abs=1e-3)
    assert calculate_rmse(b.cuda(), a.cuda()) == pytest.approx(bg[3], rel=1e-5)
    assert calculate_mae(b.cuda(), a.cuda()) == pytest.approx(bg[4], rel=1e-5)
    assert calculate_ssim(b[0].cuda(), a[0].cuda()) == pytest.approx(ref[0, 1], abs=2e-5)   # 3-D input


def test_metrics_edge_cases_and_tracker_semantics():
    z, o = torch.zeros(1, 13, 16, 16, device="cuda"), torch.ones(1, 13, 16, 16, device="cuda")
    assert calculate_sam(z, z) == pytest.approx(90.0, abs=1e-4)
    assert calculate_psnr(o, o) == float("inf")
    assert calculate_ssim(o, o) == pytest.approx(1.0, abs=1e-6)
    t = MetricsTracker()
    assert t.compute() == {"psnr": 0.0, "ssim": 0.0, "sam": 0.0, "rmse": 0.0}
    a, b = recipes.iid_tiles(4, seed=9, H=64, W=64)
    t.update(b[:2].cuda(), a[:2].cuda())
    t.update(b[2:].cuda(), a[2:].cuda())
    one = MetricsTracker()
    one.update(b.cuda(), a.cuda())
    for k in ("psnr", "ssim", "sam", "rmse"):
        assert t.compute()[k] == pytest.approx(one.compute()[k], rel=1e-12)
    t.update(o, o)                                   # a perfect image makes the mean PSNR inf, as in the reference
    assert t.compute()["psnr"] == float("inf") and t.count == 5
    t.reset()
    assert t.count == 0


def test_metrics_full_size_linearity():
    """config-C3 shaped batch (64 tiles): tracker sums are additive over shards (what the NCCL allreduce relies on)."""
    a, b = recipes.eurosat_tiles(64, seed=21)
    a, b = a.cuda(), b.cuda()
    _, s_all, _ = fused_metrics(b, a)
    parts = [fused_metrics(b[i:i + 16], a[i:i + 16])[1] for i in range(0, 64, 16)]
    np.testing.assert_allclose(sum(parts).cpu().numpy(), s_all.cpu().numpy(), rtol=1e-12)
    assert s_all[4].item() == 64


# ----------------------------------------------------------------------------------------------- loss
@pytest.mark.parametrize("name", ["iid_full", "eurosat_full", "small_ragged"])
def test_loss_forward_backward(golden_metrics, name):
    a, b = recipes.metric_cases()[name]
    g = golden_metrics
    p = b.cuda().requires_grad_(True)
    crit = CombinedLoss(1.0, 0.1)
    loss = crit(p, a.cuda())
    assert loss.ndim == 0 and loss.dtype == torch.float32
    (loss * 3.0).backward()
    ref = g[f"{name}/loss"]                                  # total, 1-ssim, mse   (reference, CPU fp32)
    assert loss.item() == pytest.approx(ref[0], rel=2e-5)
    assert crit.last_terms[1].item() == pytest.approx(ref[2], rel=2e-5)
    assert SSIMLoss()(b.cuda(), a.cuda()).item() == pytest.approx(ref[1], rel=2e-4, abs=2e-7)
    grad = p.grad / 3.0
    total, mse, sl, ograd = oloss.combined_loss_and_grad(b, a, dtype=torch.float64)    # fp64 oracle gradient
    scale = ograd.abs().max().item()
    assert (grad.cpu().double() - ograd).abs().max().item() < 2e-4 * scale
    gsum, gabs = g[f"{name}/grad_checksum"]                   # the signed sum cancels: bound it by the abs-sum
    assert grad.double().abs().sum().item() == pytest.approx(gabs, rel=2e-4)
    assert abs(grad.double().sum().item() - gsum) < 1e-6 * gabs
    # forward-only path
    with torch.no_grad():
        assert crit(b.cuda(), a.cuda()).item() == pytest.approx(ref[0], rel=2e-5)


def test_loss_weights_and_data_range():
    a, b = recipes.iid_tiles(2, seed=4, H=40, W=72)
    for wm, wsm, dr in ((0.0, 1.0, 6.0), (2.0, 0.0, 6.0), (0.5, 0.25, 2.0)):
        p = b.cuda().requires_grad_(True)
        loss = CombinedLoss(wm, wsm, data_range=dr)(p, a.cuda())
        loss.backward()
        total, _, _, ograd = oloss.combined_loss_and_grad(b, a, wm, wsm, dr, dtype=torch.float64)
        assert loss.item() == pytest.approx(total, rel=1e-5)
        assert (p.grad.cpu().double() - ograd).abs().max().item() < 2e-4 * ograd.abs().max().item() + 1e-12


def test_empty_batches_behave_like_the_reference():
    """B = 0: AddSensorNoise returns an empty tensor, MetricsTracker.update adds nothing, the model returns [0,13,S,S]."""
    import recipes
    from msd_b200.data import AddSensorNoise
    from msd_b200.models import SatMAERestoration
    from msd_b200.training import MetricsTracker
    x = torch.zeros(0, 13, 192, 192, device="cuda")
    assert AddSensorNoise()(x).shape == (0, 13, 192, 192)
    t = MetricsTracker()
    t.update(x, x)
    assert t.compute() == {"psnr": 0.0, "ssim": 0.0, "sam": 0.0, "rmse": 0.0}
    torch.manual_seed(0)
    m = SatMAERestoration(recipes.base_config(), block_layout="fallback").eval().cuda()
    with torch.no_grad():
        assert m(x).shape == (0, 13, 192, 192)


def test_tracker_prefetch_reads_the_same_sums_without_draining_the_stream():
    """prefetch() copies the running sums behind the kernels of the batch it follows; compute() then returns exactly what a
    blocking read would have, also when more work (another tracker's batch) was queued after the prefetch, and an update()
    after a prefetch invalidates the pending copy."""
    from msd_b200.training import MetricsTracker
    a, b = recipes.metric_cases()["small_c4"]
    a, b = a.cuda(), b.cuda()
    t1, t2 = MetricsTracker(), MetricsTracker()
    t1.update(b, a); t1.prefetch()
    t2.update(b * 0.5, a)                     # queued after t1's prefetch
    ref = MetricsTracker(); ref.update(b, a)
    assert t1.compute() == ref.compute()
    t1.update(b * 0.5, a)                     # pending copy is stale now
    ref.update(b * 0.5, a)
    assert t1.compute() == ref.compute() and t1.count == 2 * a.shape[0]
