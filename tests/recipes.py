"""Seeded input recipes shared by oracle/make_golden.py (which runs the reference on them) and the tests (which run
the oracle and the CUDA path on the same tensors).  Everything is generated on the CPU with explicit generators so
that the GPU box reproduces the authoring container's tensors bit for bit (checked through stored checksums)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

# (gaussian_sigma, speckle_sigma, dead_band_prob, thermal_scale)
NOISE_CONFIGS = {
    "base": (0.015, 0.008, 0.08, 0.005),      # reference configs/base.yaml:20-24
    "high": (0.035, 0.015, 0.15, 0.01),       # configs/experiments/high_noise.yaml:7-11
    "default": (0.05, 0.03, 0.25, 0.01),      # AddSensorNoise constructor defaults (transforms.py:40-44)
    "nodead": (0.02, 0.0, 0.0, 0.01),
}


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def iid_tiles(B: int, seed: int = 0, C: int = 13, H: int = 192, W: int = 192):
    """clean ~ U[-2,2] (dataset range); pred = clamp(clean + 0.05 N(0,1), -3, 3)."""
    g = _gen(seed)
    a = torch.rand(B, C, H, W, generator=g) * 4 - 2
    b = (a + 0.05 * torch.randn(B, C, H, W, generator=g)).clamp(-3, 3)
    return a, b


def eurosat_tiles(B: int, seed: int = 1, noise: float = 0.02):
    """EuroSAT-shaped smooth fields: 64x64 reflectance -> bilinear 192 -> clamp -> (x-0.5)/0.25."""
    g = _gen(seed)
    raw = torch.rand(B, 13, 64, 64, generator=g) * 0.52
    a = F.interpolate(raw, size=192, mode="bilinear", align_corners=False).clamp(0, 1)
    a = (a - 0.5) / 0.25
    b = (a + noise * torch.randn(B, 13, 192, 192, generator=g)).clamp(-3, 3)
    return a.contiguous(), b.contiguous()


def metric_cases():
    cases = {}
    cases["iid_full"] = iid_tiles(2, seed=0)
    cases["eurosat_full"] = eurosat_tiles(2, seed=1)
    cases["small_ragged"] = iid_tiles(2, seed=2, C=13, H=48, W=40)     # not a multiple of the 32x64 tile
    cases["small_c4"] = iid_tiles(3, seed=3, C=4, H=37, W=70)          # non-13-band, ragged both ways
    return cases


def small_clean():
    return torch.rand(2, 13, 8, 8, generator=_gen(5)) * 4 - 2


def model_input(B: int, seed: int = 1234):
    return torch.randn(B, 13, 192, 192, generator=_gen(seed))


def base_config():
    """The model / data part of the reference's configs/base.yaml:9-45 (pretrained_path disabled: no weights offline)."""
    return {
        "data": {"num_bands": 13, "image_size": 192, "patch_size": 16},
        "model": {
            "encoder": {"name": "satmae_vit_base", "pretrained_path": None, "freeze_layers": [0, 1, 2, 3, 4, 5],
                        "gradient_checkpointing": True, "input_channels": 13, "patch_size": 16, "embed_dim": 768,
                        "depth": 12, "num_heads": 12},
            "decoder": {"architecture": "unet_light", "channels": [384, 192, 96, 48], "use_attention": False,
                        "output_channels": 13},
        },
    }


@torch.no_grad()
def perturb_decoder_stats(sd: dict, seed: int = 7):
    """Give BatchNorm non-trivial running statistics / affine terms and the decoder biases non-zero values (at init
    BN is the identity, which would hide folding mistakes).  In place; same draws in make_golden and the tests."""
    g = _gen(seed)
    for k, v in sd.items():
        if k.endswith("running_mean"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g) * 0.5 + 0.75)
        elif "decoder" in k and k.endswith(".bias"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.05)
        elif "decoder" in k and k.endswith(".weight") and v.ndim == 1:
            v.copy_(1.0 + torch.randn(v.shape, generator=g) * 0.1)
    return sd
