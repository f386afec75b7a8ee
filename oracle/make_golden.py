"""Generate tests/golden/*.npz by RUNNING THE REFERENCE (imported from /root/reference, CPU fp32, torch eager).

    python oracle/make_golden.py            # needs /root/reference; the GPU box never runs this

The reference ships no tests or golden vectors, so these files are the pin: the oracle is checked
against them by the CPU suite, and the CUDA path is checked against the oracle (and against the full-size entries
here) on the GPU box.  Inputs that are too large to commit are regenerated from the recorded seed by
``tests/recipes.py``; an fp64 checksum of each regenerated input is stored so a drifted RNG is detected.
"""
from __future__ import annotations

import logging
import sys
from pathlib import Path

import numpy as np
import torch
import yaml

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
OUT = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import recipes  # noqa: E402  (tests/recipes.py: seeded input recipes shared with the tests)


def ref_imports():
    logging.disable(logging.WARNING)
    sys.path.insert(0, str(REF))
    from src.data.transforms import AddSensorNoise
    from src.models import SatMAERestoration
    from src.training.losses import CombinedLoss, SSIMLoss
    from src.training import metrics as M
    return AddSensorNoise, SatMAERestoration, CombinedLoss, SSIMLoss, M


def main():
    torch.set_num_threads(8)
    OUT.mkdir(parents=True, exist_ok=True)
    AddSensorNoise, SatMAERestoration, CombinedLoss, SSIMLoss, M = ref_imports()

    # ------------------------------------------------------------------ metrics + loss
    g = {}
    for name, (a, b) in recipes.metric_cases().items():
        g[f"{name}/checksum"] = np.array([a.double().sum().item(), b.double().sum().item()])
        per = []
        for i in range(a.shape[0]):
            per.append([M.calculate_psnr(b[i:i + 1], a[i:i + 1], 6.0), M.calculate_ssim(b[i:i + 1], a[i:i + 1], 6.0),
                        M.calculate_sam(b[i:i + 1], a[i:i + 1]), M.calculate_rmse(b[i:i + 1], a[i:i + 1])])
        g[f"{name}/per_image"] = np.array(per, dtype=np.float64)
        g[f"{name}/batch"] = np.array([M.calculate_psnr(b, a), M.calculate_ssim(b, a), M.calculate_sam(b, a),
                                       M.calculate_rmse(b, a), M.calculate_mae(b, a)], dtype=np.float64)
        tr = M.MetricsTracker()
        tr.update(b, a)
        c = tr.compute()
        g[f"{name}/tracker"] = np.array([c["psnr"], c["ssim"], c["sam"], c["rmse"]], dtype=np.float64)
        if a.shape[1] == 13:
            p = b.clone().requires_grad_(True)
            loss = CombinedLoss(1.0, 0.1)(p, a)
            loss.backward()
            g[f"{name}/loss"] = np.array([loss.item(), SSIMLoss()(b, a).item(), torch.nn.functional.mse_loss(b, a).item()])
            grad = p.grad.detach()
            g[f"{name}/grad_checksum"] = np.array([grad.double().sum().item(), grad.double().abs().sum().item()])
            if a.numel() <= 200_000:
                g[f"{name}/grad"] = grad.numpy()
            else:
                g[f"{name}/grad_sub"] = grad[:, :, ::8, ::8].contiguous().numpy()
        if a.numel() <= 200_000:
            g[f"{name}/a"], g[f"{name}/b"] = a.numpy(), b.numpy()
    z = torch.zeros(1, 13, 16, 16)
    o = torch.ones(1, 13, 16, 16)
    g["edge/sam_zeros"] = np.array([M.calculate_sam(z, z)])
    g["edge/psnr_same"] = np.array([M.calculate_psnr(o, o)])
    g["edge/ssim_ones"] = np.array([M.calculate_ssim(o, o)])
    np.savez_compressed(OUT / "metrics_loss.npz", **g)
    print("metrics_loss.npz:", {k: v.shape for k, v in g.items() if "checksum" not in k and v.size < 8})

    # ------------------------------------------------------------------ corruption (CPU generator: logic only)
    g = {}
    torch.manual_seed(42)
    zc = AddSensorNoise(0.0, 0.0, 0.08, 0.0)(torch.ones(8, 13, 4, 4))
    g["mask_logic/dead"] = (zc[:, :, 0, 0] == 0).numpy()
    for name, prm in recipes.NOISE_CONFIGS.items():
        x = recipes.small_clean()
        torch.manual_seed(42)
        y = AddSensorNoise(*prm)(x.clone())
        # replay the same CPU draws in the reference's order to expose them to the arithmetic oracle
        torch.manual_seed(42)
        n1 = torch.randn_like(x) if prm[0] > 0 else torch.zeros_like(x)
        n2 = torch.randn_like(x) if prm[1] > 0 else torch.zeros_like(x)
        masks = torch.stack([torch.rand(13) < prm[2] for _ in range(x.shape[0])]) if prm[2] > 0 else torch.zeros(x.shape[0], 13, dtype=torch.bool)
        n3 = torch.randn_like(x) if prm[3] > 0 else torch.zeros_like(x)
        g[f"{name}/x"], g[f"{name}/y"] = x.numpy(), y.numpy()
        g[f"{name}/n1"], g[f"{name}/n2"], g[f"{name}/n3"] = n1.numpy(), n2.numpy(), n3.numpy()
        g[f"{name}/mask"] = masks.numpy()
        g[f"{name}/w"] = torch.linspace(1.0, 2.0, 13).numpy()
    # D5: dead bands carry thermal noise only
    torch.manual_seed(42)
    y0 = AddSensorNoise(*recipes.NOISE_CONFIGS["base"])(torch.zeros(8, 13, 192, 192))
    g["d5/band_std_sample0"] = y0[0].flatten(1).std(dim=1).numpy()
    np.savez_compressed(OUT / "corrupt.npz", **g)
    print("corrupt.npz: dead rows", [np.nonzero(r)[0].tolist() for r in g["mask_logic/dead"]])

    # ------------------------------------------------------------------ model
    cfg = yaml.safe_load(open(REF / "configs" / "base.yaml"))
    cfg["model"]["encoder"]["pretrained_path"] = None
    torch.manual_seed(42)
    ref = SatMAERestoration(cfg).eval()
    sd = ref.state_dict()
    g = {}
    g["state/keys"] = np.array(list(sd.keys()))
    g["state/sum"] = np.array([v.double().sum().item() for v in sd.values()])
    g["state/abssum"] = np.array([v.double().abs().sum().item() for v in sd.values()])
    x = recipes.model_input(2)
    g["x_checksum"] = np.array([x.double().sum().item(), x.double().abs().sum().item()])
    with torch.no_grad():
        tok = ref.encoder(x)
        y = ref(x)
    g["tokens_sub"] = tok[:, ::4, ::8].contiguous().numpy()
    g["tokens_stats"] = np.array([tok.mean().item(), tok.std().item(), tok.abs().max().item()])
    g["y_sub"] = y[:, :, ::4, ::4].contiguous().numpy()
    g["y_stats"] = np.array([y.mean().item(), y.std().item(), y.abs().max().item()])
    g["y_band_mean"] = y.mean(dim=(0, 2, 3)).numpy()
    # second weight set with non-trivial BatchNorm statistics and biases (init BN is identity: mean 0, var 1)
    recipes.perturb_decoder_stats(sd)
    with torch.no_grad():
        y2 = ref(x)
    g["y2_sub"] = y2[:, :, ::4, ::4].contiguous().numpy()
    g["y2_stats"] = np.array([y2.mean().item(), y2.std().item(), y2.abs().max().item()])
    np.savez_compressed(OUT / "model.npz", **g)
    print("model.npz: y stats", g["y_stats"], "y2 stats", g["y2_stats"], "tokens", g["tokens_stats"])


if __name__ == "__main__":
    main()
