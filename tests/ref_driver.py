"""Driver executed in a subprocess by tests/test_gpu_reference_scripts.py: the reference's UNMODIFIED `Trainer`
(baseline/_ref/src/training/trainer.py: fp16 autocast + GradScaler + clip_grad_norm_ + AdamW, as scripts/train.py wires it)
running on top of libmsd_b200's SatMAERestoration / CombinedLoss / MetricsTracker / AddSensorNoise through msd_b200.run.rebind.

    python tests/ref_driver.py <ref_root> <config.yaml> <out.json>
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402


def main(ref_root: str, config_path: str, out_path: str):
    from msd_b200.run import rebind
    patched = rebind(Path(ref_root))
    from src.data import get_dataloaders                     # the reference's own loader (PIL fallback reads the PNG tiles)
    from src.models import SatMAERestoration                  # rebound -> libmsd_b200
    from src.training import CombinedLoss, Trainer            # Trainer is the reference's own class
    from src.utils import load_config
    import msd_b200.models
    assert SatMAERestoration is msd_b200.models.SatMAERestoration
    assert Trainer.__module__ == "src.training.trainer"
    config = load_config(config_path)
    config["model"]["encoder"]["pretrained_path"] = None
    torch.manual_seed(config["seed"])
    train_loader, val_loader = get_dataloaders(config)
    model = SatMAERestoration(config)
    opt_cfg = config["training"]["optimizer"]
    groups = model.get_optimizer_param_groups(lr=opt_cfg["lr"], weight_decay=opt_cfg["weight_decay"])      # scripts/train.py:144-152
    optimizer = torch.optim.AdamW(groups, betas=tuple(opt_cfg["betas"]))
    criterion = CombinedLoss(mse_weight=config["training"]["loss"]["mse_weight"], ssim_weight=config["training"]["loss"]["ssim_weight"])
    trainer = Trainer(model=model, optimizer=optimizer, criterion=criterion, config=config, scheduler=None, device="cuda")
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    train_metrics = trainer.train_epoch(train_loader, 1)
    val_metrics = trainer.validate(val_loader)
    changed = sorted(n for n, p in model.named_parameters() if not torch.equal(before[n].to(p.device), p.detach()))
    frozen_changed = [n for n in changed if any(n.startswith(f"encoder.blocks.{i}.") for i in config["model"]["encoder"]["freeze_layers"])]
    json.dump({"patched": len(patched), "train": {k: float(v) for k, v in train_metrics.items()},
               "val": {k: float(v) for k, v in val_metrics.items()}, "changed": len(changed), "frozen_changed": frozen_changed,
               "global_step": int(trainer.global_step), "scaler_scale": float(trainer.scaler.get_scale())}, open(out_path, "w"))


if __name__ == "__main__":
    main(*sys.argv[1:4])
