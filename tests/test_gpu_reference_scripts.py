"""The reference's own entry points driven through the drop-in on a GPU (SURVEY 8(a) row a18, 8(b) "wiring"): scripts/evaluate.py
and the Trainer run UNCHANGED from the staged copy under baseline/_ref (tools/stage_ref.sh; git-ignored, ships with the gpurun
snapshot) with their hot-path names rebound to libmsd_b200 by msd_b200.run.  Skipped when the reference was not staged."""
import json
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import yaml

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "baseline" / "_ref"
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (REF / "scripts" / "evaluate.py").exists(),
                                                  reason="baseline/_ref not staged (bash tools/stage_ref.sh)")]


def _make_dataset(root: Path, n: int = 20):
    """n grayscale 64x64 PNG tiles in 10 class folders: the reference's loader replicates them to 13 bands (dataset.py:137-144)"""
    from PIL import Image
    rng = np.random.default_rng(0)
    for i in range(n):
        d = root / f"Class{i % 10}"
        d.mkdir(parents=True, exist_ok=True)
        yy, xx = np.mgrid[0:64, 0:64]
        img = 110 + 60 * np.sin(xx / (3 + i % 5)) * np.cos(yy / (4 + i % 3)) + rng.normal(0, 8, (64, 64))
        Image.fromarray(np.clip(img, 0, 255).astype(np.uint8)).save(d / f"tile_{i:03d}.png")


def _config(tmp: Path, micro: int, accum: int) -> Path:
    cfg = yaml.safe_load((REF / "configs" / "base.yaml").read_text())
    cfg["data"]["root_dir"] = str(tmp / "data" / "EuroSAT_MS")
    cfg["data"]["pin_memory"] = False
    cfg["training"]["micro_batch_size"] = micro
    cfg["training"]["gradient_accumulation_steps"] = accum
    cfg["training"]["effective_batch_size"] = micro * accum
    cfg["training"]["epochs"] = 1
    p = tmp / "cfg.yaml"
    p.write_text(yaml.safe_dump(cfg))
    return p


def _env():
    env = dict(os.environ)
    env["PYTHONPATH"] = str(ROOT) + os.pathsep + env.get("PYTHONPATH", "")
    return env


def test_reference_evaluate_script_runs_unchanged_on_the_dropin(tmp_path):
    """python -m msd_b200.run <ref>/scripts/evaluate.py --checkpoint ...: load_config, SatMAERestoration(config), load_state_dict
    (strict, reference key set), get_dataloaders (noise applied per sample in __getitem__), evaluate_model -- reference code
    except for the rebound hot-path names (scripts/evaluate.py:37-65)."""
    import recipes
    from msd_b200.models import SatMAERestoration
    _make_dataset(tmp_path / "data" / "EuroSAT_MS")
    cfg = _config(tmp_path, micro=4, accum=1)
    torch.manual_seed(42)
    ckpt = tmp_path / "ckpt.pth"
    torch.save({"model_state_dict": SatMAERestoration(recipes.base_config(), block_layout="fallback").state_dict()}, ckpt)
    r = subprocess.run([sys.executable, "-m", "msd_b200.run", "--log-level", "INFO", str(REF / "scripts" / "evaluate.py"), "--config", str(cfg),
                        "--checkpoint", str(ckpt)], cwd=tmp_path, env=_env(), capture_output=True, text=True, timeout=600)
    log = r.stdout + r.stderr
    assert r.returncode == 0, log[-3000:]
    assert "rebound" in log
    m = re.search(r"PSNR: ([-\d.]+) dB", log)
    s = re.search(r"SSIM: ([-\d.]+)", log)
    a = re.search(r"SAM: ([-\d.]+)", log)
    assert m and s and a, log[-3000:]
    psnr, ssim, sam = float(m.group(1)), float(s.group(1)), float(a.group(1))
    assert np.isfinite([psnr, ssim, sam]).all() and 0 < psnr < 60 and -1 <= ssim <= 1 and 0 <= sam <= 180


def test_reference_trainer_steps_on_the_dropin(tmp_path):
    """Two optimizer steps of the reference's Trainer (fp16 autocast + GradScaler + clip 1.0 + AdamW, trainer.py:141-213) with
    the library's autograd node underneath: finite losses, trainable weights move, frozen blocks do not, validation runs."""
    _make_dataset(tmp_path / "data" / "EuroSAT_MS", n=20)          # 16 train / 4 val -> 4 micro-batches of 4, accumulation 2
    cfg = _config(tmp_path, micro=4, accum=2)
    out = tmp_path / "out.json"
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "ref_driver.py"), str(REF), str(cfg), str(out)], cwd=tmp_path,
                       env=_env(), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    res = json.loads(out.read_text())
    assert res["patched"] >= 15
    assert np.isfinite(res["train"]["loss"]) and res["train"]["loss"] > 0
    assert res["global_step"] == 2
    assert res["changed"] > 100 and res["frozen_changed"] == []
    assert np.isfinite(list(res["val"].values())).all() and res["val"]["psnr"] > 0
