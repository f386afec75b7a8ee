"""Generate tests/golden/ingest.npz by RUNNING THE REFERENCE's EuroSATMultiSpectral.__getitem__ (src/data/dataset.py:112-192)
on small grayscale PNGs written to a temporary directory (rasterio is not installed here, so the reference takes its own PIL
path, dataset.py:137-144: grayscale replicated to 13 bands; max > 10 -> /10000; bilinear 64 -> 192; clamp; (x-0.5)/0.25).

    python oracle/make_golden_ingest.py       # needs /root/reference and Pillow; the GPU box never runs this
"""
import logging
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch
from PIL import Image

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, "/root/reference")
logging.disable(logging.WARNING)
from src.data.dataset import EuroSATMultiSpectral  # noqa: E402

rng = np.random.default_rng(7)
imgs = [rng.integers(0, 256, size=(64, 64), dtype=np.uint8) for _ in range(3)]
imgs.append((np.add.outer(np.arange(64), np.arange(64)) * 2).astype(np.uint8))        # a smooth ramp
with tempfile.TemporaryDirectory() as d:
    for i, a in enumerate(imgs):
        Image.fromarray(a, mode="L").save(Path(d) / f"img_{i:02d}.png")
    ds = EuroSATMultiSpectral(d, split="train", target_size=192, noise_transform=None, train_split=1.0)
    order = [int(p.stem.split("_")[1]) for p in ds.samples]
    clean = torch.stack([ds[i][0] for i in range(len(ds))])
dn = np.stack([np.repeat(imgs[k][None], 13, axis=0) for k in order]).astype(np.uint16)
assert all(torch.equal(clean[:, 0], clean[:, c]) for c in range(13))       # replicated grayscale: every band is identical
np.savez_compressed(ROOT / "tests" / "golden" / "ingest.npz", dn=dn[:, :1], clean=clean[:, :1].numpy())   # band 0 only (size)
print("wrote ingest.npz", dn.shape, clean.shape, float(clean.min()), float(clean.max()))
