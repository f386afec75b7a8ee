"""Experiment: does a K-major SWIZZLE_128B UMMA descriptor accept an A window that starts r rows (r*128 B) into a
1024-byte swizzle atom when descriptor bits [49,52) (base offset) are set to r?  (needed to serve the three dx taps of
a 3x3 conv row from ONE TMA box).  Integer operands -> exact compare; the last r rows of each 128-row tile read past
the box and are ignored."""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))

if len(sys.argv) > 1:
    shift, bo = int(sys.argv[1]), int(sys.argv[2])
    os.environ["MSD_DBG_SHIFT"], os.environ["MSD_DBG_BO"] = str(shift), str(bo)
    import torch
    import gpu_helpers as gh
    g = torch.Generator().manual_seed(0)
    M, N, K = 512, 64, 192
    a = torch.randint(-3, 4, (M, K), generator=g).float()
    w = torch.randint(-3, 4, (N, K), generator=g).float()
    out = gh.linear(a.bfloat16().cuda(), w.bfloat16().cuda(), None, out_f32=True).cpu()
    # with the box loaded `shift` rows early and the window starting `shift` rows in, row m must still equal a[m] @ w^T
    ref = a @ w.t()
    rows = torch.arange(M)
    keep = (rows % 128) < 128 - shift
    bad = (out[keep] != ref[keep]).any(dim=1).sum().item()
    print(f"shift={shift} base_offset={bo}: {bad} / {int(keep.sum())} rows differ")
else:
    for shift, bo in ((0, 0), (1, 1), (1, 0), (2, 2), (3, 3), (7, 7), (2, 0)):
        r = subprocess.run([sys.executable, __file__, str(shift), str(bo)], capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr.strip()[-300:], flush=True)
