"""How far is the CUDA training step from fp32 autograd, compared with an emulation of its bf16 storage points in the same torch
graph (oracle/model_storage.py)?  Prints, per gradient tensor, rel-RMS(library, fp32), rel-RMS(storage oracle, fp32) and
rel-RMS(library, storage oracle).      python tools/train_storage_check.py [--batch 8]"""
import argparse, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import recipes
from msd_b200.models import SatMAERestoration
from oracle import model as omodel, model_storage

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--seed", type=int, default=11)
args = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(args.seed)
model = SatMAERestoration(recipes.base_config(), block_layout="fallback")
sd = recipes.perturb_decoder_stats({k: v.clone() for k, v in model.state_dict().items()})
model.load_state_dict(sd)
sd = {k: v.cuda() for k, v in sd.items()}
clean, noisy = recipes.eurosat_tiles(args.batch, seed=9)
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


def rel(a, b):
    return ((a - b).pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-30)).item()


y_st, ref_st = oracle(model_storage.restoration_forward_train)
y_32, ref_32 = oracle(omodel.restoration_forward_train)
y = model(noisy)
((y - clean) ** 2).mean().backward()
print(f"output: lib-fp32 {rel(y.detach(), y_32):.4f}  storage-fp32 {rel(y_st, y_32):.4f}  lib-storage {rel(y.detach(), y_st):.4f}")
rows = []
for n, p in model.named_parameters():
    if n in frozen or n.endswith(("upsample.0.bias", "head.0.bias")):
        continue
    rows.append((n, rel(p.grad, ref_32[n].grad), rel(ref_st[n].grad, ref_32[n].grad), rel(p.grad, ref_st[n].grad)))
for n, a, b, c in rows:
    print(f"{n:58s} lib-fp32 {a:.4f}  storage-fp32 {b:.4f}  ratio {a / (b + 1e-9):5.2f}  lib-storage {c:.4f}")
ratios = sorted(a / (b + 1e-9) for _, a, b, _ in rows)
print("ratio lib/storage error vs fp32: min %.2f median %.2f max %.2f" % (ratios[0], ratios[len(ratios) // 2], ratios[-1]))
