"""Three optimizer steps with graph-replayed vs eager weight packing must give identical losses and parameters."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import recipes
from msd_b200.models import SatMAERestoration
from msd_b200.training import CombinedLoss

def run(no_graph):
    if no_graph: os.environ["MSD_NO_PACK_GRAPH"] = "1"
    else: os.environ.pop("MSD_NO_PACK_GRAPH", None)
    torch.manual_seed(42)
    model = SatMAERestoration(recipes.base_config(), block_layout="fallback").cuda().train()
    opt = torch.optim.AdamW(model.get_optimizer_param_groups(1e-3, 0.05), fused=True)
    crit = CombinedLoss(1.0, 0.1)
    clean, noisy = recipes.eurosat_tiles(8, seed=3)
    clean, noisy = clean.cuda(), noisy.cuda()
    losses = []
    for _ in range(4):
        loss = crit(model(noisy), clean)
        loss.backward()
        opt.step(); opt.zero_grad(set_to_none=True)
        losses.append(loss.item())
    return losses, torch.cat([p.detach().flatten() for p in model.parameters()]).double().sum().item()

a = run(True); b = run(False)
print("eager ", a); print("graph ", b)
