"""Training-mode forward + loss at the C4 micro-batch (64) against the fp32 oracle evaluated with torch on the GPU."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import recipes
from msd_b200.models import SatMAERestoration
from msd_b200.training import CombinedLoss
from oracle import model as om

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(42)
model = SatMAERestoration(recipes.base_config(), block_layout="fallback")
sd = {k: v.clone().cuda() for k, v in model.state_dict().items()}
clean, noisy = recipes.eurosat_tiles(32, seed=3)
clean = clean.repeat(B // 32 + 1, 1, 1, 1)[:B].contiguous().cuda(); noisy = noisy.repeat(B // 32 + 1, 1, 1, 1)[:B].contiguous().cuda()
model.cuda().train()
crit = CombinedLoss(1.0, 0.1)
with torch.no_grad():
    y = model(noisy)
    y_ref = om.restoration_forward_train(sd, noisy)
    e = y - y_ref
    print(f"B={B}: out rms {y_ref.pow(2).mean().sqrt():.4f}  err rms {e.pow(2).mean().sqrt():.4f}  max {e.abs().max():.4f}  loss lib {crit(y, clean).item():.5f}  loss oracle-output {crit(y_ref.contiguous(), clean).item():.5f}")

# gradients at this batch size: library backward vs torch autograd of the oracle on the GPU (plain MSE both sides)
frozen = {n for n, p in model.named_parameters() if not p.requires_grad}
ref = {k: v.clone() for k, v in sd.items()}
for k, v in ref.items():
    if v.is_floating_point() and "running_" not in k and k not in frozen:
        v.requires_grad_(True)
((om.restoration_forward_train(ref, noisy) - clean) ** 2).mean().backward()
model.zero_grad(set_to_none=True)
((model(noisy) - clean) ** 2).mean().backward()
rows = []
for n, p in model.named_parameters():
    if n in frozen or n.endswith(("upsample.0.bias", "head.0.bias")):
        continue
    g, r = p.grad, ref[n].grad
    rows.append((((g - r).pow(2).mean().sqrt() / (r.pow(2).mean().sqrt() + 1e-30)).item(), n))
rows.sort(reverse=True)
print("worst rel-rms gradient errors:", [(round(a, 3), b) for a, b in rows[:5]], " median", round(rows[len(rows) // 2][0], 4))
