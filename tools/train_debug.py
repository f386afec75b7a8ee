"""Layer-by-layer comparison of the training forward's kept activations with the fp32 oracle (debugging aid)."""
import ctypes
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import torch.nn.functional as F
import recipes
from msd_b200 import _lib
from msd_b200.models import SatMAERestoration
from oracle import model as om

B = 2
torch.manual_seed(42)
model = SatMAERestoration(recipes.base_config(), block_layout="fallback")
sd = recipes.perturb_decoder_stats({k: v.clone() for k, v in model.state_dict().items()})
model.load_state_dict(sd)
x = recipes.model_input(B, seed=77)
model.cuda().train()
with torch.no_grad():
    y, ws = model._forward_train(x.cuda())
torch.cuda.synchronize()
lib = _lib.load()
offs = (ctypes.c_size_t * 45)()
_lib.check(lib.msd_model_train_layout(model._handle, B, offs, 45))

def view(off, shape):
    n = 1
    for s in shape: n *= s
    return ws[off:off + 2 * n].view(torch.bfloat16).view(*shape).float().cpu()

ref = {k: v.clone() for k, v in sd.items()}
with torch.no_grad():
    tok = om.encoder_forward(ref, x)
    got = view(offs[0], (B * 144, 768))
    print("xnf", (got - tok.reshape(B * 144, 768)).abs().max().item(), tok.abs().max().item())
    feat = tok.transpose(1, 2).reshape(B, 768, 12, 12)
    xx = feat
    hw = 12
    chans = [384, 192, 96, 48]
    i = 0
    for s in range(4):
        c = chans[s]; hw *= 2
        p = f"decoder.upsample_blocks.{i}."
        raw = F.conv_transpose2d(xx, ref[p + "upsample.0.weight"], ref[p + "upsample.0.bias"], stride=2)
        def cmp(name, k, t):
            g = view(offs[1 + 10 * s + k], (B, hw + 2, hw + 2, c))[:, 1:-1, 1:-1].permute(0, 3, 1, 2)
            e = (g - t).abs()
            print(f"stage {s} {name:6s} max {e.max().item():.4f} rms {e.pow(2).mean().sqrt().item():.5f} ref max {t.abs().max().item():.3f} rms {t.pow(2).mean().sqrt().item():.3f}")
        cmp("U_raw", 0, raw)
        xx = F.gelu(om._bn(raw, ref, p + "upsample.1.", training=True))
        cmp("X0", 1, xx)
        i += 1
        for rb in range(2):
            p = f"decoder.upsample_blocks.{i}."
            c1 = F.conv2d(xx, ref[p + "block.0.weight"], None, padding=1)
            cmp(f"C{1+2*rb}", 2 + 4 * rb, c1)
            y1 = F.gelu(om._bn(c1, ref, p + "block.1.", training=True))
            cmp(f"Y{1+rb}", 3 + 4 * rb, y1)
            c2 = F.conv2d(y1, ref[p + "block.3.weight"], None, padding=1)
            cmp(f"C{2+2*rb}", 4 + 4 * rb, c2)
            xx = F.gelu(xx + om._bn(c2, ref, p + "block.4.", training=True))
            cmp(f"Z{1+rb}", 5 + 4 * rb, xx)
            i += 1
