"""Driver for tests/test_gpu_train.py::test_data_parallel_backward_averages_gradients (torchrun, one rank per GPU): the gradients
`loss.backward()` leaves with SatMAERestoration.enable_data_parallel() equal the mean over ranks of the un-reduced gradients."""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import torch.distributed as dist

import recipes
from msd_b200.models import SatMAERestoration
from msd_b200.training import CombinedLoss


def main(out_path):
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(42)
    model = SatMAERestoration(recipes.base_config(), block_layout="fallback").cuda().train()
    clean, noisy = recipes.eurosat_tiles(4, seed=50 + rank)              # every rank its own micro-batch
    clean, noisy = clean.cuda(), noisy.cuda()
    crit = CombinedLoss(1.0, 0.1)

    def grads(dp):
        model.enable_data_parallel(enabled=dp)
        model.zero_grad(set_to_none=True)
        crit(model(noisy), clean).backward()
        return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}

    local_g = grads(False)
    dp_g = grads(True)
    worst = 0.0
    for n, g in local_g.items():
        mean = g.clone()
        dist.all_reduce(mean, op=dist.ReduceOp.SUM)
        mean /= world
        scale = mean.abs().max().item() + 1e-12
        # the two backward passes differ by the fp32-atomic order of the weight-gradient split-K merges only
        worst = max(worst, (dp_g[n] - mean).abs().max().item() / scale)
    same = torch.tensor([worst], device="cuda")
    dist.all_reduce(same, op=dist.ReduceOp.MAX)
    if rank == 0:
        json.dump({"world": world, "tensors": len(local_g), "worst_rel": same.item()}, open(out_path, "w"))
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
