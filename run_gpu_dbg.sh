echo 'MSD_DBG_DT=7 MSD_DBG_L2=3'
MSD_DBG_TMA=6 MSD_DBG_DT=7 MSD_DBG_L2=3 CUDA_LAUNCH_BLOCKING=1 timeout 120 python - <<'PY' 2>&1 | grep -v Warning | tail -12
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
from msd_b200.training.metrics import fused_metrics
a = torch.rand(1, 13, 192, 192, device='cuda') * 4 - 2
b = a + 0.05 * torch.randn_like(a)
t = time.time()
try:
    per, sums, raw = fused_metrics(b, a)
    torch.cuda.synchronize()
    print("ok", per.cpu(), time.time() - t)
except Exception as e:
    print("ERR", repr(e)[:300], time.time() - t)
PY

echo 'MSD_DBG_DT=2'
MSD_DBG_TMA=6 MSD_DBG_DT=2 CUDA_LAUNCH_BLOCKING=1 timeout 120 python - <<'PY' 2>&1 | grep -v Warning | tail -12
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
from msd_b200.training.metrics import fused_metrics
a = torch.rand(1, 13, 192, 192, device='cuda') * 4 - 2
b = a + 0.05 * torch.randn_like(a)
t = time.time()
try:
    per, sums, raw = fused_metrics(b, a)
    torch.cuda.synchronize()
    print("ok", per.cpu(), time.time() - t)
except Exception as e:
    print("ERR", repr(e)[:300], time.time() - t)
PY

echo 'MSD_DBG_DT=7 MSD_DBG_SW=3'
MSD_DBG_TMA=6 MSD_DBG_DT=7 MSD_DBG_SW=3 CUDA_LAUNCH_BLOCKING=1 timeout 120 python - <<'PY' 2>&1 | grep -v Warning | tail -12
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
from msd_b200.training.metrics import fused_metrics
a = torch.rand(1, 13, 192, 192, device='cuda') * 4 - 2
b = a + 0.05 * torch.randn_like(a)
t = time.time()
try:
    per, sums, raw = fused_metrics(b, a)
    torch.cuda.synchronize()
    print("ok", per.cpu(), time.time() - t)
except Exception as e:
    print("ERR", repr(e)[:300], time.time() - t)
PY

echo 'MSD_DBG_DT=3 MSD_DBG_SW=3 MSD_DBG_L2=3'
MSD_DBG_TMA=6 MSD_DBG_DT=3 MSD_DBG_SW=3 MSD_DBG_L2=3 CUDA_LAUNCH_BLOCKING=1 timeout 120 python - <<'PY' 2>&1 | grep -v Warning | tail -12
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
from msd_b200.training.metrics import fused_metrics
a = torch.rand(1, 13, 192, 192, device='cuda') * 4 - 2
b = a + 0.05 * torch.randn_like(a)
t = time.time()
try:
    per, sums, raw = fused_metrics(b, a)
    torch.cuda.synchronize()
    print("ok", per.cpu(), time.time() - t)
except Exception as e:
    print("ERR", repr(e)[:300], time.time() - t)
PY

