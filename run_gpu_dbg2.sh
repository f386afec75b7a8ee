python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_model.py -q -p no:cacheprovider -x 2>&1 | tail -3
python tools/elem_bench.py --only metrics --iters 10 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train --profile-out gpurun_out/profile_tc5.json 2>/dev/null | python -c "
import json,sys; b=json.loads(sys.stdin.read()); print(b['value'], b['ms_per_step'], b['roofline']['other_ms_per_step'], b['roofline']['ms_per_step'])"
