mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train --profile-out gpurun_out/profile_v4.json > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err; echo "bench rc=$?"; python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_v4.json'))
print({k:b[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', b['e2e']['value'])
r=b['roofline']; print({k:r[k] for k in ('achieved','frac','ms_per_step','share_of_step')}); print(r['family']['achieved'], r['family']['frac'], r['family']['ms_per_step'], r['family']['share_of_step']); print(r['other_ms_per_step'])
print(b['clocks'])
PY
python tools/elem_bench.py 2>&1 | grep -E "corrupt|metrics"
