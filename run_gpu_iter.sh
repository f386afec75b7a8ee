mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -p no:cacheprovider -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -n 3 gpurun_out/t_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train --profile-out gpurun_out/profile_tc5.json > gpurun_out/bench_tc5.json 2> gpurun_out/bench_tc5.err; echo "bench tc5 rc=$?"; python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_tc5.json'))
print({k:b[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', b['e2e']['value'], 'gemm TF/s', b['roofline']['achieved'], 'gemm ms', b['roofline']['ms_per_step'], b['roofline']['other_ms_per_step'])
p=json.load(open('gpurun_out/profile_tc5.json'))
for r in p['launches'][-6:]: print(r['i'], r['kind'], round(r['ms'],4), round(r['gflop'],1), r['tflops'] and round(r['tflops']))
PY
