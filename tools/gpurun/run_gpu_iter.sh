mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -p no:cacheprovider -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -n 3 gpurun_out/t_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/profile_v4.json > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err; echo "bench rc=$?"; python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_v4.json'))
print({k:b[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', b['e2e']['value'])
r=b['roofline']; print({k:r[k] for k in ('achieved','frac','ms_per_step','share_of_step')}); print(r['family']); print(r['other_ms_per_step'])
print(b['clocks'], b.get('train_step',{}).get('value'))
PY
