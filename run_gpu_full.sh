mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -p no:cacheprovider -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -n 4 gpurun_out/t_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/profile_tc5.json > gpurun_out/bench_tc5.json 2> gpurun_out/bench_tc5.err; echo "bench rc=$?"; python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_tc5.json'))
print({k:b[k] for k in ('value','ms_per_step')}, 'e2e', b['e2e']['value'], 'gemm TF/s', b['roofline']['achieved'], 'gemm ms', b['roofline']['ms_per_step'], b['roofline']['other_ms_per_step'], b['clocks'], 'train', b.get('train_step'), 'cpu', b.get('cpu_baseline'))
PY
tail -n 3 gpurun_out/bench_tc5.err
python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; cat gpurun_out/gemm_bench.log
