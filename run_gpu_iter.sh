mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_elementwise.py -q -p no:cacheprovider -x -k "metrics or tracker" > gpurun_out/t_metrics.log 2>&1; echo "metrics rc=$?"; tail -n 3 gpurun_out/t_metrics.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/bench_tc5.json 2> gpurun_out/bench_tc5.err; echo "bench tc5 rc=$?"; python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_tc5.json'))
print({k:b[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', b['e2e'])
PY
tail -n 3 gpurun_out/bench_tc5.err
