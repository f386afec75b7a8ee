mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_elementwise.py -q -p no:cacheprovider -x -k "metrics" > gpurun_out/t_metrics.log 2>&1; echo "metrics rc=$?"; tail -n 5 gpurun_out/t_metrics.log
timeout 120 python tools/elem_bench.py --only metrics 2>&1 | tail -2
