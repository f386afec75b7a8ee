mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -q -p no:cacheprovider -x > gpurun_out/t_model.log 2>&1; echo "model rc=$?"; tail -n 4 gpurun_out/t_model.log
timeout 120 python tools/head_bench.py 2>&1 | tail -2
