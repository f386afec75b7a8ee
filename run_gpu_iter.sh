mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -q -p no:cacheprovider -x > gpurun_out/t_model.log 2>&1; echo "model rc=$?"; tail -n 5 gpurun_out/t_model.log
python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; cat gpurun_out/gemm_bench.log
