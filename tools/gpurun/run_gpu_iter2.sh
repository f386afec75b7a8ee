mkdir -p gpurun_out
python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; cat gpurun_out/gemm_bench.log
timeout 600 python -m pytest tests/test_gpu_model.py -q -p no:cacheprovider -x -k "not mma_sync" > gpurun_out/t_model.log 2>&1; echo "model tests rc=$?"; tail -n 4 gpurun_out/t_model.log
