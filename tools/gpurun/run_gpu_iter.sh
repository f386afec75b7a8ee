mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -q -p no:cacheprovider -x -k "head or forward_vs_oracle or independence" > gpurun_out/t_head.log 2>&1; echo "head rc=$?"; tail -n 4 gpurun_out/t_head.log
timeout 120 python tools/head_bench.py 2>&1 | tail -4
