mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -q -p no:cacheprovider -x -k "attention" > gpurun_out/t_attn.log 2>&1; echo "attn rc=$?"; tail -n 5 gpurun_out/t_attn.log
timeout 120 python tools/attn_bench.py 2>&1 | tail -2
timeout 120 python tools/attn_bench.py 64 2>&1 | tail -2
