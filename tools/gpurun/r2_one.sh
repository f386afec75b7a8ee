mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -q -p no:cacheprovider -x -k "attention" 2>&1 | tail -5
echo "== four threads per row"; timeout 300 python tools/attn_bench.py 2>&1 | tee gpurun_out/attn_bench.txt
echo "== thread per row"; MSD_ATTN_THREAD_PER_ROW=1 timeout 300 python tools/attn_bench.py 2>&1 | tee -a gpurun_out/attn_bench.txt
