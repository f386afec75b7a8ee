mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -p no:cacheprovider -x -k "wgrad" 2>&1 | tail -4
echo "== N = 96 / 48 tiles"; timeout 300 python tools/bn_bench.py --wgrad 2>&1 | tee gpurun_out/wgrad_bench.txt
echo "== N = 64 tiles"; MSD_WGRAD_N64=1 timeout 300 python tools/bn_bench.py --wgrad 2>&1 | tee -a gpurun_out/wgrad_bench.txt
