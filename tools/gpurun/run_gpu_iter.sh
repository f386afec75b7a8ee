mkdir -p gpurun_out
timeout 300 python tools/train_bench.py 2>&1 | grep -E "^\{|bn_act|colsum|gemm_tn" | cut -c1-200 | head -9
timeout 600 python -m pytest tests/test_gpu_train.py -q -p no:cacheprovider -x 2>&1 | tail -2
