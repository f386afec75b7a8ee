timeout 300 python -m pytest tests/test_gpu_model.py -q -p no:cacheprovider -x -k "conv3x3" 2>&1 | tail -15
python tools/gemm_bench.py --only conv3; python tools/gemm_bench.py --only conv3_skip; python tools/gemm_bench.py --only conv4; python tools/gemm_bench.py --only conv4_skip
