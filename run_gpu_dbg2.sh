python -m pytest tests/test_gpu_elementwise.py -q -p no:cacheprovider -x 2>&1 | tail -2
python tools/elem_bench.py --only corrupt --iters 10 2>&1 | tail -2
