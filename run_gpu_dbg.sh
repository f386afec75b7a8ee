timeout 600 python -m pytest tests/test_gpu_model.py -q -p no:cacheprovider 2>&1 | tail -3
timeout 900 python -m pytest tests -q -p no:cacheprovider -m gpu 2>&1 | tail -3
