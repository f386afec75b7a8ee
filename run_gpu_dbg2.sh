timeout 300 python -m pytest tests/test_gpu_model.py -q -p no:cacheprovider -x -k "linear or conv3x3" 2>&1 | tail -15
