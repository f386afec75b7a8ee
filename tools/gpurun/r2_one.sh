mkdir -p gpurun_out
timeout 900 python tools/train_storage_check.py 2>&1 | tee gpurun_out/train_storage_check.txt | tail -100
timeout 900 python -m pytest tests/test_gpu_train.py -q -p no:cacheprovider -x -k "attention_bwd" 2>&1 | tail -3
