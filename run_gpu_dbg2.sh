python tools/train_b64_check.py 64 2>&1 | tail -2
python tools/train_b64_check.py 8 2>&1 | tail -1
