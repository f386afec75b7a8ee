mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -p no:cacheprovider -x > gpurun_out/t_train.log 2>&1; echo "train tests rc=$?"; tail -n 2 gpurun_out/t_train.log
python tools/train_bench.py > gpurun_out/train_bench.log 2>&1; grep -v "^$" gpurun_out/train_bench.log | grep "ms/step\|bn_\|colsum" | awk '{print substr($0,1,72), substr($0,150,62)}' | head -12
