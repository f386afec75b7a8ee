# training-step tests + step timing (the 48-channel raw convs / data gradients now run the rolling-row kernel)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_gpu_model.py -q -p no:cacheprovider -x > gpurun_out/t_train.log 2>&1; echo "train+model tests rc=$?"; tail -3 gpurun_out/t_train.log
timeout 600 python tools/train_bench.py > gpurun_out/train_bench.txt 2>&1; grep -E "^\{" gpurun_out/train_bench.txt | cut -c1-400
MSD_NO_ROLL=1 timeout 600 python tools/train_bench.py > gpurun_out/train_bench_noroll.txt 2>&1; grep -E "^\{" gpurun_out/train_bench_noroll.txt | cut -c1-400
