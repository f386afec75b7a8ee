mkdir -p gpurun_out
python tools/train_bench.py > gpurun_out/train_bench.log 2>&1; tail -5 gpurun_out/train_bench.log
