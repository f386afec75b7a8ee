# round 2, two GPUs: the tests that need NCCL across ranks / a second device, then the 2-rank bench (weak scaling)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -p no:cacheprovider -m gpu -k "two_ranks or data_parallel or second_device or prefetch_then_all_reduce" > gpurun_out/t_gpu_n2.log 2>&1; echo "2-gpu tests rc=$?"; tail -n 4 gpurun_out/t_gpu_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"; cut -c1-400 gpurun_out/bench_n2.json; tail -n 5 gpurun_out/bench_n2.err
