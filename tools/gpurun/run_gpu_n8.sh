mkdir -p gpurun_out
N=${NGPU:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "n$N rc=$?"
wc -l gpurun_out/bench_n$N.json; head -c 300 gpurun_out/bench_n$N.json; echo
