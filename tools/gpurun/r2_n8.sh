# round 2, N GPUs of one box: the driver's scaling command (weak scaling), reference arm under torchrun, the 27,000-tile sweep (C3)
mkdir -p gpurun_out
N=${NGPU:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "n$N rc=$?"
wc -l gpurun_out/bench_n$N.json; cut -c1-300 gpurun_out/bench_n$N.json; echo; tail -n 4 gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 -m msd_b200.sweep --tiles 27000 --batch 256 > gpurun_out/sweep_n$N.json 2> gpurun_out/sweep_n$N.err; echo "sweep rc=$?"; cat gpurun_out/sweep_n$N.json
