mkdir -p gpurun_out
N=${NGPU:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "n$N rc=$?"
head -c 400 gpurun_out/bench_n$N.json; echo; grep -v "^\*\|OMP_NUM" gpurun_out/bench_n$N.err | tail -5 | cut -c1-250
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 -m msd_b200.sweep --tiles 27000 > gpurun_out/sweep_n$N.json 2> gpurun_out/sweep_n$N.err; echo "sweep rc=$?"; tail -1 gpurun_out/sweep_n$N.json | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 tools/scene_bench.py > gpurun_out/scene_n$N.json 2> gpurun_out/scene_n$N.err; echo "scene rc=$?"; tail -1 gpurun_out/scene_n$N.json | cut -c1-400; tail -2 gpurun_out/scene_n$N.err | cut -c1-300
