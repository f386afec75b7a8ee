mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"; cat gpurun_out/bench_n2.json | cut -c1-2500; tail -5 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2>/dev/null; echo "ref n2 rc=$?"; cat gpurun_out/bench_ref_n2.json | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 -m msd_b200.sweep --tiles 2000 --batch 250 > gpurun_out/sweep_n2.json 2> gpurun_out/sweep_n2.err; echo "sweep rc=$?"; cat gpurun_out/sweep_n2.json; tail -3 gpurun_out/sweep_n2.err
python -m msd_b200.sweep --tiles 2000 --batch 250 > gpurun_out/sweep_n1.json 2>/dev/null; cat gpurun_out/sweep_n1.json
