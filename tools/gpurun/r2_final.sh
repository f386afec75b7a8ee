# round 2, end state on one GPU: the driver's sequence (GPU tests, smoke, reference arm, bench) + the event profile, the training-step
# profile and the ncu launch list of the bench command (after the plain bench exited 0)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -p no:cacheprovider -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -n 3 gpurun_out/t_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
timeout 1200 python bench.py --profile-out gpurun_out/profile.json > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
python tools/gemm_bench.py > gpurun_out/gemm_bench.txt 2>&1; tail -n 26 gpurun_out/gemm_bench.txt
timeout 600 python tools/train_bench.py > gpurun_out/train_bench.txt 2>&1; grep -E "^\{" gpurun_out/train_bench.txt | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-scene > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
