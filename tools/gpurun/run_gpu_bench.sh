mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_elementwise.py -q -p no:cacheprovider -k "philox or loss" > gpurun_out/t_elem2.log 2>&1; echo "elem2 rc=$?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/profile_tc5.json > gpurun_out/bench_tc5.json 2> gpurun_out/bench_tc5.err; echo "bench tc5 rc=$?"; cat gpurun_out/bench_tc5.json; tail -n 5 gpurun_out/bench_tc5.err
timeout 900 python bench.py --steps 5 --warmup 3 --backend 1 --no-cpu-baseline --profile-out gpurun_out/profile_hmma.json > gpurun_out/bench_hmma.json 2> gpurun_out/bench_hmma.err; echo "bench hmma rc=$?"; cat gpurun_out/bench_hmma.json; tail -n 5 gpurun_out/bench_hmma.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cat gpurun_out/bench_ref.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 360 -c 260 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"
tail -n 3 gpurun_out/t_elem2.log
