mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py -q -p no:cacheprovider -x > gpurun_out/t_model.log 2>&1; echo "model+train rc=$?"; tail -n 3 gpurun_out/t_model.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/profile_tc5.json > gpurun_out/bench_tc5.json 2> gpurun_out/bench_tc5.err; echo "bench tc5 rc=$?"; cat gpurun_out/bench_tc5.json; tail -n 3 gpurun_out/bench_tc5.err
