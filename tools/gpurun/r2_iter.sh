# round 2 iteration run: GPU tests, smoke, micro-benchmarks, training-step profile, short bench with the per-launch event profile
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -p no:cacheprovider -m gpu > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -n 12 gpurun_out/t_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
python tools/elem_bench.py 2>&1 | tee gpurun_out/elem_bench.txt
timeout 600 python tools/train_bench.py > gpurun_out/train_bench.txt 2>&1; grep -E "^\{|bn_act|colsum|gemm_tn|ln_bwd|attention_bwd|loss_|Memcpy|elementwise" gpurun_out/train_bench.txt | cut -c1-70,150-215 | head -24
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/profile.json > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
