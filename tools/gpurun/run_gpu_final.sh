mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -p no:cacheprovider -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -n 3 gpurun_out/t_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/profile_v4.json > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err; echo "bench rc=$?"; cat gpurun_out/bench_v4.json | cut -c1-600; tail -n 2 gpurun_out/bench_v4.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_v4.json 2>/dev/null; cat gpurun_out/bench_ref_v4.json | cut -c1-400
python tools/gemm_bench.py > gpurun_out/gemm_bench_v4.txt 2>&1; tail -n 30 gpurun_out/gemm_bench_v4.txt
python tools/elem_bench.py > gpurun_out/elem_bench_v4.txt 2>&1; cat gpurun_out/elem_bench_v4.txt
python tools/attn_bench.py > gpurun_out/attn_bench_v4.txt 2>&1; cat gpurun_out/attn_bench_v4.txt
python tools/head_bench.py > gpurun_out/head_bench_v4.txt 2>&1; cat gpurun_out/head_bench_v4.txt
# ncu: full captures of the new kernels, then the launch list of the bench command (the plain run above exited 0)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc5 -s 2 -c 1 -o gpurun_out/prof_attn -f python tools/attn_bench.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:metrics_fused -s 2 -c 1 -o gpurun_out/prof_metrics -f python tools/elem_bench.py --only metrics > gpurun_out/ncu_metrics.log 2>&1; echo "ncu metrics rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:corrupt_kernel -s 2 -c 1 -o gpurun_out/prof_corrupt -f python tools/elem_bench.py --only corrupt > gpurun_out/ncu_corrupt.log 2>&1; echo "ncu corrupt rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5_kernel -s 3 -c 1 -o gpurun_out/prof_head -f python tools/head_bench.py > gpurun_out/ncu_head.log 2>&1; echo "ncu head rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 350 -c 250 --csv --log-file gpurun_out/launches_v4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
