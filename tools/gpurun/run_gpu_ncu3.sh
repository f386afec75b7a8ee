mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv
python tools/gemm_bench.py > gpurun_out/gemm_bench_a.log 2>&1; cat gpurun_out/gemm_bench_a.log
MSD_NO_MERGE3=1 python tools/gemm_bench.py --only conv4 2>&1; MSD_NO_MERGE3=1 python tools/gemm_bench.py --only conv3 2>&1
python tools/gemm_bench.py --only conv4 --iters 2 > gpurun_out/plain_c4.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 2 -c 1 -o gpurun_out/prof_conv4m python tools/gemm_bench.py --only conv4 --iters 2 > gpurun_out/ncu_c4.log 2>&1; echo "ncu conv4 rc=$?"
python tools/gemm_bench.py --only convt4 --iters 2 > gpurun_out/plain_ct4.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 2 -c 1 -o gpurun_out/prof_convt4 python tools/gemm_bench.py --only convt4 --iters 2 > gpurun_out/ncu_ct4.log 2>&1; echo "ncu convt4 rc=$?"
