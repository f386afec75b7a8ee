mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 2 -c 1 -o gpurun_out/prof_convt4 -f python tools/gemm_bench.py --only convt4 --iters 2 > gpurun_out/ncu_convt4.log 2>&1; echo "ncu rc=$?"
