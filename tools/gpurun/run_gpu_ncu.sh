mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5_kernel -s 3 -c 1 -o gpurun_out/prof_head2 -f python tools/head_bench.py > gpurun_out/ncu_head2.log 2>&1; echo "ncu rc=$?"
