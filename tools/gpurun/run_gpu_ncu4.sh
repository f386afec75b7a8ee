mkdir -p gpurun_out
python tools/gemm_bench.py --only conv4 --iters 2 > gpurun_out/plain_c4.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 2 -c 1 -o gpurun_out/prof_conv4n python tools/gemm_bench.py --only conv4 --iters 2 > gpurun_out/ncu_c4.log 2>&1; echo "ncu conv4 rc=$?"
