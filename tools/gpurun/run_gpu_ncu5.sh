mkdir -p gpurun_out
for k in ${KERNELS:-convt4}; do
  python tools/gemm_bench.py --only $k --iters 2 > gpurun_out/plain_$k.log 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 2 -c 1 -o gpurun_out/prof_$k -f python tools/gemm_bench.py --only $k --iters 2 > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
