mkdir -p gpurun_out
for k in qkv fc1 conv1 conv4; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 2 -c 1 -o gpurun_out/prof_$k -f python tools/gemm_bench.py --only $k --iters 2 > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:corrupt_kernel -s 1 -c 1 -o gpurun_out/prof_corrupt -f python tools/elem_bench.py --only corrupt --iters 2 > gpurun_out/ncu_corrupt.log 2>&1; echo "ncu corrupt rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ssim_mse -s 1 -c 1 -o gpurun_out/prof_ssim -f python tools/elem_bench.py --only metrics --iters 2 > gpurun_out/ncu_ssim.log 2>&1; echo "ncu ssim rc=$?"
