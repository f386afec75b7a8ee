mkdir -p gpurun_out
python tools/elem_bench.py > gpurun_out/elem_bench.log 2>&1; cat gpurun_out/elem_bench.log
python tools/elem_bench.py --only metrics --iters 1 > gpurun_out/plain_m.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ssim_mse -s 2 -c 1 -o gpurun_out/prof_ssim python tools/elem_bench.py --only metrics --iters 1 > gpurun_out/ncu_m.log 2>&1; echo "ncu ssim rc=$?"
python tools/elem_bench.py --only corrupt --iters 1 > gpurun_out/plain_c.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corrupt_kernel -s 2 -c 1 -o gpurun_out/prof_corrupt python tools/elem_bench.py --only corrupt --iters 1 > gpurun_out/ncu_c.log 2>&1; echo "ncu corrupt rc=$?"
