mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head1x1 -s 3 -c 1 -o gpurun_out/prof_head -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_head.log 2>&1; echo "ncu rc=$?"
