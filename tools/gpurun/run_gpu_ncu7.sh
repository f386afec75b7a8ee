mkdir -p gpurun_out
for k in qkv fc2 conv1 conv3 conv4; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 2 -c 1 -o gpurun_out/prof_$k -f python tools/gemm_bench.py --only $k --iters 2 > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 360 -c 260 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu.log 2>&1; echo "ncu list rc=$?"
