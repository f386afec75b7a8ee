mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 360 -c 260 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"
tail -n 2 gpurun_out/plain.log | cut -c1-300
