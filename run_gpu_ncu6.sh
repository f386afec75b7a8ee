mkdir -p gpurun_out
python tools/elem_bench.py --only attention --iters 3 > gpurun_out/plain_att.log 2>&1 && cat gpurun_out/plain_att.log && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 2 -c 1 -o gpurun_out/prof_att -f python tools/elem_bench.py --only attention --iters 3 > gpurun_out/ncu_att.log 2>&1; echo "ncu rc=$?"
