mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc5 -s 2 -c 1 -o gpurun_out/prof_attn -f python tools/attn_bench.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_attn.log
