mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -q -p no:cacheprovider -x -k "attention" > gpurun_out/t_attn.log 2>&1; echo "attn rc=$?"; tail -n 3 gpurun_out/t_attn.log
timeout 120 python tools/attn_bench.py 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc5 -s 2 -c 1 -o gpurun_out/prof_attn -f python tools/attn_bench.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu rc=$?"
