mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_elementwise.py -q -p no:cacheprovider > gpurun_out/t_elem.log 2>&1; echo "elem rc=$?"
timeout 600 python -m pytest tests/test_gpu_model.py -q -p no:cacheprovider -k "mma_sync or layernorm or attention" > gpurun_out/t_hmma.log 2>&1; echo "hmma rc=$?"
timeout 600 python -m pytest tests/test_gpu_model.py -q -p no:cacheprovider -k "tcgen05 and linear" > gpurun_out/t_tc5_lin.log 2>&1; echo "tc5 linear rc=$?"
timeout 600 python -m pytest tests/test_gpu_model.py -q -p no:cacheprovider -k "tcgen05 and not linear" > gpurun_out/t_tc5_rest.log 2>&1; echo "tc5 rest rc=$?"
timeout 600 python -m pytest tests/test_gpu_model.py -q -p no:cacheprovider -k "folded or independence" > gpurun_out/t_model2.log 2>&1; echo "model2 rc=$?"
for f in t_elem t_hmma t_tc5_lin t_tc5_rest t_model2; do echo "=== $f"; tail -n 15 gpurun_out/$f.log; done
