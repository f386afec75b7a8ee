mkdir -p gpurun_out
python tools/graph_probe.py 2>&1 | tee gpurun_out/graph_probe.txt
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_sweep.py tests/test_gpu_scene.py -q -p no:cacheprovider -k "benchmark_batch or 27000 or writer" 2>&1 | tail -5
