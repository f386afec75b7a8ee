mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scene.py tests/test_gpu_sweep.py -q -p no:cacheprovider 2>&1 | tail -4
python tools/e2e_probe.py 2>&1 | tee gpurun_out/e2e_probe.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-scene > gpurun_out/bench20.json 2> gpurun_out/bench20.err; echo "bench rc=$?"; python -c "
import json;b=json.load(open('gpurun_out/bench20.json'));print('value',b['value'],b['ms_per_step'],'e2e',b['e2e']['value'],'f32',b['e2e_f32']['value'],b['clocks'])"
