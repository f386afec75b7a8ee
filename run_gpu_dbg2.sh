python tools/scene_bench.py 2>&1 | tail -2 | cut -c1-400
