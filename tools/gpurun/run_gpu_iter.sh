mkdir -p gpurun_out
python -m msd_b200.sweep --tiles 27000 2>/dev/null | tail -1
python -m msd_b200.sweep --tiles 27000 2>/dev/null | tail -1
python tools/scene_bench.py 2>/dev/null | tail -1
