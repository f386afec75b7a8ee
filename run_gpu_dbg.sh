mkdir -p gpurun_out
python tools/bn_cancel_probe.py > gpurun_out/bn_probe.log 2>&1; tail -20 gpurun_out/bn_probe.log
