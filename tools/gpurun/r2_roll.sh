# rolling-row 48-channel conv (csrc/conv_roll.cu): parity tests, then A/B against the tap-major kernel on the stage-4 shape
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -p no:cacheprovider -x -k "rolling or head_conv or batch_independence" 2>&1 | tail -5
{
echo "== 192x192x256";        python tools/roll_bench.py 192 192 256
echo "== 192x192x256 skip";   python tools/roll_bench.py 192 192 256 skip
echo "== head rolling"; python tools/head_bench.py 2>&1 | tail -4
} 2>&1 | grep -v "^$" | tee gpurun_out/roll_bench.txt
