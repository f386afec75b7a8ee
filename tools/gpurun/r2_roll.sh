# rolling-row 48-channel conv (csrc/conv_roll.cu): parity test, then A/B against the tap-major kernel on the stage-4 shape
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -q -p no:cacheprovider -x -k "rolling" 2>&1 | tail -5
{
echo "== 192x192x256";        python tools/roll_bench.py 192 192 256
echo "== 192x192x256 skip";   python tools/roll_bench.py 192 192 256 skip
echo "== full strips only (192 x 128 x 384)"; python tools/roll_bench.py 192 128 384
echo "== pair strips only (192 x 64 x 768)";  python tools/roll_bench.py 192 64 768
} 2>&1 | grep -v "^$" | tee gpurun_out/roll_bench.txt
