# rolling-row 48-channel conv (csrc/conv_roll.cu): parity tests, determinism, A/B timings on the stage-4 shape
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -p no:cacheprovider -x -k "rolling or head_conv or batch_independence" > gpurun_out/t_roll.log 2>&1; echo "tests rc=$?"; tail -1 gpurun_out/t_roll.log
{
echo "== 192x192x256";        python tools/roll_bench.py 192 192 256
echo "== 192x192x256 skip";   python tools/roll_bench.py 192 192 256 skip
echo "== full strips only (192 x 128 x 384)"; python tools/roll_bench.py 192 128 384 | tail -1
echo "== pair strips only (192 x 64 x 768)";  python tools/roll_bench.py 192 64 768
echo "== head rolling";       python tools/head_bench.py 2>&1 | tail -4 | head -1
echo "== tap-major";          MSD_NO_ROLL=1 python tools/roll_bench.py 192 192 256 | tail -1
} 2>&1 | grep -v "^$" | tee gpurun_out/roll_bench.txt
