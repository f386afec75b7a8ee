# training-step iteration: the training kernels' tests, the bandwidth-bound passes alone, then the step profile
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py -q -p no:cacheprovider -x > gpurun_out/t_train.log 2>&1; echo "train tests rc=$?"; tail -n 15 gpurun_out/t_train.log
timeout 300 python tools/bn_bench.py 2>&1 | tee gpurun_out/bn_bench.txt
timeout 600 python tools/train_bench.py > gpurun_out/train_bench.txt 2>&1; grep -E "^\{|bn_act|colsum|gemm_tn|ln_bwd|attention_bwd|gelu|gemm_tc5" gpurun_out/train_bench.txt | cut -c1-70,150-215 | head -30
