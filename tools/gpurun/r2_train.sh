# training-step iteration: the training kernels' tests, the bandwidth-bound passes alone, then the step profile
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py -q -p no:cacheprovider -x > gpurun_out/t_train.log 2>&1; echo "train tests rc=$?"; tail -n 15 gpurun_out/t_train.log
timeout 300 python tools/bn_bench.py 2>&1 | tee gpurun_out/bn_bench.txt
timeout 600 python tools/train_bench.py > gpurun_out/train_bench.txt 2>&1; grep -E "^\{|bn_act|colsum|gemm_tn|ln_bwd|attention_bwd|gelu|gemm_tc5" gpurun_out/train_bench.txt | cut -c1-70,150-215 | head -30
cap() {  # name, kernel regex, skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/r2_prof_$name -f "$@" > gpurun_out/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
}
cap bn_bwd1 bn_act_bwd1 3 python tools/bn_bench.py --only 192 --iters 2
cap ln_bwd ln_bwd_kernel 1 python tools/bn_bench.py --iters 2
