# profiles of the bandwidth-bound passes of the training step (one ncu --set full capture each, after a clean plain run)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -p no:cacheprovider -x -k "gelu or layernorm" 2>&1 | tail -3
cap() {  # name, kernel regex, skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  "$@" > gpurun_out/plain_$name.log 2>&1 || { echo "plain $name FAILED"; return; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/r2_prof_$name -f "$@" > gpurun_out/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
}
cap bn_bwd1 bn_act_bwd1 3 python tools/bn_bench.py --only 192 --iters 2
cap bn_bwd2 bn_act_bwd2 3 python tools/bn_bench.py --only 192 --iters 2
cap bn_fwd bn_act_fwd 3 python tools/bn_bench.py --only 192 --iters 2
cap ln_bwd ln_bwd_kernel 1 python tools/bn_bench.py --iters 2
cap gelu_bwd_cs gelu_bwd_colsum 1 python tools/bn_bench.py --iters 2
python tools/bn_bench.py 2>&1 | grep -E "gelu|layernorm"
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
