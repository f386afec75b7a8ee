# ncu --set full captures of the other decoder kernels (stage 3 conv in CTA-pair mode, the ConvTranspose pixel-shuffle GEMMs)
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  "$@" > gpurun_out/plain_$name.log 2>&1 || { echo "plain $name FAILED"; return; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/r2_prof_$name -f "$@" > gpurun_out/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
}
cap conv3 gemm_tc5 2 python tools/gemm_bench.py --only conv3 --iters 2
cap convt4 gemm_tc5 2 python tools/gemm_bench.py --only convt4 --iters 2
cap convt3 gemm_tc5 2 python tools/gemm_bench.py --only convt3 --iters 2
ls -la gpurun_out/r2_prof_conv3.ncu-rep gpurun_out/r2_prof_convt*.ncu-rep | awk '{print $5, $9}'
