# round 2, second half: `ncu --set full` captures of the rolling-row 48-channel conv (csrc/conv_roll.cu) and, beside it, of the
# tap-major kernel it replaces (MSD_NO_ROLL=1); each command ran clean without ncu first.  Summarised by tools/ncu_extract.py.
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  "$@" > gpurun_out/plain_$name.log 2>&1 || { echo "plain $name FAILED"; return; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/r2_prof_$name -f "$@" > gpurun_out/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
}
cap conv4_roll conv_roll_kernel 2 python tools/gemm_bench.py --only conv4 --iters 2
cap conv4_skip_roll conv_roll_kernel 2 python tools/gemm_bench.py --only conv4_skip --iters 2
cap head_roll conv_roll_kernel 2 python tools/head_bench.py 256
MSD_NO_ROLL=1 cap conv4 gemm_tc5_kernel 2 python tools/gemm_bench.py --only conv4 --iters 2
ls -la gpurun_out/*roll*.ncu-rep gpurun_out/r2_prof_conv4.ncu-rep | awk '{print $5, $9}'
