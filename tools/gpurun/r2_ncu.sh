# round 2 profiles: one `ncu --set full` capture per kernel (each command ran clean without ncu first), then the launch list of the
# bench command.  Reports come back in gpurun_out/ and are summarised by tools/ncu_extract.py / tools/ncu_launches.py into profiles/.
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  "$@" > gpurun_out/plain_$name.log 2>&1 || { echo "plain $name FAILED"; return; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/r2_prof_$name -f "$@" > gpurun_out/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
}
cap qkv_ln gemm_tc5_pair 2 python tools/gemm_bench.py --only qkv_ln --iters 2
cap proj_stats gemm_tc5_pair 2 python tools/gemm_bench.py --only proj_stats --iters 2
cap conv4 gemm_tc5_kernel 2 python tools/gemm_bench.py --only conv4 --iters 2
cap attn attention_tc5 2 python tools/attn_bench.py
cap corrupt corrupt_kernel 2 python tools/elem_bench.py --only corrupt --iters 2
cap metrics metrics_fused 2 python tools/elem_bench.py --only metrics --iters 2
cap loss_fwd loss_fwd 2 python tools/elem_bench.py --only loss --iters 2
cap loss_bwd loss_bwd 2 python tools/elem_bench.py --only loss --iters 2
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-scene > gpurun_out/plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-scene > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
