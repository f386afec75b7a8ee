# same-box A/B of programmatic dependent launch over the tensor-core kernel chain: MSD_NO_PDL=1 vs default, alternating
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -p no:cacheprovider -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -n 3 gpurun_out/t_gpu.log
for i in 1 2; do
  MSD_NO_PDL=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-scene > gpurun_out/ab_nopdl_$i.json 2>/dev/null
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-scene > gpurun_out/ab_pdl_$i.json 2>/dev/null
done
python - <<'PY'
import json
for tag in ("nopdl_1", "pdl_1", "nopdl_2", "pdl_2"):
    b = json.load(open(f"gpurun_out/ab_{tag}.json"))
    print(tag, "value %.0f  ms/step %.2f  e2e %.0f" % (b["value"], b["ms_per_step"], b["e2e"]["value"]), b["clocks"]["reasons"], "psnr %.6f" % b["quality"]["psnr"])
PY
MSD_NO_PDL=1 timeout 600 python tools/train_bench.py 2>&1 | grep -E "^\{" | cut -c1-200
timeout 600 python tools/train_bench.py 2>&1 | grep -E "^\{" | cut -c1-200
