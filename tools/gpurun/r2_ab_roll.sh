# same-box A/B of the rolling-row stage-4 kernels: MSD_NO_ROLL=1 (the tap-major kernels, as at the start of this session) vs the default, alternating
mkdir -p gpurun_out
for i in 1 2; do
  MSD_NO_ROLL=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-scene > gpurun_out/ab_tap_$i.json 2>/dev/null
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-scene > gpurun_out/ab_roll_$i.json 2>/dev/null
done
python - <<'PY'
import json
for tag in ("tap_1", "roll_1", "tap_2", "roll_2"):
    b = json.load(open(f"gpurun_out/ab_{tag}.json"))
    print(tag, "value %.0f  ms/step %.2f  e2e %.0f" % (b["value"], b["ms_per_step"], b["e2e"]["value"]), b["clocks"]["reasons"], "psnr %.6f" % b["quality"]["psnr"])
PY
