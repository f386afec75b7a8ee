# same-box A/B: round-1 final tree (unpacked under _r1/, its own .so) vs this tree, the driver's bench command, alternating
mkdir -p gpurun_out
for i in 1 2; do
  (cd _r1 && python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train) > gpurun_out/ab_r1_$i.json 2>/dev/null
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-scene > gpurun_out/ab_r2_$i.json 2>/dev/null
done
python - <<'PY'
import json
for tag in ("r1_1", "r2_1", "r1_2", "r2_2"):
    b = json.load(open(f"gpurun_out/ab_{tag}.json"))
    print(tag, "value %.0f  ms/step %.2f  e2e %.0f" % (b["value"], b["ms_per_step"], b["e2e"]["value"]), b.get("e2e_f32", {}).get("value"), b["clocks"]["reasons"])
PY
