"""Summarise an .ncu-rep (one kernel, --set full --import-source on): headline metrics + where warps stall.
    python tools/ncu_stalls.py gpurun_out/prof_x.ncu-rep [--top 20]
Reads the report through `ncu -i ... --page raw|source --csv` (works without a GPU)."""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
top_n = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 18


def page(name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


raw = page("raw")
hdr, units, vals = raw[0], raw[1], raw[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__cycles_elapsed.max", "launch__grid_size", "launch__registers_per_thread",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h:75s} {v[:90]} {u}")

src = page("source")
print(src[0][1][:120])
h = src[1]
ix = {k: i for i, k in enumerate(h)}
data = src[2:]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot)
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
agg = Counter()
for r in data:
    for s in stalls:
        agg[s] += int(r[ix[s]])
print("stall mix:", ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.0f}%" for k, v in agg.most_common(7)))
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top_n]:
    st = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    print(r[ix["# Samples"]].rjust(7), r[ix["Instructions Executed"]].rjust(10), r[ix["Source"]].strip()[:80].ljust(80), st)
c, n = Counter(), Counter()
for r in data:
    k = int(r[ix["Instructions Executed"]])
    c[k] += int(r[ix["# Samples"]]); n[k] += 1
print("by execution count (count, samples, #instructions):")
for k, v in sorted(c.items(), key=lambda x: -x[1])[:8]:
    print("  ", k, v, n[k])
