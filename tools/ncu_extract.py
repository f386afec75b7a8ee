"""Turn this round's `ncu --set full` captures (gpurun_out/r2_prof_*.ncu-rep, one kernel launch each) into
profiles/r2_ncu_metrics.json -- the file bench.py reads its profiler-only figures from (DRAM bytes per launch, tensor-pipe and
issue-slot activity), so that nothing measured by ncu is typed into bench.py -- and profiles/r2_ncu_full_summaries.txt.

    python tools/ncu_extract.py [gpurun_out]

Reads each report through `ncu -i ... --page raw --csv` (works without a GPU)."""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
src = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "gpurun_out"

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "launch__grid_size",
        "launch__registers_per_thread", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "dram__bytes_write.sum.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "sm__inst_executed_pipe_fma.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
        "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}


def load(rep: Path):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return None
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in WANT:
            try:
                d[h] = float(v.replace(",", "")) * UNIT.get(u, 1.0)
            except ValueError:
                d[h] = v
    return d


# label -> (report stem, elements one launch processes, for instructions-per-element figures)
TILE = 13 * 192 * 192
CAPTURES = {
    "qkv (LayerNorm folded)": ("r2_prof_qkv_ln", None), "proj + LayerNorm sums": ("r2_prof_proj_stats", None),
    "conv 192x192x48": ("r2_prof_conv4", None), "attention (tcgen05)": ("r2_prof_attn", None),
    "conv 192x192x48 (rolling rows)": ("r2_prof_conv4_roll", None), "conv 192x192x48 + skip (rolling rows)": ("r2_prof_conv4_skip_roll", None),
    "head conv + fused 1x1 (rolling rows)": ("r2_prof_head_roll", None),
    "corrupt_kernel": ("r2_prof_corrupt", 256 * TILE), "metrics_fused_kernel": ("r2_prof_metrics", 256 * TILE),
    "loss_fwd_kernel": ("r2_prof_loss_fwd", 256 * TILE), "loss_bwd_kernel": ("r2_prof_loss_bwd", 256 * TILE),
}

facts = {"note": "round-2 ncu --set full --clock-control none captures, one launch per kernel at 256 tiles (tools/gpurun/r2_ncu.sh), "
                 "extracted by tools/ncu_extract.py", "tensor_pipe_active_pct": {}}
lines = []
for label, (stem, elems) in CAPTURES.items():
    rep = src / f"{stem}.ncu-rep"
    if not rep.exists():
        continue
    d = load(rep)
    if not d:
        continue
    lines.append(f"=================== {label}  ({stem}.ncu-rep)")
    for k in WANT:
        if k in d:
            lines.append(f"{k:75s} {d[k]}")
    tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
    traffic = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    entry = {"ncu_duration_us": round(d.get("gpu__time_duration.sum", 0.0) * 1e6, 2), "dram_bytes_per_launch": traffic,
             "issue_active_pct": d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "dram_pct_of_peak": round(d.get("dram__bytes_read.sum.pct_of_peak_sustained_elapsed", 0.0) + d.get("dram__bytes_write.sum.pct_of_peak_sustained_elapsed", 0.0), 1)}
    if elems:
        entry["thread_instructions_per_element"] = round(d.get("smsp__inst_executed.sum", 0.0) * 32 / elems, 1)
    if tp is not None and "kernel" not in label and label not in ("corrupt_kernel", "metrics_fused_kernel"):
        facts["tensor_pipe_active_pct"][label] = round(tp, 1)
    if label == "qkv (LayerNorm folded)":
        facts["traffic_dominant_bytes"] = traffic
        facts["traffic_dominant_kernel"] = ("gemm_tc5_pair_kernel<256,0> on the qkv shape (M=36864, N=2304, K=768, folded-LayerNorm epilogue), "
                                            "dram__bytes_read.sum + dram__bytes_write.sum of one launch; algorithmic 230 MB (A 56.6 + W 3.5 + out 169.9)")
    facts[label if label.endswith("_kernel") else "gemm: " + label] = entry
if "loss_fwd_kernel" in facts or "loss_bwd_kernel" in facts:
    facts["loss_kernels"] = {k: facts[k] for k in ("loss_fwd_kernel", "loss_bwd_kernel") if k in facts}
(ROOT / "profiles" / "r2_ncu_metrics.json").write_text(json.dumps(facts, indent=1))
(ROOT / "profiles" / "r2_ncu_full_summaries.txt").write_text(
    "# round 2: ncu --set full --clock-control none --import-source on, one launch per kernel at B = 256 (tools/gpurun/r2_ncu.sh)\n"
    + "\n".join(lines) + "\n")
print(json.dumps(facts, indent=1)[:3000])
