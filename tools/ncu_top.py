"""Summarise an ncu report's source page: top stalled SASS instructions.  usage: python tools/ncu_top.py rep.ncu-rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
def f(r, k):
    try: return float(r[idx[k]])
    except Exception: return 0.0
tot = sum(f(r, '# Samples') for r in data)
print(rows[0][1][:100]); print("total samples", tot, "instructions", len(data))
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for r in sorted(data, key=lambda r: -f(r, '# Samples'))[:N]:
    st = sorted(((f(r, c), c) for c in stall_cols), reverse=True)[:2]
    print(f"{r[idx['Address']][-5:]} {f(r,'# Samples'):7.0f} {f(r,'# Samples')/tot:6.1%} exec={f(r,'Instructions Executed'):9.0f} {r[idx['Source']][:80]:80s} {[(int(a), b[6:]) for a, b in st]}")
