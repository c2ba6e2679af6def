"""Top stall-sample source lines/instructions of a kernel from `ncu -i rep --page source --csv` output (stdin or file)."""
import csv, sys
lines = open(sys.argv[1]).read().splitlines()
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
start = [i for i, l in enumerate(lines) if l.startswith('"Address","Source"')][0]
end = len(lines)
for i in range(start + 1, len(lines)):
    if lines[i].startswith('"Kernel Name"'):
        end = i
        break
rows = list(csv.DictReader(lines[start:end]))
key = "Warp Stall Sampling (All Samples)"
tot = sum(float(r[key] or 0) for r in rows) or 1
stalls = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
agg = {c: sum(float(r[c] or 0) for r in rows) for c in stalls}
print("stall mix:", ", ".join(f"{k[6:]} {v / tot * 100:.0f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:7]))
for r in sorted(rows, key=lambda r: -float(r[key] or 0))[:n]:
    top = max(stalls, key=lambda c: float(r[c] or 0))
    print(f"{float(r[key]) / tot * 100:5.1f}%  {top[6:]:10s} {r['Source'][:100]}")
