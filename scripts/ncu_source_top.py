"""stdin: `ncu --page source --csv` of one kernel; stdout: its top-N instructions by stall samples,
with executed count, shared-memory wavefronts (ideal / actual), for profiles/."""
import csv
import sys

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rows = list(csv.reader(sys.stdin))
if len(rows) < 3:
    sys.exit(0)
print(rows[0][1][:160] if len(rows[0]) > 1 else rows[0])
hdr = rows[1]
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    data.append(r)
col = {h: i for i, h in enumerate(hdr)}
si, ie = col["Warp Stall Sampling (All Samples)"], col["Instructions Executed"]
wi, wd = col.get("L1 Wavefronts Shared"), col.get("L1 Wavefronts Shared Ideal")
tot = sum(int(r[si] or 0) for r in data)
print(f"instructions {len(data)}  warp-instructions executed {sum(int(r[ie] or 0) for r in data)}  stall samples {tot}")
print(f"{'samples':>8s} {'share':>6s} {'executed':>10s} {'smem wf':>9s} {'ideal':>9s}  SASS")
for r in sorted(data, key=lambda r: -int(r[si] or 0))[:n]:
    print(f"{r[si]:>8s} {100 * int(r[si] or 0) / max(tot, 1):5.1f}% {r[ie]:>10s} {(r[wi] if wi else ''):>9s} {(r[wd] if wd else ''):>9s}  {r[1][:90]}")
