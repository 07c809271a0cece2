"""Top stalled SASS instructions of one kernel from `ncu -i rep --page source --csv --kernel-name K > file.csv` (diagnostic).
    python tools/ncu_hot.py file.csv [n]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ix = {c: i for i, c in enumerate(h)}
data = [r for r in rows[hi + 1:] if len(r) == len(h) and r[ix["# Samples"]].isdigit()]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:n]:
    st = sorted(((int(r[ix[c]]), c[6:]) for c in stalls), reverse=True)[:2]
    print(r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(9), r[ix["Source"]].strip()[:80].ljust(80), st)
