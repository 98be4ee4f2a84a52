"""usage: launch_times.py <ncu launch csv> <kernel substring that starts a step> [min us]: the last step's kernels and their times."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
L = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[hdr + 1:] if len(r) > vi]
idx = [i for i, (k, _) in enumerate(L) if sys.argv[2] in k]
floor = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
tot = 0.0
for k, v in L[idx[-1]:]:
    if v / 1e3 >= floor:
        print(f"{v / 1e3:9.1f} us  {k[:90]}")
    tot += v
print("total ms", tot / 1e6, "launches", len(L) - idx[-1])
