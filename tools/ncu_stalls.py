#!/usr/bin/env python3
"""Per-source-line stall-reason breakdown of an ncu --set full --import-source on report.
usage: ncu_stalls.py report.ncu-rep file-substring lo hi"""
import csv, subprocess, sys, collections, io
rep, want, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "File Path" and want in rows[i][1]:
        hdr = rows[i + 2]
        cols = [k for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        names = [hdr[k][6:] for k in cols]
        agg = collections.OrderedDict()
        line = None
        j = i + 3
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"):
            r = rows[j]; j += 1
            if len(r) < len(hdr): continue
            if r[0] != "": line = int(r[0]) if r[0].isdigit() else None; src = r[1]
            if line is None or not (lo <= line <= hi) or r[2] == "": continue
            a = agg.setdefault(line, [src.strip()[:60]] + [0] * len(cols))
            for n, k in enumerate(cols):
                try: a[1 + n] += int(r[k] or 0)
                except ValueError: pass
        for ln, a in agg.items():
            tot = sum(a[1:])
            if tot < 20: continue
            top = sorted(zip(a[1:], names), reverse=True)[:4]
            print(f"{ln:4d} {tot:6d} " + " ".join(f"{n}:{v}" for v, n in top if v) + " | " + a[0])
        i = j
    else:
        i += 1
