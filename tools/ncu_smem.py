#!/usr/bin/env python3
"""Shared-memory wavefronts per source line (total / bank-conflict excess / ideal) of one kernel in an ncu report.
usage: ncu_smem.py report.ncu-rep kernel-substring [file-substring]"""
import csv, subprocess, io, collections, sys
rep, want = sys.argv[1], sys.argv[2]
fwant = sys.argv[3] if len(sys.argv) > 3 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "File Path":
        path, func = rows[i][1], rows[i + 1][1]; hdr = rows[i + 2]; j = i + 3; body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"):
            body.append(rows[j]); j += 1
        i = j
        if want not in func or fwant not in path:
            continue
        col = {n: k for k, n in reversed(list(enumerate(hdr)))}
        cw, ce, cid = col["L1 Wavefronts Shared"], col["L1 Wavefronts Shared Excessive"], col["L1 Wavefronts Shared Ideal"]
        agg = collections.OrderedDict(); line = None; src = ''
        for r in body:
            if len(r) < len(hdr): continue
            if r[0] != '': line, src = r[0], r[1]
            if r[2] == '': continue
            try: w, e, d = int(r[cw] or 0), int(r[ce] or 0), int(r[cid] or 0)
            except ValueError: continue
            if w == 0: continue
            a = agg.setdefault(line, [src.strip()[:90], 0, 0, 0]); a[1] += w; a[2] += e; a[3] += d
        tot = sum(a[1] for a in agg.values()) or 1
        print(f"== {func[:70]}  {path.split('/')[-1]}: {tot} wavefronts")
        for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
            print(f"{ln:>4} wf {100*a[1]/tot:5.1f}%  excess {100*a[2]/tot:5.1f}%  ideal {100*a[3]/tot:5.1f}% | {a[0]}")
    else:
        i += 1
