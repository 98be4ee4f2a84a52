#!/usr/bin/env python3
"""Per-source-line roll-up of an ncu `--set full --import-source on` report.

usage: ncu_lines.py report.ncu-rep [kernel-substring] [top-N]
Sums `Instructions Executed`, stall samples and shared-memory wavefronts of every SASS instruction over the CUDA
source line it maps to (`ncu --page source --csv --print-source cuda,sass`), so the hot lines of a kernel can be
read here without a GPU.
"""
import csv, subprocess, sys, collections, io

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "File Path":
        path, func = rows[i][1], rows[i + 1][1]
        hdr = rows[i + 2]
        j = i + 3
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"):
            body.append(rows[j]); j += 1
        i = j
        if want not in func:
            continue
        col = {n: k for k, n in reversed(list(enumerate(hdr)))}
        ci, cs, cw = col["Instructions Executed"], col["# Samples"], col["L1 Wavefronts Shared"]
        agg = collections.OrderedDict()
        line, src = None, ""
        tot_i = tot_s = tot_w = 0
        for r in body:
            if len(r) < len(hdr):
                continue
            if r[0] != "":
                line, src = r[0], r[1]
            if r[2] == "":
                continue
            try:
                n, s, w = int(r[ci] or 0), int(r[cs] or 0), int(r[cw] or 0)
            except ValueError:
                continue
            a = agg.setdefault((path.split("/")[-1], line), [src.strip()[:90], 0, 0, 0, 0])
            a[1] += n; a[2] += s; a[3] += w; a[4] += 1
            tot_i += n; tot_s += s; tot_w += w
        print(f"== {func}\n   {path}: {tot_i} warp-instructions, {tot_s} samples, {tot_w} shared wavefronts")
        for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1 if "--by-inst" in sys.argv else 2])[:top]:
            print(f"{f}:{ln:>4} inst {100*a[1]/max(tot_i,1):5.1f}%  samp {100*a[2]/max(tot_s,1):5.1f}%  smem-wf {100*a[3]/max(tot_w,1):5.1f}%  sass {a[4]:4d} | {a[0]}")
    else:
        i += 1
