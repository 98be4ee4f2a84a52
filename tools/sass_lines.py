"""Per-source-line SASS opcode histogram of an object file (nvdisasm -gi), innermost inlined line.
usage: python tools/sass_lines.py build/foo.o FILE.cu [lo hi]"""
import collections, os, re, subprocess, sys, tempfile

obj, fname = sys.argv[1], sys.argv[2]
lo, hi = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, 10 ** 9)
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
    txt = ""
    for f in sorted(os.listdir(d)):
        txt += subprocess.run(["nvdisasm", "-gi", os.path.join(d, f)], capture_output=True, text=True).stdout
cnt = collections.Counter()
cur, fresh = None, True
for ln in txt.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            fresh = False
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        fresh = True
        if cur and cur[0] == fname and lo <= cur[1] <= hi:
            cnt[(cur[1], m.group(2).split(".")[0])] += 1
for (line, op), c in sorted(cnt.items()):
    print(f"{line:5d} {op:12s} {c}")
