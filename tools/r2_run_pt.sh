#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for P in 1 0; do
echo "== pair=$P"
WW_CONV12_PAIR=$P WW_TC_TRACE=1 timeout 300 python bench.py --clips 2048 --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e 2>&1 >/dev/null | grep -A26 "conv12 trace" | tail -27 | cut -c1-150
done
