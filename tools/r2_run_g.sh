#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
T=${1:-r2g}
CMD="python bench.py --workload logmel --clips 4096 --steps 1 --warmup 3"
$CMD > $O/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:logmel_tc' -s 3 -c 1 -f -o $O/${T}_lm $CMD > $O/ncu2.log 2>&1
tail -3 $O/ncu2.log
