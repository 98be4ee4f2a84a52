#!/bin/bash
# round 2, GPU call B: ncu full capture of the augment and log-mel kernels of one 8,192-clip step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
CMD="python bench.py --clips 8192 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > $O/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:augment_kernel|logmel_kernel' -s 6 -c 2 -f -o $O/r2b_full $CMD > $O/ncu2.log 2>&1
tail -3 $O/ncu2.log
