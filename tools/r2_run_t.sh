#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
WW_TC_TRACE=1 timeout 300 python bench.py --workload logmel --clips 2048 --steps 1 --warmup 3 > /dev/null 2> gpurun_out/${1:-t}_trace.err
tail -42 gpurun_out/${1:-t}_trace.err
