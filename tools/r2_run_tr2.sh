#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T=${1:-tr2}
timeout 600 python tests/probes/tf32_grad_probe.py > gpurun_out/${T}_probe.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/${T}_probe.txt | tail -20
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"
