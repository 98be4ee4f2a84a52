#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T=${1:-tr5}
timeout 900 python -m pytest tests/test_train_tc_gpu.py -q -m gpu > gpurun_out/${T}_tc.log 2>&1; echo "tc tests rc=$?"; tail -3 gpurun_out/${T}_tc.log
timeout 300 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cat gpurun_out/${T}_bench.json | cut -c1-330; tail -5 gpurun_out/${T}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"
python tools/launch_times.py gpurun_out/${T}_launches.csv pad_logmel 100
