#!/bin/bash
# training-step checks: tensor-core backward tests, the reference-parity training tests (exact kernels), then the config-5 bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T=${1:-tr}
timeout 900 python -m pytest tests/test_train_tc_gpu.py -q -m gpu -s > gpurun_out/${T}_tc.log 2>&1; echo "tc tests rc=$?"
grep -E "^B |loss after|passed|failed|Error|error" gpurun_out/${T}_tc.log | head -40
timeout 600 python -m pytest tests/test_train_gpu.py -x -q -m gpu > gpurun_out/${T}_train.log 2>&1; echo "train tests rc=$?"
tail -5 gpurun_out/${T}_train.log
timeout 300 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cat gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"
