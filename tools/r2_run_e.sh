#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
T=${1:-r2e}
timeout 300 python -m pytest tests/test_round2_gpu.py -m gpu -x -q -k "tensor_core_logmel" > $O/${T}_pytest_logmel.log 2>&1; echo "pytest logmel rc=$?"
tail -5 $O/${T}_pytest_logmel.log
WW_LOGMEL_KERNEL=tc timeout 600 python bench.py --workload logmel --steps 10 --warmup 3 > $O/${T}_lm_tc.json 2> $O/${T}_lm_tc.err; echo "tc rc=$?"; cat $O/${T}_lm_tc.json | cut -c1-200; tail -5 $O/${T}_lm_tc.err
WW_LOGMEL_KERNEL=fft timeout 600 python bench.py --workload logmel --steps 10 --warmup 3 > $O/${T}_lm_fft.json 2>/dev/null; echo "fft rc=$?"; cat $O/${T}_lm_fft.json | cut -c1-200
WW_LOGMEL_KERNEL=tc WW_TC_TRACE=1 timeout 300 python bench.py --workload logmel --clips 2048 --steps 1 --warmup 3 > /dev/null 2> $O/${T}_trace.err; tail -30 $O/${T}_trace.err
if [ "$2" = "full" ]; then timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/${T}_pytest.log; fi
