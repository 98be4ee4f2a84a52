#!/bin/bash
# round 2, GPU call A: GPU test suite, bench line (stage split), first run of the tensor-core STFT draft
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
timeout 120 ./tools/logmel_tc_draft 4096 > gpurun_out/r2a_tc_draft.log 2>&1; echo "draft rc=$?" >> gpurun_out/r2a_tc_draft.log
tail -5 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_bench.json | head -c 3000; cat gpurun_out/r2a_tc_draft.log
