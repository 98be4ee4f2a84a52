#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest ${@:-tests/test_train_tc_gpu.py} -q -m gpu > gpurun_out/tests.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/tests.log | cut -c1-300
