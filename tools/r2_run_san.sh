#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 120 python tests/probes/train_tc_once.py 3 32 || exit 1
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tests/probes/train_tc_once.py 3 32 > gpurun_out/san_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -15 gpurun_out/san_memcheck.txt | cut -c1-300
