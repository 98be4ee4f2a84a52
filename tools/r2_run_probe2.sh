#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 600 python tests/probes/train_curve_probe2.py > gpurun_out/${1:-p}_curve2.txt 2>&1; echo rc=$?; cat gpurun_out/${1:-p}_curve2.txt | cut -c1-400 | tail -52
