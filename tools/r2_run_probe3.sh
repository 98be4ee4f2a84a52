#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python tests/probes/train_curve_probe3.py > gpurun_out/${1:-p}_curve3.txt 2>&1; echo rc=$?; cat gpurun_out/${1:-p}_curve3.txt | cut -c1-200 | tail -42
