#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 600 python tests/probes/train_curve_probe.py ${2:-3e-3} > gpurun_out/${1:-p}_curve.txt 2>&1; echo rc=$?; cat gpurun_out/${1:-p}_curve.txt | cut -c1-700
