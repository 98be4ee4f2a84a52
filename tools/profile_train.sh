#!/bin/bash
# Training-step profile on ONE B200 (run under gpurun): tests, the config-5 bench line, the ncu launch list and one
# full-set capture of the five tensor-core kernels of a step (ncu only after the same command exited 0 without it),
# then the precision / training-curve probes.     usage: tools/profile_train.sh <tag>
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T=${1:-r2}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_train_tc_gpu.py tests/test_train_gpu.py -q -m gpu > $O/${T}_train_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/${T}_train_tests.log
CMD="python bench.py --workload train --steps 10 --warmup 3"
timeout 300 $CMD > $O/${T}_bench_train_1gpu.json 2> $O/${T}_train.err; echo "bench rc=$?"; cut -c1-200 $O/${T}_bench_train_1gpu.json
WW_TRAIN_KERNEL=fp32 timeout 300 python bench.py --workload train --steps 3 --warmup 3 > $O/${T}_bench_train_fp32kernels_1gpu.json 2>/dev/null; cut -c1-200 $O/${T}_bench_train_fp32kernels_1gpu.json
CMD1="python bench.py --workload train --steps 1 --warmup 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_train_launches.csv $CMD1 > $O/ncu_t1.log 2>&1; echo "ncu list rc=$?"
python tools/launch_times.py $O/${T}_train_launches.csv pad_logmel 0 > $O/${T}_train_step_kernels.txt; grep -E "total|kernel<" $O/${T}_train_step_kernels.txt
K='regex:wgrad_tap_kernel|wgrad_kernel|dgrad_kernel|conv3_kernel|conv12_kernel'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 21 -c 7 -f -o $O/${T}_train_full $CMD1 > $O/ncu_t2.log 2>&1; echo "ncu full rc=$?"; tail -2 $O/ncu_t2.log
timeout 600 python tests/probes/tf32_grad_probe.py > $O/${T}_train_precision.txt 2>&1; echo "probe rc=$?"
timeout 600 python tests/probes/train_curve_probe2.py > $O/${T}_train_curves.txt 2>&1
timeout 900 python tests/probes/train_curve_probe3.py >> $O/${T}_train_curves.txt 2>&1; echo "curves rc=$?"
timeout 900 python bench.py --no-secondary --no-cpu-baseline > $O/${T}_score_check.json 2>/dev/null; python -c "
import json; r=json.load(open('$O/${T}_score_check.json')); print(r['value'], r['ms_per_step'], r['stage_ms_per_step'])"
