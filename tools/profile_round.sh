#!/bin/bash
# Round profile on ONE B200 (run under gpurun): bench lines of every workload, then the ncu launch list and the
# full-set capture of one 8,192-clip step (= one chunk: 8 launches) (each ncu run only after the same command exited 0 without ncu).
# usage: tools/profile_round.sh <tag>      -> gpurun_out/<tag>_*.json / .csv / .ncu-rep
T=${1:-rX}; O=gpurun_out; mkdir -p $O
python bench.py > $O/${T}_bench_score_1gpu.json 2> $O/${T}_score.err
python bench.py --workload logmel --no-cpu-baseline > $O/${T}_bench_logmel_1gpu.json 2>/dev/null
python bench.py --workload stream --no-cpu-baseline > $O/${T}_bench_stream_1gpu.json 2>/dev/null
python bench.py --workload train --no-cpu-baseline > $O/${T}_bench_train_1gpu.json 2>/dev/null
python bench.py --conv-mode fp16 --no-cpu-baseline > $O/${T}_bench_score_fp16mode_1gpu.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_reference_arm.json 2>/dev/null
CMD="python bench.py --clips 8192 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
K='regex:augment_pipe_kernel|augment_kernel|logmel_kernel|conv12_kernel|conv3_kernel|pool_finish|gated_dense|fc_softmax|head_tc_pack'
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 24 -c 8 --csv --log-file $O/${T}_ncu_launches.csv $CMD > $O/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" -s 24 -c 8 -f -o $O/${T}_full $CMD > $O/ncu2.log 2>&1
tail -2 $O/ncu2.log
