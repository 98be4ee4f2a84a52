#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
T=${1:-r2s}
for K in tc fft; do
WW_LOGMEL_KERNEL=$K timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e > $O/${T}_score_$K.json 2> $O/${T}_score_$K.err; echo "$K rc=$?"
python - <<PY
import json
d=json.load(open("$O/${T}_score_$K.json"))
print("$K", d["value"], d["ms_per_step"], d["stage_ms_per_step"])
PY
done
