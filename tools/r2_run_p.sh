#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
T=${1:-r2p}
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -m gpu -x -q -k "forward or score or large_weights or full_size" > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/${T}_pytest.log
for P in 1 0; do
WW_CONV12_PAIR=$P timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e > $O/${T}_score_pair$P.json 2> $O/${T}_score_pair$P.err; echo "pair=$P rc=$?"; tail -2 $O/${T}_score_pair$P.err
python - <<PY
import json
try:
    d=json.load(open("$O/${T}_score_pair$P.json"))
    print("pair=$P", d["value"], d["ms_per_step"], d["stage_ms_per_step"])
except Exception as e: print("no json", e)
PY
done
