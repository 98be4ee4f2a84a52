#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
T=${1:-r2c}
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
tail -15 $O/${T}_pytest.log; python - <<PY
import json
d=json.load(open("$O/${T}_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"] if d.get("e2e") else None, d["stage_ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_step_frac_of_peak"], d["clocks"])
PY
