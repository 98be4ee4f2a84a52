#!/bin/bash
# whole GPU suite + the default bench line (what the driver runs at round end)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
T=${1:-full}
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<'PY'
import json,sys,glob,os
t=os.environ.get("T","full")
PY
python -c "
import json
r=json.load(open('gpurun_out/${T}_bench.json'))
print({k:r[k] for k in ('value','ms_per_step','e2e','gpu_launches') if k in r})
print('roofline',r.get('roofline'))
s=r.get('secondary',{})
for k,v in s.items(): print(k,{kk:v[kk] for kk in ('value','ms_per_step','unit') if kk in v}, v.get('roofline',{}).get('frac'))
print('stages', r.get('config',{}).get('stage_ms'))
"
