#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
N=${1:-2}
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -x -q > $O/r2_dist_pytest.log 2>&1; echo "dist pytest rc=$?"; tail -15 $O/r2_dist_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $O/r2_bench_${N}gpu.json 2> $O/r2_bench_${N}gpu.err; echo "bench rc=$?"
tail -3 $O/r2_bench_${N}gpu.err
python - <<PY
import json
d=json.load(open("$O/r2_bench_${N}gpu.json"))
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"], d["strong"])
for k,v in (d.get("secondary") or {}).items(): print(k, v["value"], v["ms_per_step"], v["config"].get("replicas_in_sync"))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 | cut -c1-300
