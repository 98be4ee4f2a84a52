#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload train --steps 10 --warmup 3 > $O/r2_bench_train_${N}gpu.json 2> $O/r2_bench_train_${N}gpu.err; echo "train bench rc=$?"
cut -c1-200 $O/r2_bench_train_${N}gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > $O/r2_bench_score_${N}gpu.json 2> $O/r2_bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$O/r2_bench_score_${N}gpu.json"))
print(d["n_gpus"], d["value"], d["ms_per_step"]); print(d["e2e"]); print(d["strong"])
for k,v in (d.get("secondary") or {}).items(): print(k, v["value"], v["ms_per_step"], v["config"].get("replicas_in_sync"))
PY
