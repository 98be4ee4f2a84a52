#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
N=${1:-8}
nvidia-smi topo -m 2>/dev/null | head -14
for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302" $d/class 2>/dev/null; then echo "$(basename $d) numa=$(cat $d/numa_node)"; fi; done | head -8
ls /sys/devices/system/node/ 2>/dev/null | head; nproc; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $O/r2_bench_${N}gpu.json 2> $O/r2_bench_${N}gpu.err; echo "bench rc=$?"
tail -3 $O/r2_bench_${N}gpu.err
python - <<PY
import json
d=json.load(open("$O/r2_bench_${N}gpu.json"))
print(d["n_gpus"], d["value"], d["ms_per_step"]); print(d["e2e"]); print(d["strong"])
for k,v in (d.get("secondary") or {}).items(): print(k, v["value"], v["ms_per_step"], v["config"].get("replicas_in_sync"))
PY
