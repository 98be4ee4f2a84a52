#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
N=${1:-2}
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q > $O/r2_dist_pytest.log 2>&1; echo "dist pytest rc=$?"; tail -6 $O/r2_dist_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload train --steps 10 --warmup 3 > $O/r2_bench_train_${N}gpu.json 2> $O/r2_bench_train_${N}gpu.err; echo "train bench rc=$?"
cut -c1-420 $O/r2_bench_train_${N}gpu.json; tail -2 $O/r2_bench_train_${N}gpu.err
