#!/bin/bash
# quick check of the head kernels on one B200 (run under gpurun): logit parity tests, then the stage split with both head kernels
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -x -q -m gpu -k "logit or forward or score or decision or predict or weights" 2>&1 | tail -8
for k in tc fp32; do
  WW_HEAD_KERNEL=$k timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-secondary --steps 5 --warmup 3 "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$k', round(d['value']), round(d['ms_per_step'],2), d.get('stage_ms_per_step'))"
done
