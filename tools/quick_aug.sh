#!/bin/bash
# quick check of the augment kernel on one B200 (run under gpurun): its parity tests, then the stage split of one bench run
python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -x -q -m gpu -k "augment or shift or resample or pcm or snr or normalise or score_with" 2>&1 | tail -5
python bench.py --no-cpu-baseline --no-e2e --steps 5 --warmup 3 "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), d.get('stage_ms_per_step'))"
