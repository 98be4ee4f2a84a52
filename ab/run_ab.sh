#!/bin/bash
# A/B on one box: alternate ab/new.so with ab/base.so; prints stage times of each run.
# usage: ab/run_ab.sh [extra bench args]
L=wakeword_jupyterlab_b200/libwakeword_b200.so
cp $L ab/keep.so
for rep in 1 2 3; do
  for v in base new; do
    cp ab/$v.so $L
    python bench.py --no-cpu-baseline --no-e2e --no-secondary "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']), round(d['ms_per_step'],2), d.get('stage_ms_per_step'))"
  done
done
cp ab/keep.so $L
