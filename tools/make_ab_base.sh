#!/bin/bash
# Build the library of the committed HEAD into ab/base.so (the "A" side of tools/run_ab.sh).
set -e
rm -rf /tmp/ab_base && git worktree add -f /tmp/ab_base HEAD -q
(cd /tmp/ab_base && python -c "from wakeword_jupyterlab_b200 import build; build.build()")
mkdir -p ab && cp /tmp/ab_base/wakeword_jupyterlab_b200/libwakeword_b200.so ab/base.so
git worktree remove --force /tmp/ab_base
