#!/bin/bash
# developer A/B on one box: single-profile step of the baseline build vs the working build, interleaved
for rep in 1 2 3; do
for lib in libpyrayhf_b200_base.so libpyrayhf_b200.so; do
  export PRHF_LIB_PATH=$PWD/pyrayhf_b200/csrc/$lib
  [ -f $PRHF_LIB_PATH ] || continue
  a=$(python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-batched | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('single ms %.4f e2e_ms %.4f'%(d['ms_per_step'], d['e2e']['ms_per_step']))")
  echo "$lib | $a"
done
done
