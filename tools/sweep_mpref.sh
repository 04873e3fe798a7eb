#!/bin/bash
# developer sweep: m-table prefetch variants of the hot loop (built with -DPRHF_MPREF=1|2)
for lib in libpyrayhf_b200.so libpyrayhf_b200_mpref1.so libpyrayhf_b200_mpref2.so libpyrayhf_b200.so; do
  export PRHF_LIB_PATH=$PWD/pyrayhf_b200/csrc/$lib
  [ -f $PRHF_LIB_PATH ] || continue
  a=$(python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-batched | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('single ms %.4f e2e_ms %.4f'%(d['ms_per_step'], d['e2e']['ms_per_step']))")
  b=$(python tools/gpu_probe.py --case 512 X 20000 5 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('batch512 ms %.3f gpts %.1f'%(d['ms_med'], d['gpts_per_s']))")
  c=$(python tools/gpu_probe.py --case 4096 X 200 5 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('n200 ms %.3f gpts %.1f'%(d['ms_med'], d['gpts_per_s']))")
  echo "$lib | $a | $b | $c"
done
