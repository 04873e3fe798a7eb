#!/bin/bash
# developer sweep: planned-mode segments per row for small batches (PRHF_PLAN_NSEG=0 lets the planner choose)
for P in ${SWEEP_P:-2 4 8 16 23}; do
  for ns in ${SWEEP_NS:-0 1 2 3 4 5 6 8 10}; do
    PRHF_PLAN_NSEG=$ns python tools/gpu_probe.py --case $P X 20000 10 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('P', $P, 'nseg', $ns, 'ms %.4f gpts %.1f live %d' % (d['ms_med'], d['gpts_per_s'], d['live_rows']))"
  done
done
