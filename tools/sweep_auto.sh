#!/bin/bash
# developer check: what the planner picks on small batches
for P in 2 3 4 6 8 12 16 23; do
  python tools/gpu_probe.py --case $P X 20000 10 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('P', $P, 'auto ms %.4f gpts %.1f live %d' % (d['ms_med'], d['gpts_per_s'], d['live_rows']))"
done
