#!/bin/bash
# developer sweep: planned-mode segments per row for the single-profile bench workload
for ns in 0 2 3 4 5 6 8 10 13 19; do
  PRHF_PLAN_NSEG=$ns python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-batched | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('nseg', $ns, 'ms_per_step %.4f'%d['ms_per_step'], 'tile_kernel_ms %.4f'%d['roofline']['kernel_ms'], 'e2e_ms %.4f'%d['e2e']['ms_per_step'])"
done
