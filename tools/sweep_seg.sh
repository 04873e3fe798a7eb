#!/bin/bash
# developer sweep: single-profile step time vs tile length
for sl in 1024 1536 2048 2560 2858 3072 3334 4096 5120 6668 10240 20000; do
  PRHF_SEG_LEN=$sl python bench.py --steps 30 --warmup 5 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('seg_len', $sl, 'ms_per_step %.4f'%d['ms_per_step'], 'e2e_ms %.4f'%d['e2e']['ms_per_step'])"
done
