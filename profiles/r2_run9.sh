#!/bin/bash
# N=1: wavefront grid = 148 - reserved SMs; interleaved repeats to see the run-to-run noise
for rep in 1 2; do
for r in 0 2 4 8 16; do
  NVB_WAVEX_RESERVED_SMS=$r python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('rep $rep reserved $r value', round(j['value'],1), 'compute us', round(j['stages']['esdf/integrate/compute']['ms_per_frame']*1000,1), 'with_color', round(j['with_color']['value'],1))"
done
done
