#!/bin/bash
# clear pass: fused kernel vs select + balanced process; ESDF / decay / pipeline parity tests with the split on
python -m pytest tests/test_gpu_parity.py tests/test_gpu_decay.py tests/test_gpu_bench_pipeline.py tests/test_gpu_esdf_slice.py -x -q -m gpu 2>&1 | tail -3
for v in 0 1; do
  NVB_CLEAR_SPLIT=$v python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('split $v value', round(j['value'],1), {k.split('/')[-1]:round(v['ms_per_frame']*1000,1) for k,v in j['stages'].items()})"
done
