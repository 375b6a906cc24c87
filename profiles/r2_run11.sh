#!/bin/bash
python -m pytest tests/test_gpu_mesh.py -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --voxel-size 0.02 --frames 16 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('2cm value', round(j['value'],1), 'mesh', json.dumps(j['with_mesh'])[:400])"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('5cm value', round(j['value'],1), 'mesh', json.dumps(j['with_mesh'])[:400], {k.split('/')[-1]:round(v['ms_per_frame']*1000,1) for k,v in j['stages'].items()})"
