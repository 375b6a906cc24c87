#!/bin/bash
# 2 GPUs: does leaving SMs to the NCCL all-gather lift the merge stall? (c2, merge on)
N=2
for r in 0 4 8; do
  NVB_WAVEX_RESERVED_SMS=$r timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_run8_r$r.json 2> gpurun_out/r2_run8_r$r.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_run8_r$r.json').read().strip().splitlines()[-1])
print('reserved $r value',round(d['value'],1),'merge',d.get('merge') and {k:d['merge'][k] for k in ('batches_per_step','ms_per_merge_median_max_over_ranks')})
PY
done
NVB_WAVEX_RESERVED_SMS=4 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-merge 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): print('no-merge reserved 4', round(json.loads(l)['value'],1))"
