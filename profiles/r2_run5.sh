#!/bin/bash
# N GPUs (gpurun --gpus N): merge test on one rank, then the workloads under torchrun
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "merge or union or dropin" 2>&1 | tail -3
for w in c2 c4 c5; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 --workload $w > gpurun_out/r2_run5_n${N}_$w.json 2> gpurun_out/r2_run5_n${N}_$w.err
  echo "bench $w N=$N rc=$?"; tail -3 gpurun_out/r2_run5_n${N}_$w.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_run5_n${N}_$w.json').read().strip().splitlines()[-1])
    print('$w N=$N value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'parity',d.get('parity_checked'),'merge',d.get('merge') and {k:d['merge'][k] for k in ('batches_per_step','ms_per_merge_median_max_over_ranks','last_union_blocks')})
except Exception as e:
    print('no bench line', e)
PY
done
