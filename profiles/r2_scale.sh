#!/bin/bash
# N GPUs (gpurun --gpus N): the bench workloads under torchrun, merge on (BatchMerger reserves 4 SMs for NCCL)
N=${1:-2}; shift
mkdir -p gpurun_out
for w in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/bench_r2_n${N}_$w.json 2> gpurun_out/bench_r2_n${N}_$w.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_r2_n${N}_$w.json').read().strip().splitlines()[-1])
    print('$w N=$N value',round(d['value'],1),'per gpu',round(d['value']/$N,1),'e2e',round(d['e2e']['value'],1),'merge',d.get('merge') and {k:d['merge'][k] for k in ('batches_per_step','ms_per_merge_median_max_over_ranks','last_union_blocks')}, d['clocks'])
except Exception as e:
    print('no bench line', e); print(open('gpurun_out/bench_r2_n${N}_$w.err').read()[-800:])
PY
done
