#!/bin/bash
# 8 GPUs, c2: (a) no merge at all -> what the ranks' different trajectories cost in a max-over-ranks number;
# (b) merge with NCCL limited to 2 channels (its CTAs then fit the 4 SMs the wavefront leaves free)
N=8
run() { tag=$1; shift
  timeout 600 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/bench_r2_n8_$tag.json 2> gpurun_out/bench_r2_n8_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_r2_n8_$tag.json').read().strip().splitlines()[-1])
    print('$tag value',round(d['value'],1),'per gpu',round(d['value']/$N,1),'merge',d.get('merge') and d['merge'].get('ms_per_merge_median_max_over_ranks'))
except Exception as e:
    print('$tag no bench line', e)
PY
}
EXTRA="--no-merge" run nomerge FOO=1
EXTRA="" run nch2 NCCL_MAX_NCHANNELS=2
