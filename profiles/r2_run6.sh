#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
run() { # label, extra args, env
  out=$(env $3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline $2 2>/dev/null | tail -1)
  python - "$1" "$out" <<'PY'
import json,sys
try:
    d=json.loads(sys.argv[2]); print(sys.argv[1],'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'merge',d.get('merge') and round(d['merge']['ms_per_merge_median_max_over_ranks'],3))
except Exception as e: print(sys.argv[1],'failed',e)
PY
}
one=$(timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1)
python - "$one" <<'PY'
import json,sys
d=json.loads(sys.argv[1]); print('N=1 value',round(d['value'],1),'ms/step',round(d['ms_per_step'],2))
PY
run "N=$N merge 1ch" "" ""
run "N=$N no merge" "--no-merge" ""
run "N=$N merge default channels" "" "NCCL_MAX_NCHANNELS=32 NCCL_MIN_NCHANNELS=2"
