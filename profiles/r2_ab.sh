#!/bin/bash
# A/B of build flags / environment switches on the bench (device-resident value only): each line "label|nvcc flags|env"
mkdir -p gpurun_out
while IFS='|' read -r label flags envs; do
  [ -z "$label" ] && continue
  NVB_EXTRA_NVCC_FLAGS="$flags" python isaac_ros_nvblox_b200/build_ext.py --force > /dev/null 2>&1
  out=$(env $envs NVB_ESDF_MODE=3 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1)
  python - "$label" "$out" <<'PY'
import json,sys
try:
    d=json.loads(sys.argv[2])
    print(sys.argv[1], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k.split('/')[-1]:round(v['ms_per_frame']*1e3,1) for k,v in d['stages'].items()})
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
done <<'CFG'
prune on||
prune off||NVB_CLEAR_PRUNE=0
prune on (2)||
prune off (2)||NVB_CLEAR_PRUNE=0
CFG
