#!/bin/bash
# A/B of the hybrid wavefront's switch point (NVB_GES_SWITCH = largest ring, in members, that runs as gather-replay).
for sw in "$@"; do
  echo "=== NVB_ESDF_MODE=2 NVB_GES_SWITCH=$sw"
  NVB_ESDF_MODE=2 NVB_GES_SWITCH=$sw python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('frames/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'wave ms', round(d['stages']['esdf/integrate/compute']['ms_per_frame'],4))"
done
