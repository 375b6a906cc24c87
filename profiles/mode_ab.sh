#!/bin/bash
# A/B of the ESDF wavefront implementations on the GPU box (NVB_ESDF_MODE: 1 four-phase, 2 gather-emulate-sweep).
for mode in "$@"; do
  echo "=== NVB_ESDF_MODE=$mode"
  NVB_ESDF_MODE=$mode python profiles/esdf_split.py 2>&1 | grep -E "^(5|9) " | sed 's/.*barrier_wait/barrier_wait/'
  NVB_ESDF_MODE=$mode python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('frames/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'wave ms', round(d['stages']['esdf/integrate/compute']['ms_per_frame'],4), {k: round(v['ms_per_frame'],4) for k,v in d['stages'].items()})"
done
