#!/bin/bash
# A/B of environment switches: each argument is an "ENV=VALUE [ENV=VALUE...]" string applied to one bench run.
for envs in "$@"; do
  echo "=== $envs"
  env $envs python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('frames/s', round(d['value']), 'e2e', round(d['e2e']['value']), {k.split('/')[-1]: round(v['ms_per_frame'],4) for k,v in d['stages'].items()})"
done
