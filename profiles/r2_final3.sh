#!/bin/bash
# after the one-sync read-back of the synchronous API: full GPU suite + the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_final3_n1.json 2> gpurun_out/bench_r2_final3_n1.err; tail -2 gpurun_out/bench_r2_final3_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2_final3_n1.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "async", round(d["e2e"]["async_api"]["value"], 1), "parity", d.get("parity_checked"), "mesh", d["with_mesh"]["value"], d["with_mesh"].get("parity_checked"), "color", round(d["with_color"]["value"], 1))
PY
