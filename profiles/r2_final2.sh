#!/bin/bash
# Round-2 closing run on one B200: full GPU suite, smoke, default bench + reference arm, ncu launch list + full raw page.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_final_n1.json 2> gpurun_out/bench_r2_final_n1.err; tail -2 gpurun_out/bench_r2_final_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_final_reference.json 2> gpurun_out/bench_r2_final_reference.err
bash profiles/r2_ncu.sh 2>&1 | tail -3
python - <<'PY'
import json
for f in ("bench_r2_final_n1", "bench_r2_final_reference"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "parity", d.get("parity_checked"), "mesh", d.get("with_mesh") and (round(d["with_mesh"]["value"], 1), d["with_mesh"].get("parity_checked")),
              "color", d.get("with_color") and round(d["with_color"]["value"], 1), "traffic", d.get("roofline") and d["roofline"].get("traffic"))
    except Exception as e:
        print(f, "no line", e)
PY
