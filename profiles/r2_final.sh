#!/bin/bash
# Round-2 measurement set on one B200 (profiles/README.md).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; tail -2 gpurun_out/bench_r2_n1.err
for w in c4 c5; do python bench.py --steps 3 --warmup 3 --workload $w > gpurun_out/bench_r2_n1_$w.json 2> gpurun_out/bench_r2_n1_$w.err; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err
python bench.py --steps 3 --warmup 3 --voxel-size 0.02 --frames 16 --no-cpu-baseline > gpurun_out/bench_r2_n1_2cm.json 2> gpurun_out/bench_r2_n1_2cm.err
python - <<'PY'
import json
for f in ("bench_r2_n1", "bench_r2_n1_c4", "bench_r2_n1_c5", "bench_r2_reference", "bench_r2_n1_2cm"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "parity", d.get("parity_checked"), "cpu", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"], 1),
              "mesh", d.get("with_mesh") and (round(d["with_mesh"]["value"], 1), d["with_mesh"].get("parity_checked")), "color", d.get("with_color") and round(d["with_color"]["value"], 1))
    except Exception as e:
        print(f, "no line", e)
PY
