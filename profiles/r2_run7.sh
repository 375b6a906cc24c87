# round 2, run 7: C++ drop-in programs (incl. the mesh one) + the default bench with the mesh leg
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cpp_dropin" 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_run7_bench.json 2> gpurun_out/r2_run7_bench.err
tail -3 gpurun_out/r2_run7_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_run7_bench.json'):
    if l.startswith('{'):
        j = json.loads(l)
        print(j['value'], j['e2e']['value'], j['parity_checked'])
        print(json.dumps(j['with_mesh'])[:900])
        print(json.dumps(j['with_color'])[:300])
PY
