#!/bin/bash
# merge kernels + MultiMapper drop-in + order KATs + 3dmatch on the GPU, then the three bench workloads at N=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_pipeline.py -x -q -m gpu -k "dropin or merge or union or hand_derived or threedmatch" > gpurun_out/r2_run4_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_run4_tests.log
tail -4 gpurun_out/r2_run4_tests.log
for w in c5; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/r2_run4_bench_$w.json 2> gpurun_out/r2_run4_bench_$w.err
  echo "bench $w rc=$?"; tail -2 gpurun_out/r2_run4_bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_run4_bench_$w.json').read().strip().splitlines()[-1])
    print('$w value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'parity',d.get('parity_checked'),'cpu',round(d['cpu_baseline']['value'],1),{k.split('/')[-1]:round(v['ms_per_frame']*1e3,1) for k,v in d['stages'].items()})
except Exception as e:
    print('no bench line', e)
PY
done
