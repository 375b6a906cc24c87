#!/bin/bash
# round 2, first GPU call: parity of the exchange-slab wavefront (mode 3) + A/B bench mode 1 vs mode 3
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_run1_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "esdf" > gpurun_out/r2_run1_esdf_tests.log 2>&1
echo "esdf tests rc=$?" >> gpurun_out/r2_run1_esdf_tests.log
tail -5 gpurun_out/r2_run1_esdf_tests.log
timeout 900 python -m pytest tests/test_gpu_bench_pipeline.py -x -q -m gpu > gpurun_out/r2_run1_pipeline_tests.log 2>&1
echo "pipeline tests rc=$?" >> gpurun_out/r2_run1_pipeline_tests.log
tail -5 gpurun_out/r2_run1_pipeline_tests.log
for mode in 1 3; do
  NVB_ESDF_MODE=$mode timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_run1_bench_mode$mode.json 2> gpurun_out/r2_run1_bench_mode$mode.err
  echo "bench mode $mode rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_run1_bench_mode$mode.json').read().strip().splitlines()[-1])
    print('mode $mode value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'parity',d.get('parity_checked'),{k:round(v['ms_per_frame']*1e3,1) for k,v in d['stages'].items()})
except Exception as e:
    print('no bench line', e)
PY
done
