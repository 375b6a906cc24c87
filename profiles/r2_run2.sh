#!/bin/bash
# round 2: parity of the exchange-slab wavefront (mode 3) + bench A/B + per-ring split of a profiling build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_occupancy.py -x -q -m gpu -k "esdf" > gpurun_out/r2_run2_esdf_tests.log 2>&1
echo "esdf tests rc=$?" >> gpurun_out/r2_run2_esdf_tests.log
tail -4 gpurun_out/r2_run2_esdf_tests.log
timeout 900 python -m pytest tests/test_gpu_bench_pipeline.py -x -q -m gpu > gpurun_out/r2_run2_pipeline_tests.log 2>&1
echo "pipeline tests rc=$?" >> gpurun_out/r2_run2_pipeline_tests.log
tail -4 gpurun_out/r2_run2_pipeline_tests.log
for mode in 3; do
  NVB_ESDF_MODE=$mode timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_run2_bench_mode$mode.json 2> gpurun_out/r2_run2_bench_mode$mode.err
  echo "bench mode $mode rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_run2_bench_mode$mode.json').read().strip().splitlines()[-1])
    print('mode $mode value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'parity',d.get('parity_checked'),{k:round(v['ms_per_frame']*1e3,1) for k,v in d['stages'].items()})
except Exception as e:
    print('no bench line', e)
PY
done
# profiling build (cycle counters in the wavefront kernel)
NVB_EXTRA_NVCC_FLAGS="-DNVB_WAVEX_PROF=1" python isaac_ros_nvblox_b200/build_ext.py --force > /dev/null 2>&1
timeout 300 python profiles/wavex_split.py 3 30 2 > gpurun_out/r2_run2_wavex_split.log 2>&1
tail -12 gpurun_out/r2_run2_wavex_split.log
