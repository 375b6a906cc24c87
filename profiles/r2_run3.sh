#!/bin/bash
# quick loop: mode-3 parity subset + bench + per-ring split of a profiling build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_occupancy.py tests/test_gpu_bench_pipeline.py tests/test_gpu_decay.py tests/test_gpu_esdf_slice.py -x -q -m gpu -k "(esdf and (3 or exchange or pruning)) or (bench_pipeline_async_80 and 3) or growth_mid or decay or slice" > gpurun_out/r2_run3_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_run3_tests.log
tail -4 gpurun_out/r2_run3_tests.log
NVB_ESDF_MODE=${MODE:-3} timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_run3_bench.json 2> gpurun_out/r2_run3_bench.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_run3_bench.json').read().strip().splitlines()[-1])
    print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'parity',d.get('parity_checked'),{k:round(v['ms_per_frame']*1e3,1) for k,v in d['stages'].items()})
except Exception as e:
    print('no bench line', e)
PY
NVB_EXTRA_NVCC_FLAGS="-DNVB_WAVEX_PROF=1" python isaac_ros_nvblox_b200/build_ext.py --force > /dev/null 2>&1
timeout 300 python profiles/wavex_split.py ${MODE:-3} 30 2 > gpurun_out/r2_run3_wavex_split.log 2>&1
tail -8 gpurun_out/r2_run3_wavex_split.log
