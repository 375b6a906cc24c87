#!/bin/bash
# A/B of wave-kernel build variants on the GPU box: for each set of -D flags, rebuild, print the ESDF time split of
# frame 9 and the bench line's frames/s. Usage: profiles/wave_ab.sh "-DNVB_WAVE_TAIL=0" "-DNVB_WAVE_TAIL=4" ...
for flags in "$@"; do
  NVB_EXTRA_NVCC_FLAGS="$flags" python isaac_ros_nvblox_b200/build_ext.py --force > /dev/null 2>&1
  echo "=== $flags"
  python profiles/esdf_split.py 2>&1 | grep -E "^9 " | sed 's/.*barrier_wait/barrier_wait/' | head -1
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('frames/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'wave ms', round(d['stages']['esdf/integrate/compute']['ms_per_frame'],4))"
done
python isaac_ros_nvblox_b200/build_ext.py --force > /dev/null 2>&1
