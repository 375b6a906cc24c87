#!/bin/bash
# Builds the wave kernel in several shapes on the GPU box and prints the ESDF time split + frames/s of each.
for v in "256 1 128 0" "256 1 128 1" "512 1 128 1"; do
  set -- $v
  NVB_EXTRA_NVCC_FLAGS="-DNVB_WAVE_THREADS=$1 -DNVB_WAVE_INLINE=$2 -DNVB_WAVE_MAXREG=$3 -DNVB_WAVE_PAD=$4" python isaac_ros_nvblox_b200/build_ext.py --force > /dev/null 2>&1
  echo "=== threads=$1 inline=$2 maxreg=$3 pad=$4"
  python profiles/esdf_split.py 2>&1 | grep -E "^9 " | sed 's/.*barrier_wait/barrier_wait/' | head -1
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('frames/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'wave ms', round(d['stages']['esdf/integrate/compute']['ms_per_frame'],3))"
done
python isaac_ros_nvblox_b200/build_ext.py --force > /dev/null 2>&1
