#!/bin/bash
# Builds the wave kernel in several shapes on the GPU box and prints the ESDF time split of each.
for v in "512 0" "512 1" "256 1" "1024 0"; do
  set -- $v
  NVB_EXTRA_NVCC_FLAGS="-DNVB_WAVE_THREADS=$1 -DNVB_WAVE_INLINE=$2" python isaac_ros_nvblox_b200/build_ext.py --force > /dev/null 2>&1
  echo "=== threads=$1 inline=$2"
  python profiles/esdf_split.py 2>&1 | grep -E "^9 |^8 " | sed 's/.*barrier_wait/barrier_wait/' | head -2
  python profiles/esdf_split.py 2>&1 | grep "per-phase" | tail -1 | cut -c1-400
done
python isaac_ros_nvblox_b200/build_ext.py --force > /dev/null 2>&1
