// nvblox/map/blox.h -- VoxelBlock<V> (reference: nvblox/include/nvblox/map/blox.h:28-67).
#pragma once
#include "nvblox/core/unified_ptr.h"
#include "nvblox/map/voxels.h"
namespace nvblox {
template <typename _VoxelType>
struct VoxelBlock {
  typedef _VoxelType VoxelType;
  static constexpr int kVoxelsPerSide = 8;
  static constexpr int kNumVoxels = 512;
  typedef unified_ptr<VoxelBlock> Ptr;             // map/blox.h:45
  typedef unified_ptr<const VoxelBlock> ConstPtr;  // map/blox.h:46
  VoxelType voxels[kVoxelsPerSide][kVoxelsPerSide][kVoxelsPerSide];
};
using TsdfBlock = VoxelBlock<TsdfVoxel>;
using EsdfBlock = VoxelBlock<EsdfVoxel>;
static_assert(sizeof(TsdfBlock) == 4096 && sizeof(EsdfBlock) == 10240, "block layout");
}  // namespace nvblox
