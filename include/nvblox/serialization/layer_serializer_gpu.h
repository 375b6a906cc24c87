// nvblox/serialization/layer_serializer_gpu.h, mesh_serializer_gpu.h -- LayerSerializerGpu<LayerType>, MeshSerializerGpu<Color>
// (reference: nvblox/include/nvblox/serialization/layer_serializer_gpu.h:32-95, mesh_serializer_gpu.h:30-139). The requested blocks
// are packed back to back by one gather kernel and one device-to-host copy per array in libnvblox_b200.so (nvb_layer_get_blocks,
// nvb_mesh_get_blocks); the results live in plain std::vectors (the reference's host_vector).
#pragma once
#include <memory>
#include <vector>
#include "nvblox/core/cuda_stream.h"
#include "nvblox/map/layer.h"
#include "nvblox/mesh/mesh_block.h"
namespace nvblox {
template <typename VoxelType>
struct SerializedLayer {
  std::vector<Index3D> block_indices;
  std::vector<VoxelType> voxels;
  std::vector<int32_t> block_offsets;  // size num_blocks + 1, in voxels; a block that is not in the layer has size 0
};
using SerializedTsdfLayer = SerializedLayer<TsdfVoxel>;
using SerializedColorLayer = SerializedLayer<ColorVoxel>;
using SerializedOccupancyLayer = SerializedLayer<OccupancyVoxel>;
using SerializedFreespaceLayer = SerializedLayer<FreespaceVoxel>;
using SerializedEsdfLayer = SerializedLayer<EsdfVoxel>;

template <class LayerType>
class LayerSerializerGpu {
 public:
  using BlockType = typename LayerType::BlockType;
  using VoxelType = typename BlockType::VoxelType;
  using SerializedLayerType = SerializedLayer<VoxelType>;
  LayerSerializerGpu() : serialized_layer_(std::make_shared<SerializedLayerType>()) {}
  std::shared_ptr<SerializedLayerType> serialize(const LayerType& layer, const std::vector<Index3D>& block_indices_to_serialize,
                                                 const CudaStream& cuda_stream = CudaStreamOwning()) {
    (void)cuda_stream;
    constexpr int kVoxelsPerBlock = 512;
    const size_t n = block_indices_to_serialize.size();
    SerializedLayerType& out = *serialized_layer_;
    out.block_indices = block_indices_to_serialize;
    out.block_offsets.assign(n + 1, 0);
    out.voxels.clear();
    if (n == 0) return serialized_layer_;
    std::vector<int32_t> raw(n * 3);
    for (size_t i = 0; i < n; i++)
      raw[3 * i] = block_indices_to_serialize[i][0], raw[3 * i + 1] = block_indices_to_serialize[i][1], raw[3 * i + 2] = block_indices_to_serialize[i][2];
    std::vector<BlockType> blocks(n);
    std::vector<uint8_t> found(n, 0);
    b200_detail::check(nvb_layer_get_blocks(layer.mapper_handle(), layer.layer_id(), raw.data(), (int32_t)n, blocks.data(), found.data()),
                       "LayerSerializerGpu::serialize", nvb_last_error());
    for (size_t i = 0; i < n; i++) {
      if (found[i]) {
        const VoxelType* v = &blocks[i].voxels[0][0][0];
        out.voxels.insert(out.voxels.end(), v, v + kVoxelsPerBlock);
      }
      out.block_offsets[i + 1] = (int32_t)out.voxels.size();
    }
    return serialized_layer_;
  }
  std::shared_ptr<SerializedLayerType> getSerializedLayer() { return serialized_layer_; }
 private:
  std::shared_ptr<SerializedLayerType> serialized_layer_;
};
using TsdfLayerSerializerGpu = LayerSerializerGpu<TsdfLayer>;
using ColorLayerSerializerGpu = LayerSerializerGpu<ColorLayer>;
using OccupancyLayerSerializerGpu = LayerSerializerGpu<OccupancyLayer>;
using FreespaceLayerSerializerGpu = LayerSerializerGpu<FreespaceLayer>;
using EsdfLayerSerializerGpu = LayerSerializerGpu<EsdfLayer>;

// ---- mesh (mesh_serializer_gpu.h)
template <typename AppearanceType>
struct SerializedMeshLayer {
  std::vector<Vector3f> vertices;
  std::vector<AppearanceType> vertex_appearances;
  std::vector<int> triangle_indices;
  std::vector<int32_t> vertex_block_offsets;          // size num_blocks + 1; also the appearances' offsets
  std::vector<int32_t> triangle_index_block_offsets;  // size num_blocks + 1
  std::vector<Index3D> block_indices;
  std::vector<int>::const_iterator triangleBlockItr(size_t block_index) const {
    return triangle_indices.begin() + triangle_index_block_offsets[block_index];
  }
  const Vector3f& getVertex(size_t block_index, size_t vertex_index) const { return vertices[vertex_block_offsets[block_index] + vertex_index]; }
  const AppearanceType& getAppearance(size_t block_index, size_t vertex_index) const {
    return vertex_appearances[vertex_block_offsets[block_index] + vertex_index];
  }
  const int& getTriangleIndex(size_t block_index, size_t triangle_index) const {
    return triangle_indices[triangle_index_block_offsets[block_index] + triangle_index];
  }
  size_t getNumVerticesInBlock(size_t block_index) const { return vertex_block_offsets[block_index + 1] - vertex_block_offsets[block_index]; }
  size_t getNumTriangleIndicesInBlock(size_t block_index) const {
    return triangle_index_block_offsets[block_index + 1] - triangle_index_block_offsets[block_index];
  }
};
using SerializedColorMeshLayer = SerializedMeshLayer<Color>;

template <typename AppearanceType>
class MeshSerializerGpu {
 public:
  using SerializedLayerType = SerializedMeshLayer<AppearanceType>;
  using MeshLayerType = MeshBlockLayer<AppearanceType>;
  MeshSerializerGpu() : serialized_mesh_(std::make_shared<SerializedLayerType>()) {}
  std::shared_ptr<SerializedLayerType> serialize(const MeshLayerType& mesh_layer, const std::vector<Index3D>& block_indices_to_serialize,
                                                 const CudaStream& cuda_stream = CudaStreamOwning()) {
    (void)cuda_stream;
    const size_t n = block_indices_to_serialize.size();
    SerializedLayerType& out = *serialized_mesh_;
    out.block_indices = block_indices_to_serialize;
    out.vertex_block_offsets.assign(n + 1, 0), out.triangle_index_block_offsets.assign(n + 1, 0);
    out.vertices.clear(), out.vertex_appearances.clear(), out.triangle_indices.clear();
    if (n == 0) return serialized_mesh_;
    std::vector<int32_t> raw(n * 3), sz(n * 3, -1);
    for (size_t i = 0; i < n; i++)
      raw[3 * i] = block_indices_to_serialize[i][0], raw[3 * i + 1] = block_indices_to_serialize[i][1], raw[3 * i + 2] = block_indices_to_serialize[i][2];
    NvbMapper* m = mesh_layer.mapper_handle();
    b200_detail::check(nvb_mesh_block_sizes(m, raw.data(), (int32_t)n, sz.data()), "MeshSerializerGpu::serialize", nvb_last_error());
    int64_t tv = 0, tt = 0, tc = 0;
    for (size_t i = 0; i < n; i++) {
      if (sz[3 * i] >= 0) tv += sz[3 * i], tt += sz[3 * i + 1], tc += sz[3 * i + 2];
      out.vertex_block_offsets[i + 1] = (int32_t)tv, out.triangle_index_block_offsets[i + 1] = (int32_t)tt;
    }
    std::vector<float> v((size_t)tv * 3 + 3);
    std::vector<uint8_t> c((size_t)tc * 4 + 4);
    out.triangle_indices.resize((size_t)tt);
    const int64_t caps[3] = {tv, tt, tc};
    b200_detail::check(nvb_mesh_get_blocks(m, raw.data(), (int32_t)n, v.data(), nullptr, out.triangle_indices.data(), c.data(), caps),
                       "MeshSerializerGpu::serialize", nvb_last_error());
    out.vertices.resize((size_t)tv), out.vertex_appearances.resize((size_t)tc);
    for (int64_t i = 0; i < tv; i++) out.vertices[i] = Vector3f(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
    for (int64_t i = 0; i < tc; i++) out.vertex_appearances[i] = b200_detail::appearanceFromRgba<AppearanceType>(&c[4 * (size_t)i]);
    return serialized_mesh_;
  }
  std::shared_ptr<SerializedLayerType> getSerializedLayer() const { return serialized_mesh_; }
 private:
  std::shared_ptr<SerializedLayerType> serialized_mesh_;
};
using ColorMeshSerializerGpu = MeshSerializerGpu<Color>;
}  // namespace nvblox
