// nvblox/mesh/mesh_block.h -- MeshBlock<AppearanceType> and the mesh layer (reference: nvblox/include/nvblox/mesh/mesh_block.h:32-83,
// nvblox/include/nvblox/map/common_names.h). In libnvblox_b200.so the mesh lives in one device arena; a MeshBlock here is the
// host copy of one block's four vectors, a ColorMeshLayer a view of the mapper's mesh layer.
#pragma once
#include <memory>
#include <vector>
#include "nvblox/core/types.h"
#include "nvblox/map/voxels.h"
#include "nvblox_b200.h"
namespace nvblox {
template <typename AppearanceType>
struct MeshBlock {
  typedef std::shared_ptr<MeshBlock> Ptr;
  typedef std::shared_ptr<const MeshBlock> ConstPtr;
  std::vector<Vector3f> vertices;
  std::vector<Vector3f> vertex_normals;
  std::vector<AppearanceType> vertex_appearances;
  std::vector<int> triangles;  // indices into `vertices`, three per triangle
  void clear() { vertices.clear(), vertex_normals.clear(), vertex_appearances.clear(), triangles.clear(); }
  size_t size() const { return vertices.size(); }
  size_t capacity() const { return vertices.capacity(); }
};
using ColorMeshBlock = MeshBlock<Color>;

namespace b200_detail {
template <typename A>
inline A appearanceFromRgba(const uint8_t* p);
template <>
inline Color appearanceFromRgba<Color>(const uint8_t* p) {
  Color c;
  c.r = p[0], c.g = p[1], c.b = p[2];
  return c;
}
}  // namespace b200_detail

template <typename AppearanceType>
class MeshBlockLayer {
 public:
  using BlockType = MeshBlock<AppearanceType>;
  explicit MeshBlockLayer(NvbMapper* m) : m_(m) {}
  NvbMapper* mapper_handle() const { return m_; }
  float block_size() const { return nvb_mapper_block_size(m_); }
  int numBlocks() const {
    int32_t n = 0;
    // before the first mesh update the layer does not exist yet: empty
    return nvb_layer_num_blocks(m_, NVB_LAYER_MESH, &n) == NVB_OK ? n : 0;
  }
  size_t size() const { return (size_t)numBlocks(); }
  size_t numAllocatedBlocks() const { return size(); }
  std::vector<Index3D> getAllBlockIndices() const {
    const int n = numBlocks();
    std::vector<Index3D> out((size_t)n);
    if (n == 0) return out;
    std::vector<int32_t> raw((size_t)n * 3 + 3);
    int32_t cnt = 0;
    b200_detail::check(nvb_layer_block_indices(m_, NVB_LAYER_MESH, raw.data(), n, &cnt), "getAllBlockIndices", nvb_last_error());
    for (int i = 0; i < n; i++) out[i] = Index3D(raw[3 * i], raw[3 * i + 1], raw[3 * i + 2]);
    return out;
  }
  bool isBlockAllocated(const Index3D& idx) const {
    const int32_t k[3] = {idx[0], idx[1], idx[2]};
    int32_t sz[3] = {-1, -1, -1};
    b200_detail::check(nvb_mesh_block_sizes(m_, k, 1, sz), "isBlockAllocated", nvb_last_error());
    return sz[0] >= 0;
  }
  // BlockLayer::getBlockAtIndex: null if the block has no mesh block; otherwise a host copy of its vectors
  typename BlockType::ConstPtr getBlockAtIndex(const Index3D& idx) const {
    const int32_t k[3] = {idx[0], idx[1], idx[2]};
    int32_t sz[3] = {-1, -1, -1};
    b200_detail::check(nvb_mesh_block_sizes(m_, k, 1, sz), "getBlockAtIndex", nvb_last_error());
    if (sz[0] < 0) return nullptr;
    auto b = std::make_shared<BlockType>();
    std::vector<float> v((size_t)sz[0] * 3 + 3), n((size_t)sz[0] * 3 + 3);
    std::vector<uint8_t> c((size_t)sz[2] * 4 + 4);
    b->triangles.resize((size_t)sz[1]);
    const int64_t caps[3] = {sz[0], sz[1], sz[2]};
    b200_detail::check(nvb_mesh_get_blocks(m_, k, 1, v.data(), n.data(), b->triangles.data(), c.data(), caps), "getBlockAtIndex",
                       nvb_last_error());
    b->vertices.resize((size_t)sz[0]), b->vertex_normals.resize((size_t)sz[0]), b->vertex_appearances.resize((size_t)sz[2]);
    for (int i = 0; i < sz[0]; i++) {
      b->vertices[i] = Vector3f(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
      b->vertex_normals[i] = Vector3f(n[3 * i], n[3 * i + 1], n[3 * i + 2]);
    }
    for (int i = 0; i < sz[2]; i++) b->vertex_appearances[i] = b200_detail::appearanceFromRgba<AppearanceType>(&c[4 * (size_t)i]);
    return b;
  }
 private:
  NvbMapper* m_;
};
using ColorMeshLayer = MeshBlockLayer<Color>;
}  // namespace nvblox
