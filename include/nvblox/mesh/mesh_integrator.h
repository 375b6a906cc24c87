// nvblox/mesh/mesh_integrator.h -- MeshIntegrator<ColorVoxel> (reference: nvblox/include/nvblox/mesh/mesh_integrator.h:39-162).
// Marching cubes + vertex welding + vertex colours run in libnvblox_b200.so (isaac_ros_nvblox_b200/csrc/nvb_mesh.cu); the layers
// passed in are views of one mapper's device map, so they only identify it.
#pragma once
#include <vector>
#include "nvblox/map/layer.h"
#include "nvblox/mesh/mesh_block.h"
namespace nvblox {
enum class DeviceType { kCPU, kGPU };
struct MeshIntegratorParams {
  float mesh_integrator_min_weight = 1e-4f;   // mesh/mesh_integrator_params.h:22-24
  bool mesh_integrator_weld_vertices = true;  // mesh/mesh_integrator_params.h:25-27
};
class ColorMeshIntegrator {
 public:
  explicit ColorMeshIntegrator(NvbMapper* m) : m_(m) {}
  float min_weight() const { return get().min_weight; }
  void min_weight(float v) { NvbMeshParams p = get(); p.min_weight = v; set(p); }
  bool weld_vertices() const { return get().weld_vertices != 0; }
  void weld_vertices(bool v) { NvbMeshParams p = get(); p.weld_vertices = v ? 1 : 0; set(p); }
  // integrateBlocksGPU (mesh_integrator.h:66-70): re-meshes the listed blocks (those present in the TSDF layer)
  bool integrateBlocksGPU(const TsdfLayer& distance_layer, const std::vector<Index3D>& block_indices, ColorMeshLayer* mesh_layer) {
    (void)distance_layer, (void)mesh_layer;
    std::vector<int32_t> raw(block_indices.size() * 3 + 3);
    for (size_t i = 0; i < block_indices.size(); i++)
      raw[3 * i] = block_indices[i][0], raw[3 * i + 1] = block_indices[i][1], raw[3 * i + 2] = block_indices[i][2];
    b200_detail::check(nvb_mesh_integrate_blocks(m_, raw.data(), (int32_t)block_indices.size(), 0), "integrateBlocksGPU", nvb_last_error());
    return true;
  }
  // integrateMeshFromDistanceField (mesh_integrator.h:58-61). The kCPU path of the reference differs from its GPU path only in
  // rounding (it is not what Mapper::updateColorMesh runs); both device types run the GPU path here.
  bool integrateMeshFromDistanceField(const TsdfLayer& distance_layer, ColorMeshLayer* mesh_layer,
                                      const DeviceType device_type = DeviceType::kGPU) {
    (void)device_type;
    return integrateBlocksGPU(distance_layer, distance_layer.getAllBlockIndices(), mesh_layer);
  }
  // updateAppearance (mesh_integrator.h:95-112): vertex colours from the colour layer (gray where it has no block)
  void updateAppearance(const ColorLayer& color_layer, const std::vector<Index3D>& block_indices, ColorMeshLayer* mesh_layer) {
    (void)color_layer, (void)mesh_layer;
    std::vector<int32_t> raw(block_indices.size() * 3 + 3);
    for (size_t i = 0; i < block_indices.size(); i++)
      raw[3 * i] = block_indices[i][0], raw[3 * i + 1] = block_indices[i][1], raw[3 * i + 2] = block_indices[i][2];
    b200_detail::check(nvb_mesh_update_color(m_, raw.data(), (int32_t)block_indices.size()), "updateAppearance", nvb_last_error());
  }
  void updateAppearance(const ColorLayer& color_layer, ColorMeshLayer* mesh_layer) {
    updateAppearance(color_layer, mesh_layer->getAllBlockIndices(), mesh_layer);
  }
 private:
  NvbMeshParams get() const {
    NvbMeshParams p;
    b200_detail::check(nvb_mapper_get_mesh_params(m_, &p), "mesh params", nvb_last_error());
    return p;
  }
  void set(const NvbMeshParams& p) { b200_detail::check(nvb_mapper_set_mesh_params(m_, &p), "mesh params", nvb_last_error()); }
  NvbMapper* m_;
};
}  // namespace nvblox
