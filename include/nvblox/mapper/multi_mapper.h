// nvblox/mapper/multi_mapper.h -- nvblox::MultiMapper (reference: nvblox/include/nvblox/mapper/multi_mapper.h:26-330,
// mapper/internal/impl/multi_mapper_impl.h), the object nvblox_ros actually drives
// (nvblox_ros/src/lib/nvblox_node.cpp:187-204, :781, :1058-1062, :1261-1264), as a facade over one or two Mappers of
// libnvblox_b200.so:
//   kStaticTsdf / kStaticOccupancy : background mapper only (multi_mapper_impl.h:27-29);
//   kDynamic                       : background = TSDF + freespace layer, updated after every depth frame
//                                    (multi_mapper_impl.h:115-119); foreground = occupancy of the depth pixels given by a
//                                    caller-supplied dynamic mask (setDynamicMask): the reference derives that mask with its
//                                    DynamicsDetection + connected-component filter (multi_mapper_impl.h:72-105), which are outside
//                                    this path;
//   kHumanWithStatic*              : the mask overload splits the depth frame: masked pixels -> foreground occupancy, the
//                                    others -> background (multi_mapper_impl.h:127-170). Only T_CM_CD = identity with one sensor
//                                    (mask in the depth camera's frame) is built: the reference's ImageMasker re-projection is
//                                    outside this path.
// updateEsdf(): 3-D or 2-D (slice) ESDF of the mappers in use (multi_mapper.cpp updateEsdfOfMapper).
#pragma once
#include <memory>
#include <optional>
#include "nvblox/mapper/mapper.h"
namespace nvblox {

enum class MappingType { kStaticTsdf, kStaticOccupancy, kDynamic, kHumanWithStaticTsdf, kHumanWithStaticOccupancy };
inline bool isHumanMapping(MappingType t) { return t == MappingType::kHumanWithStaticTsdf || t == MappingType::kHumanWithStaticOccupancy; }
inline bool isDynamicMapping(MappingType t) { return t == MappingType::kDynamic; }
inline bool isStaticOccupancy(MappingType t) { return t == MappingType::kStaticOccupancy || t == MappingType::kHumanWithStaticOccupancy; }
inline bool isUsingBothMappers(MappingType t) { return isHumanMapping(t) || isDynamicMapping(t); }

struct MultiMapperParams {
  int connected_mask_component_size_threshold = 2000;  // (accepted; the connected-component filter is outside this path)
  bool remove_small_connected_components = true;
};

class MultiMapper {
 public:
  MultiMapper(float voxel_size_m, MappingType mapping_type, EsdfMode esdf_mode, MemoryType memory_type = MemoryType::kDevice,
              std::shared_ptr<CudaStream> cuda_stream = std::make_shared<CudaStreamOwning>())
      : mapping_type_(mapping_type), esdf_mode_(esdf_mode) {
    // multi_mapper.cpp: background layer type by mapping type; the foreground mapper is always occupancy
    const ProjectiveLayerType bg = isStaticOccupancy(mapping_type)   ? ProjectiveLayerType::kOccupancy
                                   : isDynamicMapping(mapping_type) ? ProjectiveLayerType::kTsdfWithFreespace
                                                                    : ProjectiveLayerType::kTsdf;
    background_mapper_ = std::make_shared<Mapper>(voxel_size_m, BlockMemoryPoolParams(memory_type), bg, cuda_stream);
    if (isUsingBothMappers(mapping_type))
      foreground_mapper_ = std::make_shared<Mapper>(voxel_size_m, BlockMemoryPoolParams(memory_type), ProjectiveLayerType::kOccupancy, cuda_stream);
  }
  void setMultiMapperParams(const MultiMapperParams& p) { params_ = p; }
  const MultiMapperParams& getMultiMapperParams() const { return params_; }
  void setMapperParams(const MapperParams& background_mapper_params,
                       const std::optional<MapperParams>& foreground_mapper_params = std::nullopt) {
    background_mapper_->setMapperParams(background_mapper_params);
    if (foreground_mapper_ && foreground_mapper_params.has_value()) foreground_mapper_->setMapperParams(*foreground_mapper_params);
  }

  // integrateDepth(depth_frame, T_L_CD, depth_sensor, update_time_ms) -- multi_mapper.h:146-151
  template <typename SensorType>
  void integrateDepth(const DepthImage& depth_frame, const Transform& T_L_CD, const SensorType& depth_sensor,
                      const std::optional<Time>& update_time_ms = std::nullopt) {
    background_mapper_->integrateDepth(depth_frame, T_L_CD, depth_sensor);
    if (!isDynamicMapping(mapping_type_)) return;
    if (!update_time_ms.has_value()) b200_detail::check(NVB_ERR_INVALID_ARGUMENT, "MultiMapper::integrateDepth", "dynamic mapping needs update_time_ms");
    if (dynamic_mask_.has_value())
      foreground_mapper_->integrateDepth(MaskedDepthImageConstView(depth_frame, *dynamic_mask_), T_L_CD, depth_sensor);
    background_mapper_->updateFreespace(*update_time_ms, T_L_CD, depth_sensor, DepthImageConstView(depth_frame));
  }
  // The dynamic mask of the NEXT depth frame (kDynamic): >0 = dynamic pixel. Stands in for DynamicsDetection::computeDynamics.
  void setDynamicMask(const MonoImageConstView& mask) { dynamic_mask_ = mask; }

  // integrateDepth(depth_frame, mask, T_L_CD, T_CM_CD, depth_sensor, mask_sensor) -- multi_mapper.h:205-209 (human mapping)
  template <typename SensorType>
  void integrateDepth(const DepthImage& depth_frame, const MonoImage& mask, const Transform& T_L_CD, const Transform& T_CM_CD,
                      const SensorType& depth_sensor, const SensorType& /*mask_sensor*/) {
    if (!isHumanMapping(mapping_type_)) b200_detail::check(NVB_ERR_INVALID_ARGUMENT, "MultiMapper::integrateDepth", "a mask is only valid for human mapping");
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 4; c++)
        if (std::fabs(T_CM_CD(r, c) - (r == c ? 1.0f : 0.0f)) > 1e-6f)
          b200_detail::check(NVB_ERR_INVALID_ARGUMENT, "MultiMapper::integrateDepth", "only T_CM_CD = identity is built (no ImageMasker re-projection)");
    // foreground: the masked pixels; background: the others (MaskMode::kInverted)
    foreground_mapper_->integrateDepth(MaskedDepthImageConstView(depth_frame, MonoImageConstView(mask)), T_L_CD, depth_sensor);
    background_mapper_->integrateDepth(MaskedDepthImageConstView(depth_frame, MonoImageConstView(mask), MaskMode::kInverted), T_L_CD,
                                       depth_sensor);
  }

  // integrateColor(color_frame, T_L_C, sensor) -- multi_mapper.h:237-239: colour only goes to the background mapper
  template <typename SensorType>
  void integrateColor(const ColorImage& color_frame, const Transform& T_L_C, const SensorType& sensor) {
    background_mapper_->integrateColor(color_frame, T_L_C, sensor);
  }
  template <typename SensorType>
  void integrateColor(const ColorImage& color_frame, const MonoImage& mask, const Transform& T_L_C, const SensorType& sensor) {
    // human mapping: the pixels that are NOT on a human colour the static map (multi_mapper.h:241-250)
    background_mapper_->integrateColor(MaskedColorImageConstView(color_frame, MonoImageConstView(mask), MaskMode::kInverted), T_L_C, sensor);
  }

  // updateEsdf() -- multi_mapper.h:266
  void updateEsdf() {
    updateEsdfOfMapper(background_mapper_);
    if (foreground_mapper_) updateEsdfOfMapper(foreground_mapper_);
  }

  const Mapper& background_mapper() const { return *background_mapper_; }
  const Mapper& foreground_mapper() const { return *foreground_mapper_; }
  std::shared_ptr<Mapper> background_mapper() { return background_mapper_; }
  std::shared_ptr<Mapper> foreground_mapper() { return foreground_mapper_; }
  MappingType mapping_type() const { return mapping_type_; }
  EsdfMode esdf_mode() const { return esdf_mode_; }

 protected:
  void updateEsdfOfMapper(const std::shared_ptr<Mapper>& mapper) {
    if (esdf_mode_ == EsdfMode::k2D) mapper->updateEsdfSlice();
    else mapper->updateEsdf();
  }
  const MappingType mapping_type_;
  const EsdfMode esdf_mode_;
  MultiMapperParams params_;
  std::shared_ptr<Mapper> background_mapper_, foreground_mapper_;
  std::optional<MonoImageConstView> dynamic_mask_;
};
}  // namespace nvblox
