// nvblox/integrators/esdf_slicer.h -- EsdfSlicer subset (reference: nvblox/include/nvblox/integrators/esdf_slicer.h:30-118):
// the distance-map image nvblox_ros publishes for Nav2, computed on the device by nvb_esdf_slice_distance_image.
#pragma once
#include <cstdint>
#include <vector>
#include "nvblox/geometry/plane.h"
#include "nvblox/map/layer.h"
#include "nvblox/sensors/image.h"
#include "nvblox_b200.h"
namespace nvblox {
class EsdfSlicer {
 public:
  EsdfSlicer() = default;
  // getAabbOfLayerAtHeight(layer, slice_height) (:36-37): the AABB of the ESDF blocks at that height (empty if there are none)
  AxisAlignedBoundingBox getAabbOfLayerAtHeight(const EsdfLayer& layer, float slice_height) {
    float a[6];
    int32_t rows = 0, cols = 0;
    b200_detail::check(nvb_esdf_slice_distance_image(layer.mapper_handle(), slice_height, 0.0f, a, nullptr, nullptr, 0, &rows, &cols),
                       "getAabbOfLayerAtHeight", nvb_last_error());
    if (rows == 0 || cols == 0) return AxisAlignedBoundingBox();
    return AxisAlignedBoundingBox(Vector3f(a[0], a[1], a[2]), Vector3f(a[3], a[4], a[5]));
  }
  // sliceLayerToDistanceImage(layer, slice_height, unobserved_value, &aabb, &image) (:52-60): one pixel per voxel over the AABB,
  // rows along y, columns along x; an empty layer gives an empty AABB and a 0 x 0 image.
  void sliceLayerToDistanceImage(const EsdfLayer& layer, float slice_height, float unobserved_value, AxisAlignedBoundingBox* aabb,
                                 Image<float>* output_image) {
    slice(layer, slice_height, unobserved_value, aabb, output_image, nullptr);
  }
  // The occupancy grid of the same slice (occupancyGridFromSliceImage, :78-82): 100 occupied (distance < 1 cm), 0 free, -1 unknown.
  void sliceLayerToOccupancyGrid(const EsdfLayer& layer, float slice_height, float unobserved_value, AxisAlignedBoundingBox* aabb,
                                 Image<float>* output_image, std::vector<int8_t>* occupancy_grid_data) {
    slice(layer, slice_height, unobserved_value, aabb, output_image, occupancy_grid_data);
  }

 private:
  void slice(const EsdfLayer& layer, float slice_height, float unobserved_value, AxisAlignedBoundingBox* aabb, Image<float>* image,
             std::vector<int8_t>* grid) {
    float a[6];
    int32_t rows = 0, cols = 0;
    NvbMapper* m = layer.mapper_handle();
    b200_detail::check(nvb_esdf_slice_distance_image(m, slice_height, unobserved_value, a, nullptr, nullptr, 0, &rows, &cols),
                       "sliceLayerToDistanceImage", nvb_last_error());
    *image = Image<float>(rows, cols, MemoryType::kHost);
    if (grid) grid->assign((size_t)rows * cols, 0);
    if (rows == 0 || cols == 0) {
      if (aabb) *aabb = AxisAlignedBoundingBox();
      return;
    }
    b200_detail::check(nvb_esdf_slice_distance_image(m, slice_height, unobserved_value, a, image->dataPtr(), grid ? grid->data() : nullptr,
                                                     rows * cols, &rows, &cols),
                       "sliceLayerToDistanceImage", nvb_last_error());
    if (aabb) *aabb = AxisAlignedBoundingBox(Vector3f(a[0], a[1], a[2]), Vector3f(a[3], a[4], a[5]));
  }
};
}  // namespace nvblox
