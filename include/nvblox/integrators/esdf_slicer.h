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
    int32_t empty = 1;
    b200_detail::check(nvb_esdf_slice_aabb(layer.mapper_handle(), slice_height, a, &empty), "getAabbOfLayerAtHeight", nvb_last_error());
    if (empty) return AxisAlignedBoundingBox();
    return AxisAlignedBoundingBox(Vector3f(a[0], a[1], a[2]), Vector3f(a[3], a[4], a[5]));
  }
  // getCombinedAabbOfLayersAtHeight (:37-47): the box that encloses both layers' slices
  AxisAlignedBoundingBox getCombinedAabbOfLayersAtHeight(const EsdfLayer& layer_1, const EsdfLayer& layer_2, float layer_1_slice_height,
                                                         float layer_2_slice_height) {
    return getAabbOfLayerAtHeight(layer_1, layer_1_slice_height).merged(getAabbOfLayerAtHeight(layer_2, layer_2_slice_height));
  }
  // sliceLayerToDistanceImage(layer, slice_height, unobserved_value, aabb, &image) (:62-76): a GIVEN box
  void sliceLayerToDistanceImage(const EsdfLayer& layer, float slice_height, float unobserved_value, const AxisAlignedBoundingBox& aabb,
                                 Image<float>* output_image) {
    if (aabb.isEmpty()) return;
    const float a[6] = {aabb.min()[0], aabb.min()[1], aabb.min()[2], aabb.max()[0], aabb.max()[1], aabb.max()[2]};
    int32_t rows = 0, cols = 0;
    NvbMapper* m = layer.mapper_handle();
    b200_detail::check(nvb_esdf_slice_distance_image_in_aabb(m, slice_height, unobserved_value, a, nullptr, nullptr, 0, &rows, &cols),
                       "sliceLayerToDistanceImage", nvb_last_error());
    *output_image = Image<float>(rows, cols, MemoryType::kHost);
    if (rows == 0 || cols == 0) return;
    b200_detail::check(nvb_esdf_slice_distance_image_in_aabb(m, slice_height, unobserved_value, a, output_image->dataPtr(), nullptr, rows * cols,
                                                             &rows, &cols),
                       "sliceLayerToDistanceImage", nvb_last_error());
  }
  // sliceLayersToCombinedDistanceImage (:78-118; src/integrators/esdf_slicer.cu:201-240): two layers (two mappers here, e.g.
  // MultiMapper's static and dynamic maps) sliced on the merged box of their slices, element-wise minimum
  void sliceLayersToCombinedDistanceImage(const EsdfLayer& layer_1, const EsdfLayer& layer_2, float layer_1_slice_height,
                                          float layer_2_slice_height, float unobserved_value, AxisAlignedBoundingBox* aabb,
                                          Image<float>* output_image) {
    *aabb = getCombinedAabbOfLayersAtHeight(layer_1, layer_2, layer_1_slice_height, layer_2_slice_height);
    sliceLayersToCombinedDistanceImage(layer_1, layer_2, layer_1_slice_height, layer_2_slice_height, unobserved_value, *aabb, output_image);
  }
  void sliceLayersToCombinedDistanceImage(const EsdfLayer& layer_1, const EsdfLayer& layer_2, float layer_1_slice_height,
                                          float layer_2_slice_height, float unobserved_value, const AxisAlignedBoundingBox& aabb,
                                          Image<float>* output_image) {
    Image<float> slice_1(0, 0, MemoryType::kHost);
    sliceLayerToDistanceImage(layer_1, layer_1_slice_height, unobserved_value, aabb, &slice_1);
    sliceLayerToDistanceImage(layer_2, layer_2_slice_height, unobserved_value, aabb, output_image);
    if (aabb.isEmpty()) return;
    for (int r = 0; r < output_image->rows(); r++)
      for (int c = 0; c < output_image->cols(); c++)
        if (slice_1(r, c) < (*output_image)(r, c)) (*output_image)(r, c) = slice_1(r, c);  // elementWiseMinInPlace
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
