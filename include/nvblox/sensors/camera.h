// nvblox/sensors/camera.h -- nvblox::Camera (reference: nvblox/include/nvblox/sensors/camera.h:33-203),
// pinhole without distortion.
#pragma once
#include "nvblox/core/types.h"
#include "nvblox_b200.h"
namespace nvblox {
class Camera {
 public:
  Camera() = default;
  Camera(float fu, float fv, float cu, float cv, int width, int height) : c_{fu, fv, cu, cv, width, height} {}
  float fu() const { return c_.fu; }
  float fv() const { return c_.fv; }
  float cu() const { return c_.cu; }
  float cv() const { return c_.cv; }
  int width() const { return c_.width; }
  int height() const { return c_.height; }
  int cols() const { return c_.width; }
  int rows() const { return c_.height; }
  float getDepth(const Vector3f& p_C) const { return p_C[2]; }
  const NvbCamera* c_abi() const { return &c_; }
 private:
  NvbCamera c_{0, 0, 0, 0, 0, 0};
};
}  // namespace nvblox
