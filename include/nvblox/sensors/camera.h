// nvblox/sensors/camera.h -- nvblox::Camera (reference: nvblox/include/nvblox/sensors/camera.h:33-203)
// with its optional RadialTangentialDistortionParams (sensors/distortion.h:24-62).
#pragma once
#include <optional>
#include "nvblox/core/types.h"
#include "nvblox_b200.h"
namespace nvblox {
struct RadialDistortionParams {
  float k1 = 0.F, k2 = 0.F, k3 = 0.F, k4 = 0.F, k5 = 0.F, k6 = 0.F;
};
struct TangentialDistortionParams {
  float p1 = 0.F, p2 = 0.F;
};
struct RadialTangentialDistortionParams {
  RadialDistortionParams radial;
  TangentialDistortionParams tangential;
};
class Camera {
 public:
  Camera() = default;
  Camera(float fu, float fv, float cu, float cv, int width, int height,
         std::optional<RadialTangentialDistortionParams> distortion_params = std::nullopt)
      : c_{fu, fv, cu, cv, width, height, 0, 0, 0, 0, 0, 0, 0, 0, 0} {
    if (distortion_params.has_value()) {
      const auto& d = *distortion_params;
      c_.has_distortion = 1;
      c_.k1 = d.radial.k1, c_.k2 = d.radial.k2, c_.k3 = d.radial.k3;
      c_.k4 = d.radial.k4, c_.k5 = d.radial.k5, c_.k6 = d.radial.k6;
      c_.p1 = d.tangential.p1, c_.p2 = d.tangential.p2;
    }
  }
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
  NvbCamera c_{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
};
}  // namespace nvblox
