// nvblox/map/voxels.h -- TsdfVoxel / OccupancyVoxel / FreespaceVoxel / ColorVoxel / EsdfVoxel with the reference's layout
// (nvblox/include/nvblox/map/voxels.h:28-74); these are the bytes stored in HBM.
#pragma once
#include <cstdint>
#include "nvblox/core/types.h"
namespace nvblox {
struct TsdfVoxel {
  float distance = 0.0f;
  float weight = 0.0f;
};
struct OccupancyVoxel {
  float log_odds = 0.0f;
};
// nvblox::Time is a strongly typed int64 of milliseconds (core/time.h:25-70); plain int64_t here.
using Time = int64_t;
struct FreespaceVoxel {
  Time last_occupied_timestamp_ms = 0;
  Time consecutive_occupancy_duration_ms = 0;
  bool is_high_confidence_freespace = false;
};
// Color (core/color.h:28-68): three bytes, RGB.
struct Color {
  uint8_t r = 0, g = 0, b = 0;
  Color() = default;
  Color(uint8_t r_, uint8_t g_, uint8_t b_) : r(r_), g(g_), b(b_) {}
  bool operator==(const Color& o) const { return r == o.r && g == o.g && b == o.b; }
  static Color Gray() { return Color(127, 127, 127); }
  static Color Red() { return Color(255, 0, 0); }
  static Color Green() { return Color(0, 255, 0); }
  static Color Blue() { return Color(0, 0, 255); }
  static Color White() { return Color(255, 255, 255); }
  static Color Black() { return Color(0, 0, 0); }
};
struct ColorVoxel {
  Color color = Color::Gray();
  float weight = 0.0f;
};
struct EsdfVoxel {
  float squared_distance_vox = 0.0f;
  Index3D parent_direction = Index3D::Zero();
  bool is_inside = false;
  bool observed = false;
  bool is_site = false;
};
static_assert(sizeof(TsdfVoxel) == 8, "TsdfVoxel layout");
static_assert(sizeof(OccupancyVoxel) == 4, "OccupancyVoxel layout");
static_assert(sizeof(FreespaceVoxel) == 24, "FreespaceVoxel layout");
static_assert(sizeof(EsdfVoxel) == 20, "EsdfVoxel layout");
static_assert(sizeof(Color) == 3 && sizeof(ColorVoxel) == 8, "ColorVoxel layout");
}  // namespace nvblox
