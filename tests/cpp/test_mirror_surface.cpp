// Compile-and-link coverage of the mirror headers' wider surface (planar slices, EsdfSlicer, Plane): built by the CPU test suite
// (tests/test_cabi_symbols.py); with a GPU it runs a small 2-D ESDF end to end.
#include <cmath>
#include <cstdio>
#include <vector>
#include "nvblox/nvblox.h"

using namespace nvblox;

#define EXPECT(c) do { if (!(c)) { std::fprintf(stderr, "EXPECT failed: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main() {
  if (nvb_device_count() == 0) { std::fprintf(stderr, "no CUDA device\n"); return 77; }
  Camera camera(300.f, 300.f, 320.f, 240.f, 640, 480);
  DepthImage depth(480, 640, MemoryType::kUnified);
  for (int r = 0; r < 480; r++) for (int c = 0; c < 640; c++) depth(r, c) = 3.0f;
  Mapper mapper(0.05f);
  // camera looking along +x (Quaternionf(0.5, 0.5, 0.5, 0.5)), 1 m above the ground
  Transform T;
  const float R[3][3] = {{0, 0, 1}, {1, 0, 0}, {0, 1, 0}};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T.m[j * 4 + i] = R[i][j];
  T.m[14] = 1.0f;
  mapper.integrateDepth(depth, T, camera);
  const Plane ground(Vector3f(0.0f, 0.0f, 1.0f), Vector3f(0.0f, 0.0f, 0.0f));
  EXPECT(std::fabs(ground.getHeightAtXY(Vector2f(3.0f, -2.0f))) < 1e-6f);
  mapper.esdf_integrator().slice_height_above_plane_m(0.3f);
  mapper.esdf_integrator().slice_height_thickness_m(1.2f);
  mapper.esdf_integrator().esdf_slice_height(1.0f);
  mapper.updateEsdfSlice(UpdateFullLayer::kNo, ground);
  EsdfLayer esdf = mapper.esdf_layer();
  EXPECT(esdf.numBlocks() > 0);
  EsdfSlicer slicer;
  AxisAlignedBoundingBox aabb;
  Image<float> image(0, 0);
  std::vector<int8_t> grid;
  slicer.sliceLayerToOccupancyGrid(esdf, 1.0f, 1000.0f, &aabb, &image, &grid);
  EXPECT(!aabb.isEmpty() && image.rows() > 0 && image.cols() > 0 && (int)grid.size() == image.numel());
  EXPECT(!slicer.getAabbOfLayerAtHeight(esdf, 1.0f).isEmpty());
  EXPECT(slicer.getAabbOfLayerAtHeight(esdf, 40.0f).isEmpty());
  int occupied = 0;
  for (int8_t g : grid) occupied += g == 100;
  EXPECT(occupied > 0);  // the wall 3 m ahead crosses the band
  {
    // sliceLayersToCombinedDistanceImage (esdf_slicer.h:78-118): the same layer twice is the layer's own image; the given-box
    // overload on the layer's own box too
    AxisAlignedBoundingBox both;
    Image<float> combined(0, 0), own(0, 0);
    slicer.sliceLayersToCombinedDistanceImage(esdf, esdf, 1.0f, 1.0f, 1000.0f, &both, &combined);
    slicer.sliceLayerToDistanceImage(esdf, 1.0f, 1000.0f, aabb, &own);
    EXPECT(!both.isEmpty() && combined.rows() == image.rows() && combined.cols() == image.cols() && own.rows() == image.rows());
    for (int r = 0; r < image.rows(); r++)
      for (int c = 0; c < image.cols(); c++) EXPECT(combined(r, c) == image(r, c) && own(r, c) == image(r, c));
    EXPECT(slicer.getCombinedAabbOfLayersAtHeight(esdf, esdf, 1.0f, 40.0f).max()[0] == aabb.max()[0]);
  }
  std::vector<Index3D> blocks = mapper.tsdf_layer().getAllBlockIndices();
  TsdfLayer tsdf = mapper.tsdf_layer();
  mapper.esdf_integrator().integrateSlice(tsdf, blocks, ground, &esdf);
  std::printf("mirror surface ok: %d slice blocks, %d x %d image, %d occupied cells\n", esdf.numBlocks(), image.rows(), image.cols(), occupied);
  return 0;
}
