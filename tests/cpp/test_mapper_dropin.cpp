// Reads like the reference's TsdfIntegratorTestFixture tests (tests/test_tsdf_integrator.cpp:107-188,379-510):
// a plane at z = 5 m in front of an identity-pose camera, integrated through nvblox::Mapper, then checked
// voxel by voxel; then Mapper::updateEsdf and the validateEsdf-style invariants
// (tests/test_esdf_integrator.cpp:339-460). Exit code 0 = pass.
#include <cmath>
#include <cstdio>
#include "nvblox/nvblox.h"
using namespace nvblox;
#define EXPECT(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)
int main() {
  if (nvb_device_count() == 0) { std::fprintf(stderr, "no CUDA device\n"); return 77; }
  constexpr float kVoxel = 0.05f;
  Camera camera(300.f, 300.f, 320.f, 240.f, 640, 480);
  DepthImage depth(480, 640, MemoryType::kUnified);
  for (int r = 0; r < 480; r++) for (int c = 0; c < 640; c++) depth(r, c) = 5.0f;
  Mapper mapper(kVoxel);
  mapper.tsdf_integrator().max_weight(100.0f);
  EXPECT(mapper.tsdf_integrator().weighting_function_type() == WeightingFunctionType::kInverseSquareWeight);
  std::vector<Index3D> updated;
  TsdfLayer tsdf = mapper.tsdf_layer();
  mapper.tsdf_integrator().integrateFrame(MaskedDepthImageConstView(depth, kMaskActiveEverywhere), Transform::Identity(), camera, &tsdf, &updated);
  EXPECT(!updated.empty());
  EXPECT((int)updated.size() == tsdf.numBlocks());
  const float block = tsdf.block_size();
  long observed = 0;
  for (const Index3D& idx : tsdf.getAllBlockIndices()) {
    auto blk = tsdf.getBlockAtIndexHost(idx);
    EXPECT(blk != nullptr);
    EXPECT(tsdf.getBlockAtIndex(idx) != nullptr);
    for (int x = 0; x < 8; x++) for (int y = 0; y < 8; y++) for (int z = 0; z < 8; z++) {
      const TsdfVoxel& v = blk->voxels[x][y][z];
      if (v.weight > 1e-4f) {
        observed++;
        const float depth_v = idx[2] * block + (z + 0.5f) * kVoxel;
        EXPECT(std::fabs(v.weight - std::fmin(1.0f / (depth_v * depth_v), 100.0f)) < 1e-4f + 1e-5f * v.weight);
        const float expect = std::fmax(-0.2f, std::fmin(0.2f, 5.0f - depth_v));
        EXPECT(std::fabs(v.distance - expect) < 1e-4f);
      }
    }
  }
  EXPECT(observed > 10000);
  mapper.updateEsdf();
  EsdfLayer esdf = mapper.esdf_layer();
  EXPECT(esdf.numBlocks() == tsdf.numBlocks());
  long sites = 0;
  for (const Index3D& idx : esdf.getAllBlockIndices()) {
    auto blk = esdf.getBlockAtIndexHost(idx);
    for (int x = 0; x < 8; x++) for (int y = 0; y < 8; y++) for (int z = 0; z < 8; z++) {
      const EsdfVoxel& v = blk->voxels[x][y][z];
      if (!v.observed) continue;
      if (v.is_site) { sites++; EXPECT(v.squared_distance_vox == 0.0f); }
      else if (v.parent_direction != Index3D::Zero()) {
        const Index3D& p = v.parent_direction;
        EXPECT(v.squared_distance_vox == (float)(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]));
      }
    }
  }
  EXPECT(sites > 1000);
  // colour (test_color_integrator.cpp): a red frame paints the observed surface red, nothing else
  ColorImage red(480, 640, MemoryType::kUnified);
  for (int r = 0; r < 480; r++) for (int c = 0; c < 640; c++) red(r, c) = Color::Red();
  mapper.color_integrator().measurement_weight(0.3f);
  EXPECT(std::fabs(mapper.color_integrator().measurement_weight() - 0.3f) < 1e-7f);
  mapper.integrateColor(red, Transform::Identity(), camera);
  ColorLayer color = mapper.color_layer();
  EXPECT(color.numBlocks() > 0 && color.numBlocks() <= tsdf.numBlocks());
  long painted = 0;
  for (const Index3D& idx : color.getAllBlockIndices()) {
    auto blk = color.getBlockAtIndexHost(idx);
    for (int x = 0; x < 8; x++) for (int y = 0; y < 8; y++) for (int z = 0; z < 8; z++) {
      const ColorVoxel& v = blk->voxels[x][y][z];
      if (v.weight > 1e-4f) { painted++; EXPECT(v.color == Color::Red()); EXPECT(std::fabs(v.weight - 0.3f) < 1e-6f); }
      else EXPECT(v.color == Color::Gray());
    }
  }
  EXPECT(painted > 10000);
  // decay (test_tsdf_decay.cpp): the last view is spared, everything decays away without it
  const int blocks_before = tsdf.numBlocks();
  mapper.tsdf_decay_integrator().decay_factor(0.1f);
  EXPECT(std::fabs(mapper.tsdf_decay_integrator().decay_factor() - 0.1f) < 1e-7f);
  for (int i = 0; i < 6; i++) mapper.decayTsdfExcludeLastView();
  EXPECT(tsdf.numBlocks() > 0 && tsdf.numBlocks() <= blocks_before);
  for (int i = 0; i < 6 && tsdf.numBlocks() > 0; i++) mapper.decayTsdfAllVoxels();
  EXPECT(tsdf.numBlocks() == 0);
  EXPECT(esdf.numBlocks() == 0);
  EXPECT(color.numBlocks() == 0);
  std::printf("drop-in C++ API ok: %zu blocks, %ld observed voxels, %ld sites\n", updated.size(), observed, sites);
  return 0;
}
