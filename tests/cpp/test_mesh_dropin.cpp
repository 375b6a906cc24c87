// Reads like the reference's MeshTest cases (tests/test_mesh.cpp:101-155 PlaneMesh, :181-250 GPUPlaneTest, :423-505 welding;
// tests/test_mesh_appearance.cpp) through include/nvblox/: a plane at z = 5 m seen by an identity-pose camera, integrated
// through nvblox::Mapper, meshed with Mapper::updateColorMesh and with the ColorMeshIntegrator directly. Exit code 0 = pass.
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
  EXPECT(mapper.color_mesh_layer().numBlocks() == 0);  // BlankMap
  mapper.integrateDepth(depth, Transform::Identity(), camera);
  ColorMeshIntegrator integrator = mapper.color_mesh_integrator();
  EXPECT(integrator.weld_vertices() == true);
  EXPECT(std::fabs(integrator.min_weight() - 1e-4f) < 1e-9f);
  integrator.weld_vertices(false);
  mapper.updateColorMesh();
  ColorMeshLayer mesh = mapper.color_mesh_layer();
  TsdfLayer tsdf = mapper.tsdf_layer();
  EXPECT(mesh.numBlocks() > 0 && mesh.numBlocks() <= tsdf.numBlocks());
  size_t unwelded = 0;
  for (const Index3D& idx : mesh.getAllBlockIndices()) {
    ColorMeshBlock::ConstPtr b = mesh.getBlockAtIndex(idx);
    EXPECT(b != nullptr && mesh.isBlockAllocated(idx));
    EXPECT(b->vertices.size() > 0);
    EXPECT(b->vertices.size() == b->vertex_normals.size());
    EXPECT(b->vertices.size() == b->triangles.size());       // unwelded: three fresh vertices per triangle
    EXPECT(b->vertex_appearances.size() == b->vertices.size());
    unwelded += b->vertices.size();
    for (size_t i = 0; i < b->vertices.size(); i++) {
      EXPECT(std::fabs(b->vertices[i][2] - 5.0f) < 1e-3f);    // the plane
      EXPECT(std::fabs(b->vertex_normals[i][2] + 1.0f) < 1e-3f);  // facing the camera
      EXPECT(b->vertex_appearances[i].r == 127 && b->vertex_appearances[i].g == 127);  // no colour layer: Color::Gray()
    }
  }
  EXPECT(!mesh.isBlockAllocated(Index3D(1000, 1000, 1000)));
  EXPECT(mesh.getBlockAtIndex(Index3D(1000, 1000, 1000)) == nullptr);
  // welding through the integrator on an explicit list (InPlaceWeldingTest): fewer vertices, same triangle count
  integrator.weld_vertices(true);
  ColorMeshLayer out = mapper.color_mesh_layer();
  EXPECT(integrator.integrateMeshFromDistanceField(tsdf, &out, DeviceType::kGPU));
  size_t welded = 0, tris = 0;
  for (const Index3D& idx : out.getAllBlockIndices()) {
    ColorMeshBlock::ConstPtr b = out.getBlockAtIndex(idx);
    welded += b->vertices.size(), tris += b->triangles.size();
    for (int t : b->triangles) EXPECT(t >= 0 && (size_t)t < b->vertices.size());
  }
  EXPECT(welded < unwelded && tris == unwelded);
  // colours: a red frame paints the vertices red (a vertex takes the colour of the voxel it falls into, painted or not:
  // updateAppearanceBlockByClosestVoxel does not look at the colour weight)
  ColorImage red(480, 640, MemoryType::kUnified);
  for (int r = 0; r < 480; r++) for (int c = 0; c < 640; c++) red(r, c) = Color(255, 0, 0);
  mapper.integrateColor(red, Transform::Identity(), camera);
  integrator.updateAppearance(mapper.color_layer(), &out);
  size_t red_vertices = 0, all_vertices = 0;
  for (const Index3D& idx : out.getAllBlockIndices()) {
    ColorMeshBlock::ConstPtr b = out.getBlockAtIndex(idx);
    EXPECT(b->vertex_appearances.size() == b->vertices.size());
    for (const Color& c : b->vertex_appearances) {
      all_vertices++, red_vertices += (c.r == 255 && c.g == 0 && c.b == 0);
      EXPECT((c.r == 255 && c.g == 0 && c.b == 0) || (c.r == 0 && c.g == 0 && c.b == 0) || (c.r == 127 && c.g == 127 && c.b == 127));
    }
  }
  EXPECT(red_vertices * 4 > all_vertices);
  std::printf("mesh drop-in ok: %zu blocks, %zu -> %zu vertices, %zu triangles\n", out.size(), unwelded, welded, tris / 3);
  return 0;
}
