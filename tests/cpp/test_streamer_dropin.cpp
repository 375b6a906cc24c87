// Reads like the reference's serializer / streamer tests (tests/test_layer_serializer_gpu.cpp:43-95, tests/test_mesh_serializer.cpp:
// 131-210, tests/test_layer_streamer.cpp:66-300) through include/nvblox/serialization/: a plane seen by a camera, integrated and
// meshed through nvblox::Mapper, then serialized and streamed. Exit code 0 = pass.
#include <cmath>
#include <cstdio>
#include <numeric>
#include "nvblox/nvblox.h"
using namespace nvblox;
#define EXPECT(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

// SimpleLayerStreamer (test_layer_streamer.cpp:24-47): priority = x index
template <class LayerType>
class SimpleLayerStreamer : public LayerStreamerBase<LayerType> {
 public:
  const Index3DSet& index_set() const { return LayerStreamerBase<LayerType>::index_set_; }
 protected:
  std::vector<float> computePriorities(const std::vector<Index3D>& b) const override {
    std::vector<float> p;
    for (const Index3D& i : b) p.push_back(static_cast<float>(i[0]));
    return p;
  }
};

int main() {
  if (nvb_device_count() == 0) { std::fprintf(stderr, "no CUDA device\n"); return 77; }
  // ---- host logic (no layer needed)
  {
    SimpleLayerStreamer<TsdfLayer> s;
    std::vector<Index3D> idx;
    for (int i = 0; i < 100; i++) idx.push_back(Index3D((i * 37) % 101 - 50, i % 7, i % 3));
    s.markIndicesCandidates(idx);
    EXPECT(s.numCandidates() == 100);
    const std::vector<Index3D> out = s.getNBlocks(95);  // SimplePriorityTest
    EXPECT(out.size() == 95);
    for (size_t i = 1; i < out.size(); i++) EXPECT(out[i][0] <= out[i - 1][0]);
    for (const Index3D& k : s.index_set()) EXPECT(k[0] <= out.back()[0]);
    EXPECT(s.index_set().size() == 5);
    EXPECT(s.getNBlocks(0).empty() && s.numCandidates() == 5);  // RequestZero
    EXPECT(s.getNBlocks(100).size() == 5 && s.numCandidates() == 0);  // RequestMoreThanAvailable
  }
  {
    TsdfLayerStreamerOldestBlocks s;  // LayerStreamerOldestBlocks (:177-208)
    std::vector<Index3D> idx;
    for (int i = 0; i < 100; i++) idx.push_back(Index3D(i, -i, i % 5));
    s.markIndicesCandidates(idx);
    const std::vector<Index3D> a = s.getNBlocks(50, BlockExclusionParams()), b = s.getNBlocks(50, BlockExclusionParams());
    EXPECT(a.size() == 50 && b.size() == 50 && s.numCandidates() == 0);
    s.markIndicesCandidates(a), s.markIndicesCandidates(b);
    const std::vector<Index3D> a2 = s.getNBlocks(50, BlockExclusionParams()), b2 = s.getNBlocks(50, BlockExclusionParams());
    Index3DSet sa(a.begin(), a.end()), sb(b.begin(), b.end());
    for (const Index3D& k : a2) EXPECT(sa.count(k) == 1);
    for (const Index3D& k : b2) EXPECT(sb.count(k) == 1);
    BlockExclusionParams p;  // blocks whose low face is above 1 m are excluded (0.4 m blocks: z index 3 and up)
    p.exclusion_height_m = 1.0f, p.block_size_m = 0.4f;
    s.markIndicesCandidates(idx);
    for (const Index3D& k : s.getNBlocks(1000, p)) EXPECT(k[2] <= 2);
    EXPECT(s.numCandidates() == 0);
  }
  // ---- a real map: plane at z = 5 m
  constexpr float kVoxel = 0.05f;
  Camera camera(300.f, 300.f, 320.f, 240.f, 640, 480);
  DepthImage depth(480, 640, MemoryType::kUnified);
  for (int r = 0; r < 480; r++) for (int c = 0; c < 640; c++) depth(r, c) = 5.0f;
  Mapper mapper(kVoxel);
  mapper.integrateDepth(depth, Transform::Identity(), camera);
  mapper.updateColorMesh();
  TsdfLayer tsdf = mapper.tsdf_layer();
  ColorMeshLayer mesh = mapper.color_mesh_layer();
  // LayerSerializerGpu: serializeAllBlocks / serializeNoBlocks
  const std::vector<Index3D> all = tsdf.getAllBlockIndices();
  TsdfLayerSerializerGpu ser;
  auto st = ser.serialize(tsdf, all, CudaStreamOwning());
  EXPECT(st->block_indices.size() == all.size() && st->block_offsets.size() == all.size() + 1);
  EXPECT(st->voxels.size() == all.size() * 512 && st->block_offsets.back() == (int32_t)st->voxels.size());
  for (size_t i = 0; i < all.size(); i += 17) {
    auto blk = tsdf.getBlockAtIndexHost(all[i]);
    const TsdfVoxel* v = &st->voxels[st->block_offsets[i]];
    for (int q = 0; q < 512; q++) EXPECT(v[q].distance == (&blk->voxels[0][0][0])[q].distance && v[q].weight == (&blk->voxels[0][0][0])[q].weight);
  }
  EXPECT(ser.serialize(tsdf, {}, CudaStreamOwning())->voxels.empty());
  auto miss = ser.serialize(tsdf, {Index3D(999, 999, 999), all[0]}, CudaStreamOwning());
  EXPECT(miss->block_offsets[0] == 0 && miss->block_offsets[1] == 0 && miss->block_offsets[2] == 512);
  // MeshSerializerGpu: every block equals the mesh block
  const std::vector<Index3D> mall = mesh.getAllBlockIndices();
  EXPECT(!mall.empty());
  ColorMeshSerializerGpu mser;
  auto sm = mser.serialize(mesh, mall, CudaStreamOwning());
  EXPECT(sm->vertex_block_offsets.size() == mall.size() + 1 && sm->triangle_index_block_offsets.size() == mall.size() + 1);
  EXPECT((size_t)sm->vertex_block_offsets.back() == sm->vertices.size() && sm->vertices.size() == sm->vertex_appearances.size());
  size_t total_bytes = 0;
  const size_t all_mesh_vertices = sm->vertices.size();  // (a serializer hands out its one result object: later calls overwrite it)
  for (size_t i = 0; i < mall.size(); i++) {
    ColorMeshBlock::ConstPtr b = mesh.getBlockAtIndex(mall[i]);
    EXPECT(sm->getNumVerticesInBlock(i) == b->vertices.size() && sm->getNumTriangleIndicesInBlock(i) == b->triangles.size());
    for (size_t q = 0; q < b->vertices.size(); q++) {
      EXPECT(sm->getVertex(i, q)[0] == b->vertices[q][0] && sm->getVertex(i, q)[2] == b->vertices[q][2]);
      EXPECT(sm->getAppearance(i, q).r == b->vertex_appearances[q].r);
    }
    for (size_t q = 0; q < b->triangles.size(); q++) EXPECT(sm->getTriangleIndex(i, q) == b->triangles[q]);
    total_bytes += b->vertices.size() * 24 + b->vertex_appearances.size() * sizeof(Color) + b->triangles.size() * 4;
  }
  EXPECT(mser.serialize(mesh, {}, CudaStreamOwning())->vertices.empty());
  // streamers on layers: SerializeNBytes, StreamNBytes
  TsdfLayerStreamerOldestBlocks ts;
  ts.markIndicesCandidates(all);
  const size_t budget = all.size() * sizeof(TsdfBlock) / 2;
  auto half = ts.getNBytesOfSerializedBlocks(budget, tsdf, BlockExclusionParams(), CudaStreamOwning());
  EXPECT(half->voxels.size() * sizeof(TsdfVoxel) <= budget && !half->voxels.empty());
  ColorMeshLayerStreamerOldestBlocks ms;
  ms.markIndicesCandidates(mall);
  const std::vector<Index3D> h1 = ms.getNBytesOfBlocks(total_bytes / 2, mesh, BlockExclusionParams());
  const std::vector<Index3D> h2 = ms.getNBytesOfBlocks(total_bytes * 2, mesh, BlockExclusionParams());
  EXPECT(!h1.empty() && !h2.empty() && h1.size() + h2.size() == mall.size() && ms.numCandidates() == 0);
  auto bw = ms.estimateBandwidthAndSerialize(mesh, mall, "mesh", BlockExclusionParams(), (int)kLayerStreamerUnlimitedBandwidth, CudaStreamOwning());
  EXPECT(bw->block_indices.size() == mall.size());
  // LayerCakeStreamer (test_layer_cake_streamer.cpp): one streamer per layer type; a type that is not in the cake does nothing
  LayerCakeStreamer cake = LayerCakeStreamer::create<TsdfLayer, ColorMeshLayer>();
  EXPECT(cake.getPtr<TsdfLayer>() != nullptr && cake.getPtr<ColorMeshLayer>() != nullptr && cake.getPtr<EsdfLayer>() == nullptr);
  auto c1 = cake.estimateBandwidthAndSerialize(tsdf, all, "tsdf", BlockExclusionParams(), (int)kLayerStreamerUnlimitedBandwidth, CudaStreamOwning());
  EXPECT(c1 && c1->block_indices.size() == all.size() && c1->voxels.size() == all.size() * 512);
  auto c2 = cake.serializeAllBlocks(mesh, mall, CudaStreamOwning());
  EXPECT(c2 && c2->vertices.size() == all_mesh_vertices && all_mesh_vertices > 0);
  EXPECT(cake.getSerializedLayer<ColorMeshLayer>() == c2);
  EsdfLayer esdf = mapper.esdf_layer();
  EXPECT(!cake.estimateBandwidthAndSerialize(esdf, all, "esdf", BlockExclusionParams(), -1, CudaStreamOwning()));
  cake.add<TsdfLayer>();  // already there: ignored
  EXPECT(cake.get<TsdfLayer>().numCandidates() == 0);
  std::printf("streamer drop-in ok: %zu tsdf blocks, %zu mesh blocks, %zu mesh bytes\n", all.size(), mall.size(), total_bytes);
  return 0;
}
