// nvblox/serialization/layer_streamer.h -- LayerStreamerBase / LayerStreamerOldestBlocks (reference:
// nvblox/include/nvblox/serialization/layer_streamer.h:31-262, internal/impl/layer_streamer_impl.h): which blocks of a layer go
// out under a block / byte / bandwidth budget -- candidates not streamed yet, highest priority first, optional exclusion above
// a height or outside a radius -- and their serialization through the GPU serializers. Host-side bookkeeping, header-only.
// Where the reference leaves an order undefined (unordered_set walk, std::sort on equal priorities) equal priorities go out in
// (x, y, z) order here.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <deque>
#include <functional>
#include <limits>
#include <map>
#include <numeric>
#include <optional>
#include <set>
#include <string>
#include <vector>
#include "nvblox/serialization/layer_serializer_gpu.h"
namespace nvblox {
struct BlockExclusionParams {
  std::optional<Vector3f> exclusion_center_m = std::nullopt;
  std::optional<float> exclusion_height_m = std::nullopt;
  std::optional<float> exclusion_radius_m = std::nullopt;
  std::optional<float> block_size_m = std::nullopt;
};
using ExcludeBlockFunctor = std::function<bool(const Index3D&)>;
namespace b200_detail {
struct Index3DLess {
  bool operator()(const Index3D& a, const Index3D& b) const {
    if (a[0] != b[0]) return a[0] < b[0];
    if (a[1] != b[1]) return a[1] < b[1];
    return a[2] < b[2];
  }
};
// sizeInBytes of a block for the byte budgets; false if the layer does not hold the block
template <typename V>
inline bool blockBytes(const VoxelBlockLayer<V>& layer, const Index3D& idx, size_t* bytes) {
  if (layer.getBlockAtIndex(idx) == nullptr) return false;
  *bytes = sizeof(VoxelBlock<V>);
  return true;
}
template <typename A>
inline bool blockBytes(const MeshBlockLayer<A>& layer, const Index3D& idx, size_t* bytes) {
  const int32_t k[3] = {idx[0], idx[1], idx[2]};
  int32_t sz[3] = {-1, -1, -1};
  check(nvb_mesh_block_sizes(layer.mapper_handle(), k, 1, sz), "sizeInBytes", nvb_last_error());
  if (sz[0] < 0) return false;
  // mesh_block.h:86-93: vertices + normals (12 B each), appearances, triangle indices
  *bytes = (size_t)sz[0] * 2 * sizeof(Vector3f) + (size_t)sz[2] * sizeof(A) + (size_t)sz[1] * sizeof(int);
  return true;
}
template <class LayerType>
struct StreamerTraits;
template <typename V>
struct StreamerTraits<VoxelBlockLayer<V>> {
  using Serializer = LayerSerializerGpu<VoxelBlockLayer<V>>;
  using Serialized = SerializedLayer<V>;
};
template <typename A>
struct StreamerTraits<MeshBlockLayer<A>> {
  using Serializer = MeshSerializerGpu<A>;
  using Serialized = SerializedMeshLayer<A>;
};
}  // namespace b200_detail
template <class LayerType>
using SerializedLayerType = typename b200_detail::StreamerTraits<LayerType>::Serialized;
using Index3DSet = std::set<Index3D, b200_detail::Index3DLess>;
using BlockIndexToLastPublishedIndexMap = std::map<Index3D, int64_t, b200_detail::Index3DLess>;

class LayerStreamerInterface {
 public:
  virtual ~LayerStreamerInterface() = default;
};

template <class _LayerType>
class LayerStreamerBase : public LayerStreamerInterface {
 public:
  using LayerType = _LayerType;
  virtual ~LayerStreamerBase() = default;
  void markIndicesCandidates(const std::vector<Index3D>& block_indices) { index_set_.insert(block_indices.begin(), block_indices.end()); }
  int numCandidates() const { return (int)index_set_.size(); }
  void clear() { index_set_.clear(); }
  void setExclusionFunctors(std::vector<ExcludeBlockFunctor> f) { exclude_block_functors_ = std::move(f); }
  // getNBlocks (layer_streamer_impl.h:51-67)
  std::vector<Index3D> getNBlocks(const int num_blocks) {
    int streamed = 0;
    return getHighestPriorityBlocks([&streamed, num_blocks](const Index3D&) -> StreamStatus {
      const bool ok = streamed < num_blocks;
      if (ok) ++streamed;
      return {ok, false, !ok};
    });
  }
  // getNBytesOfBlocks (:69-92): the running sum includes the block that crosses the limit, which is not streamed
  std::vector<Index3D> getNBytesOfBlocks(const size_t num_bytes, const LayerType& layer) {
    size_t streamed = 0;
    return getHighestPriorityBlocks([&streamed, num_bytes, &layer](const Index3D& idx) -> StreamStatus {
      size_t bytes = 0;
      if (!b200_detail::blockBytes(layer, idx, &bytes)) return {false, true, false};
      streamed += bytes;
      const bool ok = streamed < num_bytes;
      return {ok, false, !ok};
    });
  }
  std::shared_ptr<SerializedLayerType<LayerType>> getNBytesOfSerializedBlocks(const size_t num_bytes, const LayerType& layer,
                                                                            const CudaStream& cuda_stream) {
    return serializer_.serialize(layer, getNBytesOfBlocks(num_bytes, layer), cuda_stream);
  }
  std::shared_ptr<SerializedLayerType<LayerType>> serializeAllBlocks(const LayerType& layer, const std::vector<Index3D>& block_indices,
                                                                   const CudaStream& cuda_stream) {
    return serializer_.serialize(layer, block_indices, cuda_stream);
  }
  std::shared_ptr<SerializedLayerType<LayerType>> getSerializedLayer() { return serializer_.getSerializedLayer(); }

 protected:
  virtual std::vector<float> computePriorities(const std::vector<Index3D>& block_indices) const = 0;
  struct StreamStatus {
    bool should_block_be_streamed = false;
    bool block_index_invalid = false;
    bool streaming_limit_reached = false;
  };
  using StreamStatusFunctor = std::function<StreamStatus(const Index3D&)>;
  void excludeBlocks(std::vector<Index3D>* block_indices) const {
    if (exclude_block_functors_.empty()) return;
    std::vector<Index3D> kept;
    kept.reserve(block_indices->size());
    for (const Index3D& idx : *block_indices) {
      bool exclude = false;
      for (const ExcludeBlockFunctor& f : exclude_block_functors_)
        if (f(idx)) { exclude = true; break; }
      if (!exclude) kept.push_back(idx);
    }
    *block_indices = std::move(kept);
  }
  // getHighestPriorityBlocks (:104-165): excluded blocks leave the tracking set for good
  std::vector<Index3D> getHighestPriorityBlocks(StreamStatusFunctor get_stream_status) {
    std::vector<Index3D> index_vec(index_set_.begin(), index_set_.end());
    index_set_.clear();
    excludeBlocks(&index_vec);
    const std::vector<float> priorities = computePriorities(index_vec);
    std::vector<int> order(priorities.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](const int a, const int b) { return priorities[a] > priorities[b]; });
    std::vector<Index3D> out;
    int rest_from = -1;
    for (int i = 0; i < (int)index_vec.size(); i++) {
      const Index3D& idx = index_vec[order[i]];
      const StreamStatus s = get_stream_status(idx);
      if (s.should_block_be_streamed) out.push_back(idx);
      else if (!s.block_index_invalid) index_set_.insert(idx);
      if (s.streaming_limit_reached) { rest_from = i + 1; break; }
    }
    if (rest_from > 0)
      for (size_t i = rest_from; i < order.size(); i++) index_set_.insert(index_vec[order[i]]);
    return out;
  }
  std::vector<ExcludeBlockFunctor> exclude_block_functors_;
  Index3DSet index_set_;
  typename b200_detail::StreamerTraits<LayerType>::Serializer serializer_;
};

template <class _LayerType>
class LayerStreamerOldestBlocks : public LayerStreamerBase<_LayerType> {
 public:
  using LayerType = _LayerType;
  using Base = LayerStreamerBase<_LayerType>;
  std::vector<Index3D> getNBlocks(const int num_blocks, const BlockExclusionParams& p) {
    setupExclusionFunctors(p);
    const std::vector<Index3D> out = Base::getNBlocks(num_blocks);
    updateBlocksLastPublishIndex(out);
    return out;
  }
  std::vector<Index3D> getNBytesOfBlocks(const size_t num_bytes, const LayerType& layer, const BlockExclusionParams& p) {
    setupExclusionFunctors(p);
    const std::vector<Index3D> out = Base::getNBytesOfBlocks(num_bytes, layer);
    updateBlocksLastPublishIndex(out);
    return out;
  }
  std::shared_ptr<SerializedLayerType<LayerType>> getNBytesOfSerializedBlocks(const size_t num_bytes, const LayerType& layer,
                                                                            const BlockExclusionParams& p, const CudaStream& cuda_stream) {
    return Base::serializer_.serialize(layer, getNBytesOfBlocks(num_bytes, layer, p), cuda_stream);
  }
  // estimateBandwidthAndSerialize (:314-352): budget of this call = bandwidth limit x the measured period of the calls
  std::shared_ptr<SerializedLayerType<LayerType>> estimateBandwidthAndSerialize(const LayerType& layer, const std::vector<Index3D>& blocks_to_serialize,
                                                                              const std::string& layer_name, const BlockExclusionParams& p,
                                                                              const int bandwidth_limit_mbps, const CudaStream& cuda_stream) {
    (void)layer_name;
    Base::markIndicesCandidates(blocks_to_serialize);
    const double now = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    ticks_.push_back(now);
    if (ticks_.size() > 100) ticks_.pop_front();
    float rate_hz = 0.0f;
    if (ticks_.size() >= 2 && ticks_.back() > ticks_.front()) rate_hz = (float)((ticks_.size() - 1) / (ticks_.back() - ticks_.front()));
    rate_hz = std::max(1.0f, std::min(100.0f, rate_hz));
    const float megabits_per_update = bandwidth_limit_mbps * (1.0f / rate_hz);
    const size_t num_bytes = bandwidth_limit_mbps < 0 ? std::numeric_limits<size_t>::max() : (size_t)(megabits_per_update * (1e6f / 8.0f));
    return getNBytesOfSerializedBlocks(num_bytes, layer, p, cuda_stream);
  }

 protected:
  std::vector<float> computePriorities(const std::vector<Index3D>& block_indices) const override {
    std::vector<float> out;
    out.reserve(block_indices.size());
    for (const Index3D& idx : block_indices) out.push_back(computePriority(idx));
    return out;
  }
  float computePriority(const Index3D& idx) const {
    const auto it = last_published_map_.find(idx);
    if (it == last_published_map_.end()) return static_cast<float>(std::numeric_limits<int64_t>::max());
    return static_cast<float>(-1 * it->second);
  }
  void updateBlocksLastPublishIndex(const std::vector<Index3D>& block_indices) {
    for (const Index3D& idx : block_indices) last_published_map_[idx] = publishing_index_;
    ++publishing_index_;
  }
  void setupExclusionFunctors(const BlockExclusionParams& p) {
    std::vector<ExcludeBlockFunctor> f;
    if (p.exclusion_height_m.has_value() && p.block_size_m.has_value() && p.exclusion_height_m.value() > 0.0f)
      f.push_back(getExcludeAboveHeightFunctor(p.exclusion_height_m.value(), p.block_size_m.value()));
    if (p.block_size_m.has_value() && p.exclusion_center_m.has_value() && p.exclusion_radius_m.has_value() && p.exclusion_radius_m.value() > 0.0f)
      f.push_back(getExcludeOutsideRadiusFunctor(p.exclusion_radius_m.value(), p.exclusion_center_m.value(), p.block_size_m.value()));
    Base::setExclusionFunctors(f);
  }
  static ExcludeBlockFunctor getExcludeAboveHeightFunctor(const float exclusion_height_m, const float block_size_m) {
    return [exclusion_height_m, block_size_m](const Index3D& idx) { return static_cast<float>(idx[2]) * block_size_m > exclusion_height_m; };
  }
  static ExcludeBlockFunctor getExcludeOutsideRadiusFunctor(const float radius_m, const Vector3f& center_m, const float block_size_m) {
    const float r2 = radius_m * radius_m;
    return [r2, center_m, block_size_m](const Index3D& idx) {
      // getCenterPositionFromBlockIndex (core/internal/impl/indexing_impl.h:65-69)
      float d[3];
      for (int a = 0; a < 3; a++) d[a] = block_size_m * (static_cast<float>(idx[a]) + 0.5f) - center_m[a];
      return d[0] * d[0] + (d[1] * d[1] + d[2] * d[2]) > r2;
    };
  }
  int64_t publishing_index_ = 0;
  BlockIndexToLastPublishedIndexMap last_published_map_;
  std::deque<double> ticks_;
};
constexpr float kLayerStreamerUnlimitedBandwidth = -1.0F;
using ColorMeshLayerStreamerOldestBlocks = LayerStreamerOldestBlocks<ColorMeshLayer>;
using TsdfLayerStreamerOldestBlocks = LayerStreamerOldestBlocks<TsdfLayer>;
using EsdfLayerStreamerOldestBlocks = LayerStreamerOldestBlocks<EsdfLayer>;
using OccupancyLayerStreamerOldestBlocks = LayerStreamerOldestBlocks<OccupancyLayer>;
using FreespaceLayerStreamerOldestBlocks = LayerStreamerOldestBlocks<FreespaceLayer>;
using ColorLayerStreamerOldestBlocks = LayerStreamerOldestBlocks<ColorLayer>;
}  // namespace nvblox
