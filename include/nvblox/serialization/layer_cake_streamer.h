// nvblox/serialization/layer_cake_streamer.h -- LayerCakeStreamer (reference: nvblox/include/nvblox/serialization/
// layer_cake_streamer.h:25-98, internal/impl/layer_cake_streamer_impl.h): one LayerStreamerOldestBlocks per layer TYPE behind one
// object; a request for a layer type that is not in the cake does nothing and returns an empty pointer.
#pragma once
#include <memory>
#include <string>
#include <typeindex>
#include <unordered_map>
#include <vector>
#include "nvblox/serialization/layer_streamer.h"
namespace nvblox {
class LayerCakeStreamer {
 public:
  LayerCakeStreamer() = default;
  template <typename LayerType>
  std::shared_ptr<SerializedLayerType<LayerType>> estimateBandwidthAndSerialize(const LayerType& layer, const std::vector<Index3D>& blocks_to_serialize,
                                                                              const std::string& layer_name, const BlockExclusionParams& p,
                                                                              const int bandwidth_limit_mbps, const CudaStream& cuda_stream) {
    auto* s = getPtr<LayerType>();
    if (!s) return std::shared_ptr<SerializedLayerType<LayerType>>();
    return s->estimateBandwidthAndSerialize(layer, blocks_to_serialize, layer_name, p, bandwidth_limit_mbps, cuda_stream);
  }
  template <typename LayerType>
  std::shared_ptr<SerializedLayerType<LayerType>> serializeAllBlocks(const LayerType& layer, const std::vector<Index3D>& block_indices,
                                                                   const CudaStream& cuda_stream) {
    auto* s = getPtr<LayerType>();
    if (!s) return std::shared_ptr<SerializedLayerType<LayerType>>();
    return s->serializeAllBlocks(layer, block_indices, cuda_stream);
  }
  template <typename LayerType>
  std::shared_ptr<SerializedLayerType<LayerType>> getSerializedLayer() {
    auto* s = getPtr<LayerType>();
    return s ? s->getSerializedLayer() : std::shared_ptr<SerializedLayerType<LayerType>>();
  }
  // one streamer of each type at most (a second add() of the same type is ignored)
  template <typename LayerType>
  void add() {
    if (streamers_.count(typeid(LayerType)) == 0)
      streamers_.emplace(std::type_index(typeid(LayerType)), std::make_unique<LayerStreamerOldestBlocks<LayerType>>());
  }
  template <typename LayerType>
  LayerStreamerOldestBlocks<LayerType>* getPtr() {
    auto it = streamers_.find(std::type_index(typeid(LayerType)));
    return it == streamers_.end() ? nullptr : dynamic_cast<LayerStreamerOldestBlocks<LayerType>*>(it->second.get());
  }
  template <typename LayerType>
  const LayerStreamerOldestBlocks<LayerType>* getConstPtr() const {
    auto it = streamers_.find(std::type_index(typeid(LayerType)));
    return it == streamers_.end() ? nullptr : dynamic_cast<const LayerStreamerOldestBlocks<LayerType>*>(it->second.get());
  }
  template <typename LayerType>
  const LayerStreamerOldestBlocks<LayerType>& get() const {
    const auto* p = getConstPtr<LayerType>();
    b200_detail::check(p ? 0 : 1, "LayerCakeStreamer::get", "layer type not in the cake");
    return *p;
  }
  template <typename... LayerTypes>
  static LayerCakeStreamer create() {
    LayerCakeStreamer cake;
    (cake.template add<LayerTypes>(), ...);
    return cake;
  }
 private:
  std::unordered_map<std::type_index, std::unique_ptr<LayerStreamerInterface>> streamers_;
};
}  // namespace nvblox
