// nvblox/core/unified_ptr.h -- what VoxelBlock<V>::Ptr / ConstPtr are on this path
// (reference: unified_ptr<T>, nvblox/include/nvblox/core/unified_ptr.h:48-170; BlockType::Ptr, map/blox.h:45-46).
//
// The reference's layers hand out reference-counted pointers into their block pool; a pointer to a kDevice block cannot be
// dereferenced on the host there either. Here blocks live in the mapper's slab, so the handle is a non-owning view: get() is the
// device address (what nvblox_ros's kernels take), valid until the block is deallocated or the layer grows.
#pragma once
#include <cstddef>
#include "nvblox/core/types.h"
namespace nvblox {
template <typename T>
class unified_ptr {
 public:
  unified_ptr() = default;
  unified_ptr(std::nullptr_t) {}
  explicit unified_ptr(T* device_ptr) : p_(device_ptr) {}
  template <typename U>
  unified_ptr(const unified_ptr<U>& o) : p_(o.get()) {}
  T* get() const { return p_; }
  explicit operator bool() const { return p_ != nullptr; }
  bool operator==(std::nullptr_t) const { return p_ == nullptr; }
  bool operator!=(std::nullptr_t) const { return p_ != nullptr; }
  MemoryType memory_type() const { return MemoryType::kDevice; }
 private:
  T* p_ = nullptr;
};
}  // namespace nvblox
