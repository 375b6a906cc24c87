// nvblox/core/cuda_stream.h -- CudaStream / CudaStreamOwning (reference: nvblox/include/nvblox/core/cuda_stream.h:27-106) as far
// as the depth-integration path's interfaces mention them.
//
// In the reference every Mapper works on the stream it is handed. Here a mapper owns its three streams (compute, copy, ESDF
// side stream: DESIGN.md section 7) inside libnvblox_b200.so and the calls of this header set are synchronous, so a stream passed
// to a constructor is accepted and kept for source compatibility only; synchronize() is therefore a no-op (there is never
// outstanding work the caller could observe). The header does not pull in the CUDA runtime: cudaStream_t is spelled as the opaque
// pointer it is.
#pragma once
#include <memory>
namespace nvblox {
class CudaStream {
 public:
  virtual ~CudaStream() = default;
  void synchronize() const {}
  void* get() const { return stream_; }  // cudaStream_t
  operator void*() const { return stream_; }
 protected:
  void* stream_ = nullptr;  // the default stream
};
class CudaStreamOwning : public CudaStream {
 public:
  explicit CudaStreamOwning(unsigned int /*cudaStreamFlags*/ = 1u) {}
};
class CudaStreamNonOwning : public CudaStream {
 public:
  explicit CudaStreamNonOwning(void* stream) { stream_ = stream; }
};
}  // namespace nvblox
