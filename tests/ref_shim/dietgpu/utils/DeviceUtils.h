// shim (test support only): the few device helpers the reference's tests use.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CUDA_VERIFY(X)                                                                        \
  do {                                                                                        \
    cudaError_t err__ = (X);                                                                  \
    if (err__ != cudaSuccess) {                                                               \
      std::fprintf(stderr, "CUDA error %d (%s) at %s:%d\n", (int)err__, cudaGetErrorString(err__), __FILE__, __LINE__); \
      std::abort();                                                                           \
    }                                                                                         \
  } while (0)

namespace dietgpu {

inline int getCurrentDevice() {
  int d = 0;
  CUDA_VERIFY(cudaGetDevice(&d));
  return d;
}

// owning stream handle, convertible to cudaStream_t
class CudaStream {
 public:
  explicit CudaStream(int flags = cudaStreamDefault) { CUDA_VERIFY(cudaStreamCreateWithFlags(&stream_, flags)); }
  CudaStream(const CudaStream&) = delete;
  CudaStream& operator=(const CudaStream&) = delete;
  CudaStream(CudaStream&& o) noexcept : stream_(o.stream_) { o.stream_ = nullptr; }
  CudaStream& operator=(CudaStream&& o) noexcept {
    if (this != &o) {
      reset();
      stream_ = o.stream_;
      o.stream_ = nullptr;
    }
    return *this;
  }
  ~CudaStream() { reset(); }
  cudaStream_t get() { return stream_; }
  operator cudaStream_t() { return stream_; }
  static CudaStream make() { return CudaStream(); }
  static CudaStream makeNonBlocking() { return CudaStream(cudaStreamNonBlocking); }

 private:
  void reset() {
    if (stream_) {
      cudaStreamSynchronize(stream_);
      cudaStreamDestroy(stream_);
      stream_ = nullptr;
    }
  }
  cudaStream_t stream_ = nullptr;
};

}  // namespace dietgpu
