// dietgpu_b200_compat.hpp -- the reference's C++ API surface, header-only, on top of the C ABI.
//
// A caller of facebookresearch/dietgpu's C++ API (dietgpu/ans/GpuANSCodec.h:65-341,
// dietgpu/float/GpuFloatCodec.h:31-292, dietgpu/utils/StackDeviceMemory.h) can include this
// header instead, link libdietgpu_b200.so, and keep its call sites: same namespace, same
// function names, same argument order and meaning, same status types.  Every function forwards
// to one dgb_* entry point of include/dietgpu_b200.h.
//
// Error behaviour mirrors the reference: API misuse / CUDA failure aborts the process with a
// message (the reference's glog CHECK, utils/DeviceUtils.h:33-39); a checksum mismatch is
// returned in the status struct (ans/GpuANSDecode.cuh:581-590).
//
// Written for this repository from the API's documented contract; it shares no source text with
// the reference headers.
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

#include "dietgpu_b200.h"

namespace dietgpu {

constexpr int kANSRequiredAlignment = DGB_ANS_REQUIRED_ALIGNMENT;  // ans/GpuANSCodec.h:16
constexpr int kANSDefaultProbBits = DGB_ANS_DEFAULT_PROB_BITS;     // ans/GpuANSCodec.h:20

namespace detail {
[[noreturn]] inline void fail(const char* what, int code) {
  std::fprintf(stderr, "dietgpu_b200: %s failed: %s (code %d, cudaError %d)\n", what,
               dgb_error_string(code), code, dgb_last_cuda_error());
  std::abort();
}
inline void check(const char* what, int code) {
  if (code != DGB_OK && code != DGB_ERR_CHECKSUM) fail(what, code);
}
}  // namespace detail

// utils/StackDeviceMemory.h: caller-provided scratch with an allocation fallback.  Only what the
// codec API needs is reproduced: construction, usage high-water mark, and (internally) one
// reservation per call.  When the region is too small the call allocates a temporary region with
// cudaMalloc, warns on stderr and frees it after synchronising the stream -- the behaviour of
// utils/StackDeviceMemory.cpp:119-139.
class StackDeviceMemory {
 public:
  StackDeviceMemory(int device, size_t allocPerDevice) : device_(device), owned_(true), size_(allocPerDevice) {
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    if (cudaMalloc(&base_, allocPerDevice) != cudaSuccess) detail::fail("StackDeviceMemory cudaMalloc", DGB_ERR_CUDA);
    cudaSetDevice(prev);
  }
  StackDeviceMemory(int device, void* p, size_t size) : device_(device), owned_(false), base_(p), size_(size) {}
  StackDeviceMemory(const StackDeviceMemory&) = delete;
  StackDeviceMemory& operator=(const StackDeviceMemory&) = delete;
  StackDeviceMemory(StackDeviceMemory&& o) noexcept { *this = std::move(o); }
  StackDeviceMemory& operator=(StackDeviceMemory&& o) noexcept {
    std::swap(device_, o.device_); std::swap(owned_, o.owned_); std::swap(base_, o.base_);
    std::swap(size_, o.size_); std::swap(maxUsed_, o.maxUsed_);
    return *this;
  }
  ~StackDeviceMemory() { if (owned_ && base_) cudaFree(base_); }

  int getDevice() const { return device_; }
  size_t getSizeAvailable() const { return alignedSize(); }
  size_t getSizeTotal() const { return size_; }
  size_t getMaxMemoryUsage() const { return maxUsed_; }
  void resetMaxMemoryUsage() { maxUsed_ = 0; }

  // One scratch region for the duration of a call.
  struct Lease {
    void* ptr = nullptr;
    size_t bytes = 0;
    void* fallback = nullptr;
    cudaStream_t stream = nullptr;
    Lease() = default;
    Lease(const Lease&) = delete;
    Lease(Lease&& o) noexcept { std::swap(ptr, o.ptr); std::swap(bytes, o.bytes); std::swap(fallback, o.fallback); std::swap(stream, o.stream); }
    ~Lease() {
      if (fallback) {
        cudaStreamSynchronize(stream);
        cudaFree(fallback);
      }
    }
  };
  Lease lease(size_t need, cudaStream_t stream) {
    maxUsed_ = std::max(maxUsed_, need);
    Lease l;
    l.stream = stream;
    l.bytes = need;
    if (alignedSize() >= need) {
      l.ptr = alignedBase();
      return l;
    }
    std::fprintf(stderr,
                 "dietgpu_b200 WARNING: temporary memory of %zu bytes requested, %zu available; "
                 "falling back to cudaMalloc (synchronises the stream)\n", need, alignedSize());
    if (cudaMalloc(&l.fallback, need) != cudaSuccess) detail::fail("temporary cudaMalloc", DGB_ERR_CUDA);
    l.ptr = l.fallback;
    return l;
  }

 private:
  void* alignedBase() const {
    auto a = reinterpret_cast<uintptr_t>(base_);
    return reinterpret_cast<void*>((a + 255) & ~uintptr_t(255));
  }
  size_t alignedSize() const {
    if (!base_) return 0;
    size_t pad = reinterpret_cast<uintptr_t>(alignedBase()) - reinterpret_cast<uintptr_t>(base_);
    return size_ > pad ? size_ - pad : 0;
  }
  int device_ = 0;
  bool owned_ = false;
  void* base_ = nullptr;
  size_t size_ = 0;
  size_t maxUsed_ = 0;
};

inline StackDeviceMemory makeStackMemory(size_t bytes = 256 * 1024 * 1024) {
  int dev = 0;
  cudaGetDevice(&dev);
  return StackDeviceMemory(dev, bytes);
}

// ---- ans/GpuANSCodec.h:22-59 ----
inline uint32_t getMaxCompressedSize(uint32_t uncompressedBytes) { return dgb_ans_max_compressed_size(uncompressedBytes); }

struct ANSCodecConfig {
  ANSCodecConfig() : probBits(kANSDefaultProbBits), useChecksum(false) {}
  explicit ANSCodecConfig(int pb, bool checksum = false) : probBits(pb), useChecksum(checksum) {}
  int probBits;
  bool useChecksum;
};

enum class ANSDecodeError : uint32_t { None = 0, ChecksumMismatch = 1 };

struct ANSDecodeStatus {
  ANSDecodeStatus() : error(ANSDecodeError::None) {}
  ANSDecodeError error;
  std::vector<std::pair<int, std::string>> errorInfo;
};

// ---- float/GpuFloatCodec.h:18-97 ----
enum class FloatType : uint32_t { kUndefined = 0, kFloat16 = 1, kBFloat16 = 2, kFloat32 = 3 };

inline uint32_t getMaxFloatCompressedSize(FloatType ft, uint32_t size) {
  return dgb_float_max_compressed_size((int)ft, size);
}

struct FloatCodecConfig {
  FloatCodecConfig() : floatType(FloatType::kFloat16), useChecksum(false), is16ByteAligned(false) {}
  FloatCodecConfig(FloatType ft, const ANSCodecConfig& ansConf, bool align, bool checksum = false)
      : floatType(ft), useChecksum(checksum), ansConfig(ansConf), is16ByteAligned(align) {}
  FloatType floatType;
  bool useChecksum;
  ANSCodecConfig ansConfig;
  bool is16ByteAligned;  // accepted for compatibility; the fused decode needs no alignment hint
};
using FloatCompressConfig = FloatCodecConfig;
using FloatDecompressConfig = FloatCodecConfig;

enum class FloatDecompressError : uint32_t { None = 0, ChecksumMismatch = 1 };

struct FloatDecompressStatus {
  FloatDecompressStatus() : error(FloatDecompressError::None) {}
  FloatDecompressError error;
  std::vector<std::pair<int, std::string>> errorInfo;
};

namespace detail {
template <typename Status, typename Err>
Status makeStatus(int rc, const std::vector<uint8_t>& mismatch, Err mismatchErr) {
  Status st;
  if (rc == DGB_ERR_CHECKSUM) {
    st.error = mismatchErr;
    for (size_t i = 0; i < mismatch.size(); ++i)
      if (mismatch[i]) st.errorInfo.emplace_back((int)i, "Checksum mismatch in batch member " + std::to_string(i) + "\n");
  }
  return st;
}
inline uint32_t maxOf(const uint32_t* v, uint32_t n) {
  uint32_t m = 0;
  for (uint32_t i = 0; i < n; ++i) m = std::max(m, v[i]);
  return m;
}
}  // namespace detail

// ---- encode (ans/GpuANSCodec.h:65-164) ----
inline void ansEncodeBatchStride(StackDeviceMemory& res, const ANSCodecConfig& config, uint32_t numInBatch,
                                 const void* in_dev, uint32_t inPerBatchSize, uint32_t inPerBatchStride,
                                 const uint32_t* histogram_dev, void* out_dev, uint32_t outPerBatchStride,
                                 uint32_t* outBatchSize_dev, cudaStream_t stream) {
  auto l = res.lease(dgb_ans_encode_temp_bytes(numInBatch, inPerBatchSize), stream);
  detail::check("ansEncodeBatchStride",
                dgb_ans_encode_stride(l.ptr, l.bytes, config.probBits, config.useChecksum, numInBatch, in_dev,
                                      inPerBatchSize, inPerBatchStride, histogram_dev, out_dev, outPerBatchStride,
                                      outBatchSize_dev, stream));
}

inline void ansEncodeBatchPointer(StackDeviceMemory& res, const ANSCodecConfig& config, uint32_t numInBatch,
                                  const void** in, const uint32_t* inSize, const uint32_t* histogram_dev,
                                  void** out, uint32_t* outSize_dev, cudaStream_t stream) {
  auto l = res.lease(dgb_ans_encode_temp_bytes(numInBatch, detail::maxOf(inSize, numInBatch)), stream);
  detail::check("ansEncodeBatchPointer",
                dgb_ans_encode_pointer(l.ptr, l.bytes, config.probBits, config.useChecksum, numInBatch, in, inSize,
                                       histogram_dev, out, outSize_dev, stream));
}

inline void ansEncodeBatchSplitSize(StackDeviceMemory& res, const ANSCodecConfig& config, uint32_t numInBatch,
                                    const void* in_dev, const uint32_t* inSplitSizes, const uint32_t* histogram_dev,
                                    void* out_dev, uint32_t outStride, uint32_t* outSize_dev, cudaStream_t stream) {
  auto l = res.lease(dgb_ans_encode_temp_bytes(numInBatch, detail::maxOf(inSplitSizes, numInBatch)), stream);
  detail::check("ansEncodeBatchSplitSize",
                dgb_ans_encode_split_size(l.ptr, l.bytes, config.probBits, config.useChecksum, numInBatch, in_dev,
                                          inSplitSizes, histogram_dev, out_dev, outStride, outSize_dev, stream));
}

// ---- decode (ans/GpuANSCodec.h:170-303) ----
inline ANSDecodeStatus ansDecodeBatchStride(StackDeviceMemory& res, const ANSCodecConfig& config, uint32_t numInBatch,
                                            const void* in_dev, uint32_t inPerBatchStride, void* out_dev,
                                            uint32_t outPerBatchStride, uint32_t outPerBatchCapacity,
                                            uint8_t* outSuccess_dev, uint32_t* outSize_dev, cudaStream_t stream) {
  auto l = res.lease(dgb_ans_decode_temp_bytes(numInBatch), stream);
  std::vector<uint8_t> mm(numInBatch);
  int rc = dgb_ans_decode_stride(l.ptr, l.bytes, config.probBits, config.useChecksum, numInBatch, in_dev,
                                 inPerBatchStride, out_dev, outPerBatchStride, outPerBatchCapacity, outSuccess_dev,
                                 outSize_dev, mm.data(), stream);
  detail::check("ansDecodeBatchStride", rc);
  return detail::makeStatus<ANSDecodeStatus>(rc, mm, ANSDecodeError::ChecksumMismatch);
}

inline ANSDecodeStatus ansDecodeBatchPointer(StackDeviceMemory& res, const ANSCodecConfig& config, uint32_t numInBatch,
                                             const void** in, void** out, const uint32_t* outCapacity,
                                             uint8_t* outSuccess_dev, uint32_t* outSize_dev, cudaStream_t stream) {
  auto l = res.lease(dgb_ans_decode_temp_bytes(numInBatch), stream);
  std::vector<uint8_t> mm(numInBatch);
  int rc = dgb_ans_decode_pointer(l.ptr, l.bytes, config.probBits, config.useChecksum, numInBatch, in, out,
                                  outCapacity, outSuccess_dev, outSize_dev, mm.data(), stream);
  detail::check("ansDecodeBatchPointer", rc);
  return detail::makeStatus<ANSDecodeStatus>(rc, mm, ANSDecodeError::ChecksumMismatch);
}

inline ANSDecodeStatus ansDecodeBatchSplitSize(StackDeviceMemory& res, const ANSCodecConfig& config,
                                               uint32_t numInBatch, const void** in, void* out_dev,
                                               const uint32_t* outSplitSizes, uint8_t* outSuccess_dev,
                                               uint32_t* outSize_dev, cudaStream_t stream) {
  auto l = res.lease(dgb_ans_decode_temp_bytes(numInBatch), stream);
  std::vector<uint8_t> mm(numInBatch);
  int rc = dgb_ans_decode_split_size(l.ptr, l.bytes, config.probBits, config.useChecksum, numInBatch, in, out_dev,
                                     outSplitSizes, outSuccess_dev, outSize_dev, mm.data(), stream);
  detail::check("ansDecodeBatchSplitSize", rc);
  return detail::makeStatus<ANSDecodeStatus>(rc, mm, ANSDecodeError::ChecksumMismatch);
}

// ---- information (ans/GpuANSCodec.h:309-341) ----
inline void ansGetCompressedInfo(StackDeviceMemory& res, const void** in, uint32_t numInBatch,
                                 uint32_t* outSizes_dev, uint32_t* outChecksum_dev, cudaStream_t stream) {
  auto l = res.lease(sizeof(void*) * (size_t)numInBatch + 256, stream);
  detail::check("ansGetCompressedInfo",
                dgb_ans_get_compressed_info(l.ptr, l.bytes, in, 0, numInBatch, outSizes_dev, outChecksum_dev, stream));
}
inline void ansGetCompressedInfoDevice(StackDeviceMemory&, const void** in_dev, uint32_t numInBatch,
                                       uint32_t* outSizes_dev, uint32_t* outChecksum_dev, cudaStream_t stream) {
  detail::check("ansGetCompressedInfoDevice",
                dgb_ans_get_compressed_info(nullptr, 0, in_dev, 1, numInBatch, outSizes_dev, outChecksum_dev, stream));
}

// ---- float codec (float/GpuFloatCodec.h:103-292) ----
inline void floatCompress(StackDeviceMemory& res, const FloatCompressConfig& config, uint32_t numInBatch,
                          const void** in, const uint32_t* inSize, void** out, uint32_t* outSize_dev,
                          cudaStream_t stream) {
  auto l = res.lease(dgb_float_compress_temp_bytes((int)config.floatType, numInBatch, detail::maxOf(inSize, numInBatch)), stream);
  detail::check("floatCompress",
                dgb_float_compress_pointer(l.ptr, l.bytes, (int)config.floatType, config.ansConfig.probBits,
                                           config.useChecksum, numInBatch, in, inSize, out, outSize_dev, stream));
}

inline void floatCompressSplitSize(StackDeviceMemory& res, const FloatCompressConfig& config, uint32_t numInBatch,
                                   const void* in_dev, const uint32_t* inSplitSizes, void* out_dev,
                                   uint32_t outStride, uint32_t* outSize_dev, cudaStream_t stream) {
  auto l = res.lease(dgb_float_compress_temp_bytes((int)config.floatType, numInBatch, detail::maxOf(inSplitSizes, numInBatch)), stream);
  detail::check("floatCompressSplitSize",
                dgb_float_compress_split_size(l.ptr, l.bytes, (int)config.floatType, config.ansConfig.probBits,
                                              config.useChecksum, numInBatch, in_dev, inSplitSizes, out_dev, outStride,
                                              outSize_dev, stream));
}

inline FloatDecompressStatus floatDecompress(StackDeviceMemory& res, const FloatDecompressConfig& config,
                                             uint32_t numInBatch, const void** in, void** out,
                                             const uint32_t* outCapacity, uint8_t* outSuccess_dev,
                                             uint32_t* outSize_dev, cudaStream_t stream) {
  auto l = res.lease(dgb_float_decompress_temp_bytes((int)config.floatType, numInBatch, 0), stream);
  std::vector<uint8_t> mm(numInBatch);
  int rc = dgb_float_decompress_pointer(l.ptr, l.bytes, (int)config.floatType, config.ansConfig.probBits,
                                        config.useChecksum, numInBatch, in, out, outCapacity, outSuccess_dev,
                                        outSize_dev, mm.data(), stream);
  detail::check("floatDecompress", rc);
  return detail::makeStatus<FloatDecompressStatus>(rc, mm, FloatDecompressError::ChecksumMismatch);
}

inline FloatDecompressStatus floatDecompressSplitSize(StackDeviceMemory& res, const FloatDecompressConfig& config,
                                                      uint32_t numInBatch, const void** in, void* out_dev,
                                                      const uint32_t* outSplitSizes, uint8_t* outSuccess_dev,
                                                      uint32_t* outSize_dev, cudaStream_t stream) {
  auto l = res.lease(dgb_float_decompress_temp_bytes((int)config.floatType, numInBatch, 0), stream);
  std::vector<uint8_t> mm(numInBatch);
  int rc = dgb_float_decompress_split_size(l.ptr, l.bytes, (int)config.floatType, config.ansConfig.probBits,
                                           config.useChecksum, numInBatch, in, out_dev, outSplitSizes,
                                           outSuccess_dev, outSize_dev, mm.data(), stream);
  detail::check("floatDecompressSplitSize", rc);
  return detail::makeStatus<FloatDecompressStatus>(rc, mm, FloatDecompressError::ChecksumMismatch);
}

inline void floatGetCompressedInfo(StackDeviceMemory& res, const void** in, uint32_t numInBatch,
                                   uint32_t* outSizes_dev, uint32_t* outTypes_dev, uint32_t* outChecksum_dev,
                                   cudaStream_t stream) {
  auto l = res.lease(sizeof(void*) * (size_t)numInBatch + 256, stream);
  detail::check("floatGetCompressedInfo",
                dgb_float_get_compressed_info(l.ptr, l.bytes, in, 0, numInBatch, outSizes_dev, outTypes_dev,
                                              outChecksum_dev, stream));
}
inline void floatGetCompressedInfoDevice(StackDeviceMemory&, const void** in_dev, uint32_t numInBatch,
                                         uint32_t* outSizes_dev, uint32_t* outTypes_dev,
                                         uint32_t* outChecksum_dev, cudaStream_t stream) {
  detail::check("floatGetCompressedInfoDevice",
                dgb_float_get_compressed_info(nullptr, 0, in_dev, 1, numInBatch, outSizes_dev, outTypes_dev,
                                              outChecksum_dev, stream));
}

}  // namespace dietgpu
