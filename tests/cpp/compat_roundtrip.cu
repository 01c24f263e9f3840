// Round-trip tests through the C++ compatibility layer (include/dietgpu_b200_compat.hpp), shaped
// like the reference's own gtests so the two can be read side by side:
//   ANSTest.{ZeroSized,BatchPointer,BatchPointerLarge,BatchStride}   (dietgpu/ans/ANSTest.cu:243-282)
//   FloatTest.{Batch,LargeBatch,BatchSize1}                          (dietgpu/float/FloatTest.cu:270-311)
// Same generators (std::mt19937(10) + exponential_distribution for bytes, mt19937(10+n) +
// normal_distribution for floats), same size lists, checksum on.  Plain asserts instead of gtest.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "../../include/dietgpu_b200_compat.hpp"

using namespace dietgpu;

#define REQUIRE(c)                                                        \
  do {                                                                    \
    if (!(c)) {                                                           \
      std::fprintf(stderr, "%s:%d REQUIRE(%s) failed\n", __FILE__, __LINE__, #c); \
      std::exit(1);                                                       \
    }                                                                     \
  } while (0)

template <typename T>
struct Dev {
  T* p = nullptr;
  size_t n = 0;
  explicit Dev(size_t count) : n(count) { REQUIRE(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)) == cudaSuccess); }
  Dev(const std::vector<T>& h) : Dev(h.size()) { if (n) cudaMemcpy(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice); }
  ~Dev() { cudaFree(p); }
  std::vector<T> host() const { std::vector<T> h(n); if (n) cudaMemcpy(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost); return h; }
};

static std::vector<uint8_t> generateSymbols(int num, float lambda) {
  std::mt19937 gen(10 + num);
  std::exponential_distribution<float> dist(lambda);
  std::vector<uint8_t> out(num);
  for (auto& v : out) v = (uint8_t)(std::min(dist(gen), 1.0f) * 255.0f);
  return out;
}

static void runAnsBatchPointer(StackDeviceMemory& res, int prec, const std::vector<uint32_t>& sizes, float lambda) {
  const uint32_t n = (uint32_t)sizes.size();
  uint32_t maxSize = 0;
  for (auto s : sizes) maxSize = std::max(maxSize, s);
  const uint32_t outStride = getMaxCompressedSize(maxSize);
  std::vector<std::vector<uint8_t>> orig;
  std::vector<Dev<uint8_t>*> in, dec;
  Dev<uint8_t> enc((size_t)n * outStride);
  std::vector<const void*> inPtr(n), encPtrC(n);
  std::vector<void*> encPtr(n), decPtr(n);
  for (uint32_t i = 0; i < n; ++i) {
    orig.push_back(generateSymbols(sizes[i], lambda));
    in.push_back(new Dev<uint8_t>(orig.back()));
    dec.push_back(new Dev<uint8_t>(sizes[i]));
    inPtr[i] = in[i]->p;
    encPtr[i] = enc.p + (size_t)i * outStride;
    encPtrC[i] = encPtr[i];
    decPtr[i] = dec[i]->p;
  }
  Dev<uint32_t> encSize(n), decSize(n);
  Dev<uint8_t> ok(n);
  ANSCodecConfig cfg(prec, true);
  ansEncodeBatchPointer(res, cfg, n, inPtr.data(), sizes.data(), nullptr, encPtr.data(), encSize.p, 0);
  auto es = encSize.host();
  for (auto s : es) REQUIRE(s % 16 == 0 && s <= outStride);  // ANSTest.cu:131-135
  auto st = ansDecodeBatchPointer(res, cfg, n, encPtrC.data(), decPtr.data(), sizes.data(), ok.p, decSize.p, 0);
  REQUIRE(st.error == ANSDecodeError::None);
  auto oks = ok.host();
  auto ds = decSize.host();
  for (uint32_t i = 0; i < n; ++i) {
    REQUIRE(oks[i] == 1 && ds[i] == sizes[i]);
    REQUIRE(dec[i]->host() == orig[i]);
    delete in[i];
    delete dec[i];
  }
}

static uint16_t toBf16(float f) { uint32_t u; std::memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }

template <typename W>
static std::vector<W> generateFloats(int num, FloatType ft) {
  std::mt19937 gen(10 + num);
  std::normal_distribution<float> dist;
  std::vector<W> out(num);
  for (auto& v : out) {
    float f = dist(gen);
    if (ft == FloatType::kFloat32) { uint32_t u; std::memcpy(&u, &f, 4); v = (W)u; }
    else if (ft == FloatType::kBFloat16) v = (W)toBf16(f);
    else v = (W)(toBf16(f * 0.25f) ^ 0x0101);  // any 16-bit pattern exercises the fp16 split
  }
  return out;
}

template <typename W>
static void runFloatBatch(StackDeviceMemory& res, FloatType ft, int prec, const std::vector<uint32_t>& sizes, bool misalign) {
  const uint32_t n = (uint32_t)sizes.size();
  uint32_t maxSize = 0;
  for (auto s : sizes) maxSize = std::max(maxSize, s);
  const uint32_t outStride = getMaxFloatCompressedSize(ft, maxSize);
  std::vector<std::vector<W>> orig;
  std::vector<Dev<W>*> in, dec;
  Dev<uint8_t> enc((size_t)n * outStride);
  std::vector<const void*> inPtr(n), encPtrC(n);
  std::vector<void*> encPtr(n), decPtr(n);
  const size_t off = misalign ? 1 : 0;  // FloatTest.cu:276-282: word-aligned but not 16 B aligned
  for (uint32_t i = 0; i < n; ++i) {
    orig.push_back(generateFloats<W>(sizes[i], ft));
    in.push_back(new Dev<W>(sizes[i] + off));
    dec.push_back(new Dev<W>(sizes[i] + off));
    if (sizes[i]) cudaMemcpy(in[i]->p + off, orig[i].data(), sizes[i] * sizeof(W), cudaMemcpyHostToDevice);
    inPtr[i] = in[i]->p + off;
    encPtr[i] = enc.p + (size_t)i * outStride;
    encPtrC[i] = encPtr[i];
    decPtr[i] = dec[i]->p + off;
  }
  Dev<uint32_t> encSize(n), decSize(n);
  Dev<uint8_t> ok(n);
  FloatCodecConfig cfg(ft, ANSCodecConfig(prec, false), !misalign, true);
  floatCompress(res, cfg, n, inPtr.data(), sizes.data(), encPtr.data(), encSize.p, 0);
  auto st = floatDecompress(res, cfg, n, encPtrC.data(), decPtr.data(), sizes.data(), ok.p, decSize.p, 0);
  REQUIRE(st.error == FloatDecompressError::None);
  auto oks = ok.host();
  auto ds = decSize.host();
  for (uint32_t i = 0; i < n; ++i) {
    REQUIRE(oks[i] == 1 && ds[i] == sizes[i]);
    std::vector<W> got(sizes[i]);
    if (sizes[i]) cudaMemcpy(got.data(), dec[i]->p + off, sizes[i] * sizeof(W), cudaMemcpyDeviceToHost);
    REQUIRE(got == orig[i]);
    delete in[i];
    delete dec[i];
  }
}

int main() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    std::fprintf(stderr, "no CUDA device\n");
    return 2;
  }
  auto res = makeStackMemory(64 << 20);
  // ANSTest.ZeroSized / BatchPointer
  runAnsBatchPointer(res, 10, {0}, 10.0f);
  for (int prec : {9, 10, 11})
    for (float lambda : {1.0f, 10.0f, 100.0f, 1000.0f})
      for (auto sizes : std::vector<std::vector<uint32_t>>{{1}, {1, 1}, {4096, 4095, 4096}, {1234, 2345, 3456}, {10000, 10013, 10000}})
        runAnsBatchPointer(res, prec, sizes, lambda);
  // ANSTest.BatchPointerLarge
  {
    std::mt19937 gen(10);
    std::uniform_int_distribution<uint32_t> dist(100, 10000);
    std::vector<uint32_t> sizes(100);
    for (auto& s : sizes) s = dist(gen);
    runAnsBatchPointer(res, 10, sizes, 20.0f);
  }
  // ANSTest.BatchStride: 13 x 8208 through the stride API
  {
    const uint32_t n = 13, sz = 8208, ostride = getMaxCompressedSize(sz);
    std::vector<uint8_t> all;
    for (uint32_t i = 0; i < n; ++i) { auto v = generateSymbols(sz + i, 25.0f); all.insert(all.end(), v.begin(), v.begin() + sz); }
    Dev<uint8_t> in(all), enc((size_t)n * ostride), dec((size_t)n * sz), ok(n);
    Dev<uint32_t> es(n), dsz(n);
    ANSCodecConfig cfg(10, true);
    ansEncodeBatchStride(res, cfg, n, in.p, sz, sz, nullptr, enc.p, ostride, es.p, 0);
    auto st = ansDecodeBatchStride(res, cfg, n, enc.p, ostride, dec.p, sz, sz, ok.p, dsz.p, 0);
    REQUIRE(st.error == ANSDecodeError::None);
    REQUIRE(dec.host() == all);
  }
  // FloatTest.Batch / LargeBatch / BatchSize1
  for (int prec : {9, 10})
    for (bool mis : {false, true}) {
      for (auto sizes : std::vector<std::vector<uint32_t>>{{1}, {13, 4096, 9999}, {512 * 1024}}) {
        runFloatBatch<uint16_t>(res, FloatType::kFloat16, prec, sizes, mis);
        runFloatBatch<uint16_t>(res, FloatType::kBFloat16, prec, sizes, mis);
        runFloatBatch<uint32_t>(res, FloatType::kFloat32, prec, sizes, mis);
      }
    }
  {
    std::vector<uint32_t> sizes(256);
    for (uint32_t i = 0; i < 256; ++i) sizes[i] = 1000 + 37 * i;
    runFloatBatch<uint16_t>(res, FloatType::kBFloat16, 10, sizes, false);
  }
  // a too-small scratch region must fall back to cudaMalloc, not fail (utils/StackDeviceMemory.cpp:119-139)
  {
    StackDeviceMemory tiny(0, nullptr, 0);
    runAnsBatchPointer(tiny, 10, {5000, 123}, 10.0f);
    REQUIRE(tiny.getMaxMemoryUsage() > 0);
  }
  std::printf("compat_roundtrip: all passed\n");
  return 0;
}
