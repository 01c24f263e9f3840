// dietgpu_b200_device.cuh -- warp / CTA level device API of the B200-native rANS codec.
//
// The reference lists "CUB-like APIs for fusing warp-oriented ANS into user kernels" as a planned
// extension (README.md:105); it exports nothing at device level.  This header is that API: a user
// kernel can histogram and normalise its own data, entropy-code 4 KiB blocks with one warp each
// straight out of (and into) shared memory, and write archives the library (and the reference)
// decodes -- without a round trip through global memory and without a separate codec launch.
//
// Wire semantics are the library's and the reference's:
//   * blockNormalizedPdf      ans/GpuANSStatistics.cuh:178-341  (bit-identical, incl. the symbol-id quirk)
//   * warpEncodeBlock         ans/GpuANSEncode.cuh:49-211       (32 interleaved lane states, emit order = lane order)
//   * blockBuildDecodeLut     ans/GpuANSDecode.cuh:405-476
//   * warpDecodeBlock         ans/GpuANSDecode.cuh:55-217
//   * blockWriteArchive       ans/GpuANSEncode.cuh:515-628 / ans/GpuANSUtils.cuh:67-227 (archive layout)
// The library's own kernels (dietgpu_b200/csrc/encode.cu, decode.cu) use the same normalisation
// routine from this header and tuned variants of the two block coders (cp.async / TMA fed, PTX emit
// blocks); the versions here take plain pointers so that they compose with any user kernel.
//
// Header-only, sm_100a (any sm_70+ compiles).  Include from a .cu file; no host code, no linking.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace dietgpu_b200 {
namespace device {

constexpr uint32_t kBlockBytes = 4096;  // ans/GpuANSUtils.cuh:37
constexpr uint32_t kNumSymbols = 256;
constexpr uint32_t kStateMin = 1u << 15;                      // ans/GpuANSUtils.cuh:46-49
constexpr uint32_t kAnsMagicVersion = (0xd00du << 16) | 1u;   // ans/GpuANSUtils.cuh:52-55,105-107
constexpr uint32_t kAnsHeaderBytes = 32, kAnsPdfBytes = 512;

// worst-case u16 words of one block: every symbol costs at most probBits bits; +8 for 16 B padding
__host__ __device__ constexpr uint32_t maxBlockWords(int probBits) { return 256u * (uint32_t)probBits + 8u; }
// archive bytes that are not block streams (ans/GpuANSUtils.cuh:68-81)
__host__ __device__ constexpr uint32_t archiveOverhead(uint32_t numBlocks) {
  return kAnsHeaderBytes + kAnsPdfBytes + 128u * numBlocks + 8u * ((numBlocks + 1u) / 2u * 2u);
}

// ---- statistics ------------------------------------------------------------------------------
// CTA of exactly 256 threads, thread t <-> symbol t.  count = occurrences of symbol t, total = sum
// of all counts (> 0).  Returns this symbol's pdf (sums to 2^probBits over the CTA) and writes the
// exclusive prefix (cdf).  Bit-identical to the reference's normalizeProbabilitiesFromHistogram:
// fp32 quantisation (IEEE divide, truncation), descending sort on (q << 16 | symbol) -- done here
// by counting ranks --, then the two fix-up branches, of which the "add" branch indexes by SYMBOL ID
// (ans/GpuANSStatistics.cuh:258-273).  Contains CTA barriers: call it uniformly.
__device__ inline uint32_t blockNormalizedPdf(uint32_t count, uint32_t total, int probBits, uint32_t* cdfOut) {
  __shared__ uint32_t sKey[kNumSymbols];
  __shared__ uint32_t sQByRank[kNumSymbols];
  __shared__ uint32_t sSymByRank[kNumSymbols];
  __shared__ uint32_t sPdf[kNumSymbols];
  __shared__ uint32_t sWarp[8];

  const uint32_t t = threadIdx.x;
  const uint32_t K = 1u << probBits;
  // :215-218 fp32 quantisation, IEEE divide, truncation
  float ratio = __fdiv_rn(__uint2float_rn(count), __uint2float_rn(total));
  uint32_t q = __float2uint_rz(__fmul_rn((float)K, ratio));
  if (count > 0 && q == 0) q = 1;

  uint32_t incl = q;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) incl += __shfl_xor_sync(0xffffffffu, incl, d);
  __syncthreads();  // the static arrays may still be read by a previous call
  if ((t & 31) == 0) sWarp[t >> 5] = incl;
  const uint32_t key = (q << 16) | t;
  sKey[t] = key;
  __syncthreads();
  int sum = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += (int)sWarp[w];

  // rank = number of keys strictly greater (descending order, keys unique)
  uint32_t rank = 0;
#pragma unroll 8
  for (int j = 0; j < (int)kNumSymbols; ++j) rank += (sKey[j] > key);
  sQByRank[rank] = q;
  sSymByRank[rank] = t;
  __syncthreads();

  // from here thread t owns RANK t
  uint32_t qr = sQByRank[t];
  const uint32_t symr = sSymByRank[t];
  int diff = (int)K - sum;
  if (diff > 0) {
    // :258-273: +1 to every entry whose SYMBOL ID < min(diff, 256), repeated
    while (diff > 0) {
      int it = diff < (int)kNumSymbols ? diff : (int)kNumSymbols;
      if ((int)symr < it) qr += 1;
      diff -= it;
    }
  } else if (diff < 0) {
    // :274-315: -1 from the smallest entries still > 1, by rank, iterated
    diff = -diff;
    while (diff > 0) {
      int g = __syncthreads_count(qr > 1);
      int it = diff < g ? diff : g;
      if (it <= 0) break;
      if ((int)t >= g - it && (int)t < g) qr -= 1;
      diff -= it;
    }
  }
  sPdf[symr] = qr;
  __syncthreads();

  // back to thread t == symbol t; exclusive scan -> cdf (:336-341)
  const uint32_t pdf = sPdf[t];
  uint32_t inc = pdf;
  const uint32_t lane = t & 31u, warp = t >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t v = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= (uint32_t)d) inc += v;
  }
  __syncthreads();
  if (lane == 31) sWarp[warp] = inc;
  __syncthreads();
  uint32_t base = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w)
    if ((uint32_t)w < warp) base += sWarp[w];
  *cdfOut = base + inc - pdf;
  return pdf;
}

// ---- encoder table -----------------------------------------------------------------------------
// One entry per symbol (8 B, shared memory).  state / pdf == hi32(state * magic) >> shift for every
// coder state < 2^31 with magic = ceil(2^(32+shift) / pdf), shift = ceil(log2 pdf) - 1 (the quotient
// the reference's round-up magic computes, ans/GpuANSStatistics.cuh:343-358, without its add);
// pdf == 1 uses magic = 2^32 - 1 (quotient state - 1) and carries the missing 2^probBits - 1 in the
// cdf term.  pack = shift | (2^probBits - pdf) << 5 | cdf term << 20.
struct EncodeEntry {
  uint32_t magic, pack;
};
struct EncodeTable {
  EncodeEntry e[kNumSymbols];
};

__device__ inline EncodeEntry makeEncodeEntry(uint32_t pdf, uint32_t cdf, int probBits) {
  const uint32_t K = 1u << probBits;
  uint32_t shift = 0, magic = 0, cdfTerm = cdf;
  if (pdf > 1) {
    shift = 31u - (uint32_t)__clz((int)(pdf - 1));
    magic = (uint32_t)(((1ull << (32 + shift)) + pdf - 1) / pdf);
  } else if (pdf == 1) {
    magic = 0xffffffffu;
    cdfTerm = cdf + (K - 1u);
  }
  EncodeEntry e;
  e.magic = magic;
  e.pack = shift | ((K - pdf) << 5) | (cdfTerm << 20);
  return e;
}

// CTA of 256 threads: counts (shared or global, 256 words) -> table (shared) and the u16 pdf an archive
// stores (shared or global, 256 entries).  total must equal the sum of counts and be > 0.
__device__ inline void blockBuildEncodeTable(const uint32_t* counts, uint32_t total, int probBits, EncodeTable* table,
                                             uint16_t* pdfOut) {
  uint32_t cdf;
  const uint32_t pdf = blockNormalizedPdf(counts[threadIdx.x], total, probBits, &cdf);
  table->e[threadIdx.x] = makeEncodeEntry(pdf, cdf, probBits);
  pdfOut[threadIdx.x] = (uint16_t)pdf;
  __syncthreads();
}

// ---- block coder -------------------------------------------------------------------------------
__device__ inline uint32_t laneMaskLt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
__device__ inline uint32_t laneMaskGe() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_ge;" : "=r"(m));
  return m;
}

// One warp encodes bytes in[0, n), 1 <= n <= 4096 (row r, lane l <-> byte 32 r + l), into outWords
// (room for maxBlockWords(probBits)); returns the number of u16 words written.  *laneStateOut is this
// lane's final coder state (the 32 of them are the block's ANSWarpState).  `in` and `outWords` may be
// shared or global memory; all 32 lanes must call.
__device__ inline uint32_t warpEncodeBlock(const uint8_t* in, uint32_t n, const EncodeTable* table, int probBits,
                                           uint16_t* outWords, uint32_t* laneStateOut) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lt = laneMaskLt();
  const uint32_t thrNegScale = 0u - (1u << (31 - probBits));
  uint32_t state = kStateMin, pos = 0;
  for (uint32_t r = 0; r < n; r += 32) {
    const bool valid = r + lane < n;
    const EncodeEntry e = table->e[valid ? in[r + lane] : 0];
    const uint32_t kmp = (e.pack >> 5) & 0xfffu;            // 2^probBits - pdf
    const uint32_t thr = kmp * thrNegScale + 0x80000000u;   // pdf << (31 - probBits)
    // renormalise: lanes whose state is too large emit their low 16 bits, in lane order
    const bool emit = valid && state >= thr;
    const uint32_t vote = __ballot_sync(0xffffffffu, emit);
    if (emit) {
      outWords[pos + __popc(vote & lt)] = (uint16_t)state;
      state >>= 16;
    }
    pos += __popc(vote);
    // x' = (x / pdf) * 2^probBits + x % pdf + cdf  ==  (x / pdf) * (2^probBits - pdf) + x + cdf
    if (valid) {
      const uint32_t div = __umulhi(state, e.magic) >> (e.pack & 31u);
      state = div * kmp + state + (e.pack >> 20);
    }
  }
  *laneStateOut = state;
  return pos;
}

// ---- decoder -----------------------------------------------------------------------------------
// LUT entry per state residue: [31:20] pdf  [19:8] residue - cdf  [7:0] symbol.
// CTA of THREADS threads (multiple of 32, >= 32): u16 pdf[256] (shared or global) -> lut[2^PB] (shared).
// Returns false (uniformly) if the pdf does not sum to 2^PB.
template <int PB, int THREADS>
__device__ inline bool blockBuildDecodeLut(const uint16_t* pdf, uint32_t* lut) {
  __shared__ uint32_t sCdfStart[kNumSymbols + 1];
  const uint32_t t = threadIdx.x;
  __syncthreads();
  if (t < 32) {
    // 8 symbols per lane, warp scan
    uint32_t p[8], s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { p[k] = pdf[t * 8 + k]; s += p[k]; }
    uint32_t inc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t v = __shfl_up_sync(0xffffffffu, inc, d);
      if (t >= (uint32_t)d) inc += v;
    }
    uint32_t c = inc - s;
#pragma unroll
    for (int k = 0; k < 8; ++k) { sCdfStart[t * 8 + k] = c; c += p[k]; }
    if (t == 31) sCdfStart[kNumSymbols] = c;
  }
  __syncthreads();
  const bool ok = sCdfStart[kNumSymbols] == (1u << PB);
  if (ok) {
    for (uint32_t sym = t >> 5; sym < kNumSymbols; sym += THREADS / 32) {
      const uint32_t begin = sCdfStart[sym], p = sCdfStart[sym + 1] - begin;
      for (uint32_t j = t & 31u; j < p; j += 32u) lut[begin + j] = (p << 20) | (j << 8) | sym;
    }
  }
  __syncthreads();
  return ok;
}

// One warp decodes a block of n bytes (1 <= n <= 4096): laneState = this lane's stored state,
// words[0, numWords) = the block's stream (shared or global), lut = blockBuildDecodeLut's table.
// Writes out[0, n) (shared or global).  Returns false if the stream is inconsistent (wrong word
// count or final states), which the reference does not check.
template <int PB>
__device__ inline bool warpDecodeBlock(uint32_t laneState, const uint16_t* words, uint32_t numWords, uint32_t n,
                                       const uint32_t* lut, uint8_t* out) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t ge = laneMaskGe();
  uint32_t state = laneState, pos = numWords;
  bool ok = true;
  const uint32_t lastRow = (n - 1u) / 32u * 32u;
  for (int64_t r = lastRow; r >= 0; r -= 32) {
    const bool valid = (uint32_t)r + lane < n;
    if (valid) {
      const uint32_t e = lut[state & ((1u << PB) - 1u)];
      out[(uint32_t)r + lane] = (uint8_t)e;
      state = (e >> 20) * (state >> PB) + ((e >> 8) & 0xfffu);
    }
    // refill: lanes whose state dropped below 2^15 pop one word each, highest lane first
    const bool rd = valid && state < kStateMin;
    const uint32_t vote = __ballot_sync(0xffffffffu, rd);
    const uint32_t cnt = (uint32_t)__popc(vote);
    if (cnt > pos) { ok = false; break; }
    if (rd) state = (state << 16) + words[pos - (uint32_t)__popc(vote & ge)];
    pos -= cnt;
  }
  ok = ok && pos == 0 && state == kStateMin;
  return __all_sync(0xffffffffu, ok);
}

// ---- archive -----------------------------------------------------------------------------------
// Pointers into an archive of numBlocks blocks at `archive` (16 B aligned).
struct ArchiveLayout {
  uint8_t* header;       // 32 B
  uint16_t* pdf;         // 256 x u16
  uint32_t* states;      // numBlocks x 32 lane states
  uint2* blockWords;     // numBlocks x {uncompressed << 16 | words, word offset}
  uint16_t* data;        // block streams, each padded to 8 words
};
__device__ inline ArchiveLayout archiveLayout(uint8_t* archive, uint32_t numBlocks) {
  ArchiveLayout a;
  a.header = archive;
  a.pdf = reinterpret_cast<uint16_t*>(archive + kAnsHeaderBytes);
  a.states = reinterpret_cast<uint32_t*>(archive + kAnsHeaderBytes + kAnsPdfBytes);
  a.blockWords = reinterpret_cast<uint2*>(archive + kAnsHeaderBytes + kAnsPdfBytes + 128u * numBlocks);
  a.data = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(a.blockWords) + 8u * ((numBlocks + 1u) / 2u * 2u));
  return a;
}

// One thread writes the 32 B header (undefined bits of the reference's header are zero here).
__device__ inline void writeArchiveHeader(uint8_t* archive, uint32_t numBlocks, uint32_t uncompressedBytes,
                                          uint32_t totalWords, int probBits) {
  uint4* h = reinterpret_cast<uint4*>(archive);
  h[0] = make_uint4(kAnsMagicVersion, numBlocks, uncompressedBytes, totalWords);
  h[1] = make_uint4((uint32_t)probBits, 0u, 0u, 0u);
}

}  // namespace device
}  // namespace dietgpu_b200
