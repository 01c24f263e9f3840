// decode.cu -- batched rANS decode (bytes) and fused float decompress
// (fp16/bf16/fp32) for sm_100a.  Two launches per call:
//
//   P   planKernel     reads every member's header(s), validates them,
//                      writes outSuccess / outSize (ans/GpuANSDecode.cuh:
//                      326-341 semantics) and builds the flat block index
//                      (exclusive scan of block counts) the decode kernel
//                      partitions.  The reference instead sizes its grid blind
//                      (ans/GpuANSDecode.cuh:500-523) and runs a separate LUT
//                      kernel through global memory (:405-476).
//   D   decodeKernel   each CTA owns a contiguous run of 4 KiB blocks; it
//                      builds the 2^pb-entry decode LUT of the member in
//                      shared memory straight from the archive's pdf, then
//                      every warp decodes whole blocks (ans/GpuANSDecode.cuh:
//                      55-217, 274-297 restated).  The block's compressed
//                      words and lane states are staged into shared memory
//                      with one TMA bulk copy (cp.async.bulk + mbarrier) so
//                      the loop-carried refill load is an LDS, not a global
//                      load; float kinds join the decoded byte with the stored
//                      byte(s) and write the float word directly (any
//                      alignment -- the reference's two-pass path,
//                      float/GpuFloatDecompress.cuh:622-694, is not needed).
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace dgb {

namespace {

struct DecodeScratch {
  MemberDesc* members;   // [n]; planKernel fills work0 = first flat block
  uint32_t* totals;      // [0] = total blocks, [1] = unused
  uint32_t* checksum;    // [n] checksum of decoded output (use_checksum only; zeroed)
  uint32_t* archiveChecksum;  // [n]
  uint32_t* sizes;       // [n] decoded size in bytes (for the checksum pass)
};

// Where the ANS archive of member `md` starts, and the float-level fields.
struct ArchiveView {
  const uint8_t* ans;      // ANS archive base
  const uint8_t* non;      // float kinds: stored (non-compressed) plane(s)
  uint32_t floatWords;     // float kinds: size field of the float header
  bool ok;
};

template <int KIND>
__device__ __forceinline__ ArchiveView openArchive(const uint8_t* in) {
  ArchiveView v;
  if (KIND == kKindBytes) {
    v.ans = in;
    v.non = nullptr;
    v.floatWords = 0;
    v.ok = true;
  } else {
    // float/GpuFloatUtils.cuh:26-74 GpuFloatHeader
    const uint4 fh = __ldg(reinterpret_cast<const uint4*>(in));
    v.ok = fh.x == kFloatMagicVersion && (int)(fh.z & 0xfu) == KIND;
    v.floatWords = fh.y;
    v.non = in + kFloatHeaderBytes;
    v.ans = v.non + floatNonCompBytes(KIND, fh.y);
  }
  return v;
}

// ---------------------------------------------------------------------------
// P: plan.  One CTA; member i handled by thread (i % blockDim) in rounds.
// ---------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(1024)
planKernel(DecodeScratch sc, uint32_t n, int pb, uint8_t* __restrict__ outSuccess,
           uint32_t* __restrict__ outSize, bool wantChecksum) {
  __shared__ uint32_t sWarp[32];
  uint32_t carry = 0;
  for (uint32_t base = 0; base < n; base += blockDim.x) {
    const uint32_t i = base + threadIdx.x;
    uint32_t blocks = 0;
    if (i < n) {
      const MemberDesc md = sc.members[i];
      const ArchiveView av = openArchive<KIND>(static_cast<const uint8_t*>(md.in));
      bool ok = av.ok;
      uint32_t need = 0, nb = 0, storedChecksum = 0;
      if (ok) {
        const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(av.ans));
        const uint4 h1 = __ldg(reinterpret_cast<const uint4*>(av.ans) + 1);
        nb = h0.y;
        need = h0.z;  // uncompressed bytes == float words for float kinds
        ok = h0.x == kAnsMagicVersion && (int)(h1.x & 0xfu) == pb && nb == divUp(need, kBlockBytes);
        if (KIND != kKindBytes) ok = ok && need == av.floatWords;
        storedChecksum = KIND == kKindBytes
            ? h1.y
            : __ldg(reinterpret_cast<const uint32_t*>(md.in) + 3);
      }
      // ans/GpuANSDecode.cuh:326-337: success iff capacity suffices; size reported regardless
      const bool success = ok && md.size >= need;
      if (outSuccess) outSuccess[i] = success ? 1 : 0;
      if (outSize) outSize[i] = ok ? need : 0u;
      if (wantChecksum) {
        sc.archiveChecksum[i] = storedChecksum;
        sc.sizes[i] = success ? need : 0u;
      }
      blocks = success ? nb : 0u;
    }
    uint32_t tot;
    const uint32_t excl = blockExclusiveScan<1024>(blocks, sWarp, &tot);
    if (i < n) sc.members[i].work0 = carry + excl;
    carry += tot;
  }
  if (threadIdx.x == 0) sc.totals[0] = carry;
}

// ---------------------------------------------------------------------------
// Output writers (ans/BatchProvider.cuh:16-37 BatchWriter, float/
// GpuFloatDecompress.cuh:391-486 JoinFloatWriter restated).  A writer is bound
// to one 4 KiB block; `at(row)` yields a cursor for a group of rows so that the
// unrolled loop addresses rows with compile-time offsets (row J of the group is
// element J*32 from the cursor).  `prefetch` fetches the stored byte(s) of a
// row ahead of the dependent decode chain.
// ---------------------------------------------------------------------------
template <int KIND>
struct RowWriter;

template <>
struct RowWriter<kKindBytes> {
  uint8_t* out;
  struct Pre {};
  struct Cursor { uint8_t* o; };
  __device__ __forceinline__ void setBlock(const ArchiveView&, void* outBase, uint32_t block, uint32_t lane) {
    out = static_cast<uint8_t*>(outBase) + (size_t)block * kBlockBytes + lane;
    __builtin_assume(__isGlobal(out));
  }
  __device__ __forceinline__ Cursor at(uint32_t row) const { return Cursor{out + row * 32u}; }
  template <int J>
  __device__ __forceinline__ Pre prefetch(const Cursor&) const { return Pre{}; }
  template <int J>
  __device__ __forceinline__ void write(const Cursor& c, uint32_t entry, Pre) const {
    c.o[J * 32] = (uint8_t)entry;
  }
};

template <int KIND>
struct RowWriter16 {
  uint16_t* out;
  const uint8_t* non;
  typedef uint32_t Pre;
  struct Cursor { uint16_t* o; const uint8_t* n; };
  __device__ __forceinline__ void setBlock(const ArchiveView& av, void* outBase, uint32_t block, uint32_t lane) {
    out = static_cast<uint16_t*>(outBase) + (size_t)block * kBlockBytes + lane;
    non = av.non + (size_t)block * kBlockBytes + lane;
    __builtin_assume(__isGlobal(out));
    __builtin_assume(__isGlobal(non));
  }
  __device__ __forceinline__ Cursor at(uint32_t row) const { return Cursor{out + row * 32u, non + row * 32u}; }
  template <int J>
  __device__ __forceinline__ Pre prefetch(const Cursor& c) const { return __ldg(c.n + J * 32); }
  template <int J>
  __device__ __forceinline__ void write(const Cursor& c, uint32_t entry, Pre nc) const {
    uint32_t v;
    if (KIND == kKindF16) {
      // float/GpuFloatUtils.cuh:117-119: comp * 256 + nonComp  (bytes: [non, comp])
      v = __byte_perm(entry, nc, 0x4404);
    } else {
      // float/GpuFloatUtils.cuh:149-159: (comp:non) rotated right by one within 16 bits.
      // x = comp<<24 | non<<16 ; (non : x) >> 17 leaves comp<<7 | non>>1 | (non&1)<<15 in the
      // low 16 bits (higher bits are dropped by the 16-bit store)
      v = __funnelshift_r(__byte_perm(entry, nc, 0x0444), nc, 17);
    }
    c.o[J * 32] = (uint16_t)v;
  }
};
template <> struct RowWriter<kKindF16> : RowWriter16<kKindF16> {};
template <> struct RowWriter<kKindBF16> : RowWriter16<kKindBF16> {};

template <>
struct RowWriter<kKindF32> {
  uint32_t* out;
  const uint16_t* non2;
  const uint8_t* non1;
  struct Pre { uint32_t lo, hi; };
  struct Cursor { uint32_t* o; const uint16_t* n2; const uint8_t* n1; };
  __device__ __forceinline__ void setBlock(const ArchiveView& av, void* outBase, uint32_t block, uint32_t lane) {
    out = static_cast<uint32_t*>(outBase) + (size_t)block * kBlockBytes + lane;
    non2 = reinterpret_cast<const uint16_t*>(av.non) + (size_t)block * kBlockBytes + lane;
    non1 = av.non + 2u * (size_t)roundUp(av.floatWords, 8u) + (size_t)block * kBlockBytes + lane;
    __builtin_assume(__isGlobal(out));
    __builtin_assume(__isGlobal(non2));
    __builtin_assume(__isGlobal(non1));
  }
  __device__ __forceinline__ Cursor at(uint32_t row) const {
    return Cursor{out + row * 32u, non2 + row * 32u, non1 + row * 32u};
  }
  template <int J>
  __device__ __forceinline__ Pre prefetch(const Cursor& c) const {
    Pre p;
    p.lo = __ldg(c.n2 + J * 32);
    p.hi = __ldg(c.n1 + J * 32);
    return p;
  }
  template <int J>
  __device__ __forceinline__ void write(const Cursor& c, uint32_t entry, Pre p) const {
    // float/GpuFloatUtils.cuh:187-190: rotate right by one
    const uint32_t v = __byte_perm(p.lo, __byte_perm(p.hi, entry, 0x0040), 0x5410);
    c.o[J * 32] = __funnelshift_r(v, v, 1);
  }
};

// ---------------------------------------------------------------------------
// One decode step for a full row (ans/GpuANSDecode.cuh:55-105 restated).
// LUT entry: [31:20] pdf, [19:8] s-cdf, [7:0] symbol  (this kernel's own
// layout; the LUT never leaves shared memory).
// `words` points one past the last unread word of the block's stream, either
// in shared memory (staged) or in global memory.
// ---------------------------------------------------------------------------
template <int PB>
__device__ __forceinline__ uint32_t decodeStep(uint32_t& state, const uint32_t* __restrict__ lut,
                                               const uint16_t*& words, uint32_t geMask) {
  constexpr uint32_t mask = (1u << PB) - 1u;
  const uint32_t e = lut[state & mask];
  state = (e >> 20) * (state >> PB) + ((e >> 8) & 0xfffu);
  const bool rd = state < kStateMin;
  const uint32_t vote = __ballot_sync(0xffffffffu, rd);
  if (rd) {
    const uint32_t w = *(words - __popc(vote & geMask));
    state = (state << 16) + w;
  }
  words -= __popc(vote);
  return e;
}

template <int PB>
__device__ __forceinline__ uint32_t decodeStepPartial(bool valid, uint32_t& state,
                                                      const uint32_t* __restrict__ lut,
                                                      const uint16_t*& words, uint32_t geMask) {
  constexpr uint32_t mask = (1u << PB) - 1u;
  const uint32_t e = lut[state & mask];
  if (valid) state = (e >> 20) * (state >> PB) + ((e >> 8) & 0xfffu);
  const bool rd = valid && state < kStateMin;
  const uint32_t vote = __ballot_sync(0xffffffffu, rd);
  if (rd) {
    const uint32_t w = *(words - __popc(vote & geMask));
    state = (state << 16) + w;
  }
  words -= __popc(vote);
  return e;
}

template <int KIND, int PB, int J>
struct RowGroup {
  // rows J-1 .. 0 of the group, highest first (the decoder walks rows backwards)
  template <typename PreArr>
  static __device__ __forceinline__ void load(const RowWriter<KIND>& wr,
                                              const typename RowWriter<KIND>::Cursor& c, PreArr& pre) {
    pre[J - 1] = wr.template prefetch<J - 1>(c);
    RowGroup<KIND, PB, J - 1>::load(wr, c, pre);
  }
  template <typename PreArr>
  static __device__ __forceinline__ void run(uint32_t& state, const uint32_t* __restrict__ lut,
                                             const uint16_t*& words, uint32_t geMask,
                                             const RowWriter<KIND>& wr,
                                             const typename RowWriter<KIND>::Cursor& c, PreArr& pre) {
    const uint32_t e = decodeStep<PB>(state, lut, words, geMask);
    wr.template write<J - 1>(c, e, pre[J - 1]);
    RowGroup<KIND, PB, J - 1>::run(state, lut, words, geMask, wr, c, pre);
  }
};
template <int KIND, int PB>
struct RowGroup<KIND, PB, 0> {
  template <typename PreArr>
  static __device__ __forceinline__ void load(const RowWriter<KIND>&, const typename RowWriter<KIND>::Cursor&, PreArr&) {}
  template <typename PreArr>
  static __device__ __forceinline__ void run(uint32_t&, const uint32_t* __restrict__, const uint16_t*&, uint32_t,
                                             const RowWriter<KIND>&, const typename RowWriter<KIND>::Cursor&, PreArr&) {}
};

template <int KIND, int PB>
__device__ __forceinline__ void decodeBlockWarp(uint32_t state, const uint16_t* wordsEnd,
                                                uint32_t n, const uint32_t* __restrict__ lut,
                                                const RowWriter<KIND>& wr, uint32_t lane) {
  typedef typename RowWriter<KIND>::Pre Pre;
  const uint32_t geMask = laneMaskGe();
  const uint16_t* words = wordsEnd;
  uint32_t row = n >> 5;  // number of full rows; the partial row (if any) has this index
  const uint32_t rem = n & 31u;
  if (rem) {
    const bool valid = lane < rem;
    const typename RowWriter<KIND>::Cursor c = wr.at(row);
    Pre pre[1] = {};
    if (valid) pre[0] = wr.template prefetch<0>(c);
    const uint32_t e = decodeStepPartial<PB>(valid, state, lut, words, geMask);
    if (valid) wr.template write<0>(c, e, pre[0]);
  }
  constexpr int U = 8;
  while (row >= (uint32_t)U) {
    row -= U;
    const typename RowWriter<KIND>::Cursor c = wr.at(row);
    Pre pre[U];
    RowGroup<KIND, PB, U>::load(wr, c, pre);
    RowGroup<KIND, PB, U>::run(state, lut, words, geMask, wr, c, pre);
  }
  while (row > 0) {
    --row;
    const typename RowWriter<KIND>::Cursor c = wr.at(row);
    Pre pre[1];
    pre[0] = wr.template prefetch<0>(c);
    const uint32_t e = decodeStep<PB>(state, lut, words, geMask);
    wr.template write<0>(c, e, pre[0]);
  }
}

// LUT build from the archive's u16 pdf[256] (ans/GpuANSDecode.cuh:405-476
// restated; runs inside the decode CTA).  blockDim.x == WARPS*32.
template <int PB, int WARPS>
__device__ void buildLut(const uint8_t* __restrict__ ans, uint32_t* __restrict__ lut,
                         uint32_t* sPdf, uint32_t* sCdf, uint32_t* sWarp) {
  constexpr int T = WARPS * 32;
  constexpr int PER = (kNumSymbols + T - 1) / T;
  const uint16_t* pdfIn = reinterpret_cast<const uint16_t*>(ans + kAnsHeaderBytes);
  // contiguous PER symbols per thread so one block scan gives the cdf
  uint32_t p[PER];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const uint32_t sym = threadIdx.x * PER + k;
    p[k] = sym < kNumSymbols ? (uint32_t)__ldg(pdfIn + sym) : 0u;
    s += p[k];
  }
  uint32_t tot;
  uint32_t c = blockExclusiveScan<T>(s, sWarp, &tot);
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const uint32_t sym = threadIdx.x * PER + k;
    if (sym < kNumSymbols) { sPdf[sym] = p[k]; sCdf[sym] = c; }
    c += p[k];
  }
  __syncthreads();
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  for (uint32_t sym = warp; sym < kNumSymbols; sym += WARPS) {
    const uint32_t pdf = sPdf[sym], begin = sCdf[sym];
    for (uint32_t j = lane; j < pdf; j += 32u) {
      if (begin + j < (1u << PB)) lut[begin + j] = (pdf << 20) | (j << 8) | sym;
    }
  }
  __syncthreads();
}

template <int KIND, int PB, int WARPS, bool STAGE>
__global__ void __launch_bounds__(WARPS * 32)
decodeKernel(DecodeScratch sc, uint32_t n) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr uint32_t K = 1u << PB;
  constexpr uint32_t slotBytes = 128u + maxBlockWords(PB) * 2u;  // lane states + stream
  __shared__ __align__(16) uint32_t lut[K];  // static: constant base address for the hot LDS
  uint32_t* sPdf = reinterpret_cast<uint32_t*>(smem);
  uint32_t* sCdf = sPdf + kNumSymbols;
  uint32_t* sWarp = sCdf + kNumSymbols;           // 32 words
  uint32_t* sMisc = sWarp + 32;                   // 32 words
  unsigned long long* sBar = reinterpret_cast<unsigned long long*>(sMisc + 32);  // [WARPS]
  uint8_t* sSlots = reinterpret_cast<uint8_t*>(sBar + ((WARPS + 1) & ~1));

  const uint32_t t = threadIdx.x, lane = t & 31u;
  // shuffle makes the warp index provably warp-uniform, so the vote in the hot loop needs no
  // divergence check (BRA.DIV)
  const uint32_t warp = __shfl_sync(0xffffffffu, t >> 5, 0);
  uint8_t* mySlot = sSlots + (size_t)warp * slotBytes;
  const uint32_t myBar = smemAddr(sBar + warp);
  uint32_t phase = 0;
  if (STAGE) {
    if (lane == 0) mbarInit(myBar, 1);
    fenceBarrierInit();
    __syncthreads();
  }

  const uint32_t total = __ldcg(sc.totals);
  // contiguous, balanced run of flat blocks for this CTA
  const uint64_t g = gridDim.x;
  uint32_t cur = (uint32_t)((uint64_t)total * blockIdx.x / g);
  const uint32_t end = (uint32_t)((uint64_t)total * (blockIdx.x + 1) / g);

  while (cur < end) {
    // member containing flat block `cur`
    if (t == 0) {
      uint32_t lo = 0, hi = n;
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldcg(&sc.members[mid].work0) <= cur) lo = mid; else hi = mid;
      }
      // skip members that contribute no blocks (failed / empty): the search lands on the
      // last member whose start <= cur, which is the one owning `cur`
      sMisc[0] = lo;
    }
    __syncthreads();
    const uint32_t m = sMisc[0];
    const MemberDesc md = sc.members[m];
    const ArchiveView av = openArchive<KIND>(static_cast<const uint8_t*>(md.in));
    const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(av.ans));
    const uint32_t nb = h0.y;
    const uint32_t memberFirst = __ldcg(&sc.members[m].work0);
    const uint32_t memberEnd = min(end, memberFirst + nb);
    buildLut<PB, WARPS>(av.ans, lut, sPdf, sCdf, sWarp);

    const uint8_t* pStates = av.ans + kAnsHeaderBytes + kAnsPdfBytes;
    const uint2* pBlockWords = reinterpret_cast<const uint2*>(pStates + 128u * (size_t)nb);
    const uint16_t* pData = reinterpret_cast<const uint16_t*>(
        reinterpret_cast<const uint8_t*>(pBlockWords) + 8u * (size_t)roundUp(nb, 2u));
    const bool canStage = STAGE && ((reinterpret_cast<uintptr_t>(av.ans) & 15u) == 0);

    RowWriter<KIND> wr;
    for (uint32_t fb = cur + warp; fb < memberEnd; fb += WARPS) {
      const uint32_t block = fb - memberFirst;
      const uint2 bw = __ldg(pBlockWords + block);
      const uint32_t blockLen = bw.x >> 16, words = bw.x & 0xffffu;
      const uint16_t* stream = pData + bw.y;
      wr.setBlock(av, md.out, block, lane);
      // two call sites on purpose: the compiler then knows the address space of the stream
      // (LDS for the staged copy, LDG for the direct path) instead of a generic pointer
      if (canStage && words <= maxBlockWords(PB) && (bw.y & 7u) == 0u) {
        const uint32_t streamBytes = roundUp(words, 8u) * 2u;
        __syncwarp();
        if (lane == 0) {
          mbarExpectTx(myBar, 128u + streamBytes);
          bulkLoad(smemAddr(mySlot), pStates + 128u * (size_t)block, 128u, myBar);
          if (streamBytes) bulkLoad(smemAddr(mySlot + 128), stream, streamBytes, myBar);
        }
        mbarWait(myBar, phase);
        phase ^= 1u;
        const uint32_t state = reinterpret_cast<const uint32_t*>(mySlot)[lane];
        decodeBlockWarp<KIND, PB>(state, reinterpret_cast<const uint16_t*>(mySlot + 128) + words,
                                  blockLen, lut, wr, lane);
      } else {
        const uint32_t state = __ldg(reinterpret_cast<const uint32_t*>(pStates) + block * 32u + lane);
        decodeBlockWarp<KIND, PB>(state, stream + words, blockLen, lut, wr, lane);
      }
    }
    cur = memberEnd;
    __syncthreads();  // everyone done with this member's LUT before it is rebuilt
  }
}

// XOR checksum of decoded outputs (ans/GpuChecksum.cuh:26-93 semantics: XOR of
// all bytes folded to 8 bits).  grid = (n, Y).
__global__ void __launch_bounds__(256)
checksumKernel(DecodeScratch sc) {
  const uint32_t m = blockIdx.x;
  const uint8_t* p = static_cast<const uint8_t*>(sc.members[m].out);
  const uint32_t size = sc.sizes[m];
  uint32_t x = 0;
  for (uint32_t i = blockIdx.y * blockDim.x + threadIdx.x; i < size; i += gridDim.y * blockDim.x)
    x ^= p[i];
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) x ^= __shfl_xor_sync(0xffffffffu, x, d);
  if ((threadIdx.x & 31) == 0 && x) atomicXor(sc.checksum + m, x & 0xffu);
}

size_t alignUp256(size_t v) { return (v + 255) & ~size_t(255); }

struct DecodePlan {
  size_t members, totals, checksum, archiveChecksum, sizes, total;
};

DecodePlan planDecodeScratch(uint32_t n) {
  DecodePlan p{};
  size_t o = 0;
  p.members = o; o = alignUp256(o + sizeof(MemberDesc) * (size_t)n);
  p.totals = o; o = alignUp256(o + 16);
  p.checksum = o; o = alignUp256(o + 4 * (size_t)n);
  p.archiveChecksum = o; o = alignUp256(o + 4 * (size_t)n);
  p.sizes = o; o = alignUp256(o + 4 * (size_t)n);
  p.total = o;
  return p;
}

int smCountD() {
  static int cached = 0;
  if (!cached) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    if (cached <= 0) cached = 148;
  }
  return cached;
}

template <int KIND, int PB, int WARPS, bool STAGE>
int launchDecode(const DecodeScratch& sc, uint32_t n, uint64_t blockBound, cudaStream_t stream) {
  auto kern = decodeKernel<KIND, PB, WARPS, STAGE>;
  constexpr uint32_t K = 1u << PB;
  size_t smemBytes = (2 * kNumSymbols + 64) * 4 + ((WARPS + 1) & ~1) * 8;
  (void)K;
  if (STAGE) smemBytes += (size_t)WARPS * (128u + maxBlockWords(PB) * 2u);
  static int perSm = 0;  // per instantiation; one device per process (one rank per GPU)
  if (perSm == 0) {
    DGB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
    int occ = 0;
    DGB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smemBytes));
    perSm = std::max(occ, 1);
  }
  const uint64_t want = (blockBound + WARPS - 1) / WARPS;
  const uint32_t grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)perSm * smCountD()));
  kern<<<grid, WARPS * 32, smemBytes, stream>>>(sc, n);
  DGB_CUDA_TRY(cudaGetLastError());
  return DGB_OK;
}

template <int KIND, int PB>
int launchDecodeW(const DecodeScratch& sc, uint32_t n, uint64_t blockBound, cudaStream_t stream) {
  const Options& opt = options();
  const bool stage = opt.decode_stage != 0;
  switch (opt.decode_warps) {
    case 2:
      return stage ? launchDecode<KIND, PB, 2, true>(sc, n, blockBound, stream)
                   : launchDecode<KIND, PB, 2, false>(sc, n, blockBound, stream);
    case 8:
      return stage ? launchDecode<KIND, PB, 8, true>(sc, n, blockBound, stream)
                   : launchDecode<KIND, PB, 8, false>(sc, n, blockBound, stream);
    default:
      return stage ? launchDecode<KIND, PB, 4, true>(sc, n, blockBound, stream)
                   : launchDecode<KIND, PB, 4, false>(sc, n, blockBound, stream);
  }
}

template <int KIND>
int decodeKind(const DecodeScratch& sc, int pb, bool checksum, uint32_t n, uint64_t blockBound,
               uint8_t* outSuccess, uint32_t* outSize, cudaStream_t stream) {
  planKernel<KIND><<<1, 1024, 0, stream>>>(sc, n, pb, outSuccess, outSize, checksum);
  DGB_CUDA_TRY(cudaGetLastError());
  switch (pb) {
    case 9: return launchDecodeW<KIND, 9>(sc, n, blockBound, stream);
    case 10: return launchDecodeW<KIND, 10>(sc, n, blockBound, stream);
    case 11: return launchDecodeW<KIND, 11>(sc, n, blockBound, stream);
    default: return DGB_ERR_INVALID_ARG;
  }
}

}  // namespace

size_t decodeTempBytes(int /*kind*/, uint32_t n) { return planDecodeScratch(n).total + 256; }

int decodeBatch(int kind, void* temp, size_t tempBytes, int pb, bool checksum, uint32_t n,
                const HostMember* members, uint8_t* outSuccess_dev, uint32_t* outSize_dev,
                uint8_t* mismatchHost, cudaStream_t stream) {
  if (n == 0) return DGB_OK;
  if (pb < 9 || pb > 11) return DGB_ERR_INVALID_ARG;
  const DecodePlan dp = planDecodeScratch(n);
  if (!temp || tempBytes < dp.total) return DGB_ERR_TEMP_TOO_SMALL;
  if (reinterpret_cast<uintptr_t>(temp) & 255u) return DGB_ERR_INVALID_ARG;

  std::vector<MemberDesc> desc(n);
  uint64_t blockBound = 0;
  const uint32_t wordBytes = kind == kKindF32 ? 4u : (kind == kKindBytes ? 1u : 2u);
  for (uint32_t i = 0; i < n; ++i) {
    if (!members[i].in || (members[i].size && !members[i].out)) return DGB_ERR_INVALID_ARG;
    // headers are read as 16 B vectors; the reference imposes the same alignment
    if (reinterpret_cast<uintptr_t>(members[i].in) & 15u) return DGB_ERR_INVALID_ARG;
    if (reinterpret_cast<uintptr_t>(members[i].out) & (wordBytes - 1u)) return DGB_ERR_INVALID_ARG;
    desc[i].in = members[i].in;
    desc[i].out = members[i].out;
    desc[i].size = members[i].size;  // capacity
    desc[i].work0 = 0;
    blockBound += divUp(members[i].size, kBlockBytes);
  }
  uint8_t* base = static_cast<uint8_t*>(temp);
  DecodeScratch sc;
  sc.members = reinterpret_cast<MemberDesc*>(base + dp.members);
  sc.totals = reinterpret_cast<uint32_t*>(base + dp.totals);
  sc.checksum = reinterpret_cast<uint32_t*>(base + dp.checksum);
  sc.archiveChecksum = reinterpret_cast<uint32_t*>(base + dp.archiveChecksum);
  sc.sizes = reinterpret_cast<uint32_t*>(base + dp.sizes);
  DGB_CUDA_TRY(cudaMemcpyAsync(sc.members, desc.data(), sizeof(MemberDesc) * n,
                               cudaMemcpyHostToDevice, stream));
  if (checksum) DGB_CUDA_TRY(cudaMemsetAsync(sc.checksum, 0, 4 * (size_t)n, stream));

  int rc;
  switch (kind) {
    case kKindBytes: rc = decodeKind<kKindBytes>(sc, pb, checksum, n, blockBound, outSuccess_dev, outSize_dev, stream); break;
    case kKindF16: rc = decodeKind<kKindF16>(sc, pb, checksum, n, blockBound, outSuccess_dev, outSize_dev, stream); break;
    case kKindBF16: rc = decodeKind<kKindBF16>(sc, pb, checksum, n, blockBound, outSuccess_dev, outSize_dev, stream); break;
    case kKindF32: rc = decodeKind<kKindF32>(sc, pb, checksum, n, blockBound, outSuccess_dev, outSize_dev, stream); break;
    default: return DGB_ERR_INVALID_ARG;
  }
  if (rc != DGB_OK) return rc;

  if (checksum) {
    // ans/GpuANSDecode.cuh:555-591 / float/GpuFloatDecompress.cuh:698-733: checksum the
    // output, compare with the archive's on the host (this path synchronises).
    // Float kinds: the checksum covers the first `size` BYTES; sc.sizes holds
    // the word count, which is exactly that byte count.
    const int sms = smCountD();
    dim3 grid(n, std::max(1, std::min(64, 4 * sms / (int)std::max(1u, n))));
    checksumKernel<<<grid, 256, 0, stream>>>(sc);
    DGB_CUDA_TRY(cudaGetLastError());
    std::vector<uint32_t> got(n), want(n), sizes(n);
    DGB_CUDA_TRY(cudaMemcpyAsync(got.data(), sc.checksum, 4 * (size_t)n, cudaMemcpyDeviceToHost, stream));
    DGB_CUDA_TRY(cudaMemcpyAsync(want.data(), sc.archiveChecksum, 4 * (size_t)n, cudaMemcpyDeviceToHost, stream));
    DGB_CUDA_TRY(cudaMemcpyAsync(sizes.data(), sc.sizes, 4 * (size_t)n, cudaMemcpyDeviceToHost, stream));
    DGB_CUDA_TRY(cudaStreamSynchronize(stream));
    bool bad = false;
    for (uint32_t i = 0; i < n; ++i) {
      // members that were skipped (capacity / bad header) are reported through outSuccess
      const bool mm = sizes[i] != 0 && got[i] != want[i];
      if (mismatchHost) mismatchHost[i] = mm ? 1 : 0;
      bad = bad || mm;
    }
    if (bad) return DGB_ERR_CHECKSUM;
  }
  return DGB_OK;
}

// ---------------------------------------------------------------------------
// Header queries (ans/GpuANSInfo.cuh:16-37, float/GpuFloatInfo.cuh:19-41)
// ---------------------------------------------------------------------------
namespace {
__global__ void infoKernel(int kind, const void* const* in, uint32_t n, uint32_t* outSizes,
                           uint32_t* outTypes, uint32_t* outChecksum) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t* h = static_cast<const uint32_t*>(in[i]);
  if (kind == kKindBytes) {
    const bool ok = h[0] == kAnsMagicVersion;
    if (outSizes) outSizes[i] = ok ? h[2] : 0u;  // UNCOMPRESSED bytes (ans/GpuANSInfo.cuh:27-29)
    if (outChecksum) outChecksum[i] = ok ? h[5] : 0u;
  } else {
    const bool ok = h[0] == kFloatMagicVersion;
    if (outSizes) outSizes[i] = ok ? h[1] : 0u;  // float words
    if (outTypes) outTypes[i] = ok ? (h[2] & 0xfu) : 0u;
    if (outChecksum) outChecksum[i] = ok ? h[3] : 0u;
  }
}
}  // namespace

int getInfo(int kind, void* temp, size_t tempBytes, const void* const* in, bool inIsDevice,
            uint32_t n, uint32_t* outSizes, uint32_t* outTypes, uint32_t* outChecksum,
            cudaStream_t stream) {
  if (n == 0) return DGB_OK;
  if (!in) return DGB_ERR_INVALID_ARG;
  const void* const* in_dev = in;
  if (!inIsDevice) {
    if (!temp || tempBytes < sizeof(void*) * (size_t)n) return DGB_ERR_TEMP_TOO_SMALL;
    DGB_CUDA_TRY(cudaMemcpyAsync(temp, in, sizeof(void*) * (size_t)n, cudaMemcpyHostToDevice, stream));
    in_dev = static_cast<const void* const*>(temp);
  }
  infoKernel<<<divUp(n, 128), 128, 0, stream>>>(kind, in_dev, n, outSizes, outTypes, outChecksum);
  DGB_CUDA_TRY(cudaGetLastError());
  return DGB_OK;
}

}  // namespace dgb
