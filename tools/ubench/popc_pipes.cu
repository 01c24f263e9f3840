// Micro-benchmark: which issue pipes do POPC / REDUX / SHFL / VOTE share on sm_100a?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o popc_pipes popc_pipes.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) loopKernel(uint32_t* out, uint32_t seed) {
  uint32_t a = seed + threadIdx.x, b = seed * 3u + threadIdx.x, acc = 0;
  const uint32_t lt = (1u << (threadIdx.x & 31)) - 1u;
#pragma unroll 8
  for (int i = 0; i < kIters; ++i) {
    const bool p = (a ^ (uint32_t)i) & 4u;
    const uint32_t v = __ballot_sync(0xffffffffu, p);
    if (MODE == 0) {            // prefix popc only
      acc += __popc(v & lt);
    } else if (MODE == 1) {     // prefix popc + total popc
      acc += __popc(v & lt);
      b += __popc(v);
    } else if (MODE == 2) {     // prefix popc + redux total
      acc += __popc(v & lt);
      b += __reduce_add_sync(0xffffffffu, p ? 1u : 0u);
    } else if (MODE == 3) {     // redux only
      b += __reduce_add_sync(0xffffffffu, p ? 1u : 0u);
    } else if (MODE == 4) {     // prefix popc + shfl of lane 31's inclusive count
      const uint32_t pre = __popc(v & lt);
      acc += pre;
      b += __shfl_sync(0xffffffffu, pre + (p ? 1u : 0u), 31);
    } else if (MODE == 5) {     // vote only
      acc += v;
    } else if (MODE == 6) {     // two redux
      b += __reduce_add_sync(0xffffffffu, p ? 1u : 0u);
      acc += __reduce_add_sync(0xffffffffu, a & 1u);
    }
    a = a * 1664525u + b;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + acc;
}

template <int MODE>
void run(const char* name, uint32_t* out, int ctas) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  loopKernel<MODE><<<ctas, 256>>>(out, 1);
  cudaEventRecord(e0);
  loopKernel<MODE><<<ctas, 256>>>(out, 2);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  // warp-iterations per SMSP
  const double warpsPerSmsp = (double)ctas * 8 / sms / 4;
  const double cyc = ms * 1e-3 * clk * 1e3 / (warpsPerSmsp * kIters);
  printf("%-28s ctas=%5d  %.3f ms  ~%.2f cycles per warp-iteration per SMSP (at %d MHz nominal)\n", name, ctas, ms,
         cyc, clk / 1000);
}

int main() {
  uint32_t* out;
  cudaMalloc(&out, 148 * 8 * 256 * 4 * 2);
  for (int ctas : {148 * 4, 148 * 8}) {
    run<5>("vote only", out, ctas);
    run<0>("vote+popc(prefix)", out, ctas);
    run<1>("vote+popc+popc", out, ctas);
    run<2>("vote+popc+redux", out, ctas);
    run<3>("vote+redux", out, ctas);
    run<6>("vote+redux+redux", out, ctas);
    run<4>("vote+popc+shfl", out, ctas);
  }
  return 0;
}
