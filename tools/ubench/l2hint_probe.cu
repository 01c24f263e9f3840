// Which .L2::cache_hint forms run on this GPU?  One kernel per form, error checked after each.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t pol(int k) {
  uint64_t p;
  if (k == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else if (k == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
template <int FORM, int POL>
__global__ void k(const uint4* in, uint4* out) {
  __shared__ uint4 sm[64];
  const uint64_t p = pol(POL);
  uint4 v = make_uint4(1, 2, 3, 4);
  if (FORM == 0) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(in + threadIdx.x), "l"(p));
  if (FORM == 1) asm volatile("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(in + threadIdx.x), "l"(p));
  if (FORM == 2) asm volatile("st.global.L2::cache_hint.v2.u32 [%0], {%1,%2}, %3;" ::"l"(out + threadIdx.x), "r"(v.x), "r"(v.y), "l"(p) : "memory");
  if (FORM == 3) asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(out + threadIdx.x), "r"(v.x), "l"(p) : "memory");
  if (FORM == 4) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(sm) + threadIdx.x * 16;
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(in + threadIdx.x), "l"(p) : "memory");
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    v = sm[threadIdx.x];
  }
  if (FORM == 5) asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(in + threadIdx.x), "l"(p));
  out[64 + threadIdx.x] = v;
}
template <int FORM, int POL>
void run(const char* name, const uint4* in, uint4* out) {
  k<FORM, POL><<<1, 32>>>(in, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-44s policy %d: %s\n", name, POL, cudaGetErrorString(e));
  if (e != cudaSuccess) { cudaDeviceReset(); exit(0); }
}
int main() {
  uint4 *in, *out;
  cudaMalloc(&in, 4096);
  cudaMalloc(&out, 4096);
  cudaMemset(in, 1, 4096);
#define ALL(F, N) run<F, 0>(N, in, out); run<F, 1>(N, in, out); run<F, 2>(N, in, out);
  ALL(1, "ld.global.L2::cache_hint.v4")
  ALL(5, "ld.global.nc.L2::cache_hint.v4")
  ALL(2, "st.global.L2::cache_hint.v2")
  ALL(3, "st.global.L2::cache_hint.u32")
  ALL(4, "cp.async.cg.L2::cache_hint 16")
  ALL(0, "ld.global.nc.L1::no_allocate.L2::cache_hint")
  return 0;
}
