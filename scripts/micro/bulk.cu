// Microbenchmark: (1) column walk at restricted occupancy, (2) the same bytes staged through
// shared memory with cp.async.bulk + mbarrier (whole [C x T] tile in flight per CTA).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int C>
__global__ void colwalk_occ(const float* __restrict__ x, float* __restrict__ out, int HW) {
  extern __shared__ float dummy[];
  const int n = blockIdx.y;
  const int hw = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (hw >= HW) return;
  const float* p = x + (size_t)n * C * HW + hw;
  float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 8
  for (int c = 0; c < C; ++c) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p + (size_t)c * HW));
    acc.x = fmaxf(acc.x, v.x); acc.y = fmaxf(acc.y, v.y); acc.z = fmaxf(acc.z, v.z); acc.w = fmaxf(acc.w, v.w);
  }
  *reinterpret_cast<float4*>(out + (size_t)n * HW + hw) = acc;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// one CTA = T anchors x C channel rows, all rows requested up front with bulk copies
template <int C, int T>
__global__ void __launch_bounds__(T) bulk_tile(const float* __restrict__ x, float* __restrict__ out, int HW) {
  extern __shared__ __align__(128) float tile[];   // [C][T]
  __shared__ __align__(8) unsigned long long bar;
  const int n = blockIdx.y, hw0 = blockIdx.x * T;
  const int cnt = min(T, HW - hw0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)(C * cnt * 4)) : "memory");
    __syncwarp();
    for (int c = threadIdx.x; c < C; c += 32) {
      const float* src = x + ((size_t)n * C + c) * HW + hw0;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(tile + c * T)), "l"(src), "r"((uint32_t)(cnt * 4)), "r"(smem_u32(&bar)) : "memory");
    }
  }
  // wait phase 0
  uint32_t done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  }
  if (threadIdx.x < cnt) {
    float acc = 0.f;
#pragma unroll 12
    for (int c = 0; c < C; ++c) acc = fmaxf(acc, tile[c * T + threadIdx.x]);
    out[(size_t)n * HW + hw0 + threadIdx.x] = acc;
  }
}
template <typename F> float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(a);
  for (int i = 0; i < 20; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / 20;
}
int main() {
  const int N = 64, C = 108, HW = 16800;
  size_t elems = (size_t)N * C * HW;
  float *x, *out;
  cudaMalloc(&x, elems * 4); cudaMalloc(&out, (size_t)N * HW * 4);
  cudaMemset(x, 0, elems * 4);
  double mb = elems * 4 / 1e6;
  cudaFuncSetAttribute(colwalk_occ<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int smem_kb : {0, 16, 32, 48, 72, 100}) {   // 128-thread CTAs; smem limits CTAs/SM: 227/smem
    dim3 g((HW / 4 + 127) / 128, N);
    float a = timeit([&] { colwalk_occ<C><<<g, 128, smem_kb * 1024>>>(x, out, HW); });
    int ctas = smem_kb ? 227 / (smem_kb + 1) : 16;
    printf("colwalk vec4 u8, 128 thr, %3d KB smem (~%2d CTAs/SM = %4d thr): %.1f us (%.0f GB/s)\n", smem_kb, ctas > 16 ? 16 : ctas,
           (ctas > 16 ? 16 : ctas) * 128, a * 1e3, mb / a);
  }
  {
    constexpr int T = 256;
    cudaFuncSetAttribute(bulk_tile<C, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, C * T * 4);
    dim3 g((HW + T - 1) / T, N);
    float a = timeit([&] { bulk_tile<C, T><<<g, T, C * T * 4>>>(x, out, HW); });
    printf("bulk tile C x %d (%d KB smem, 2 CTAs/SM): %.1f us (%.0f GB/s)\n", T, C * T * 4 / 1024, a * 1e3, mb / a);
  }
  {
    constexpr int T = 128;
    cudaFuncSetAttribute(bulk_tile<C, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, C * T * 4);
    dim3 g((HW + T - 1) / T, N);
    float a = timeit([&] { bulk_tile<C, T><<<g, T, C * T * 4>>>(x, out, HW); });
    printf("bulk tile C x %d (%d KB smem, 4 CTAs/SM): %.1f us (%.0f GB/s)\n", T, C * T * 4 / 1024, a * 1e3, mb / a);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
