// Microbenchmark: achievable DRAM read bandwidth for the NCHW "column walk" (each thread walks
// C channel planes of stride HW) versus a linear read of the same bytes.  Working set 464 MB >> L2.
#include <cstdio>
#include <cuda_runtime.h>
template <int UNROLL, int C, int VEC>
__global__ void colwalk(const float* __restrict__ x, float* __restrict__ out, int HW) {
  const int n = blockIdx.y;
  const int hw = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (hw >= HW) return;
  const float* p = x + (size_t)n * C * HW + hw;
  float acc[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
#pragma unroll UNROLL
  for (int c = 0; c < C; ++c) {
    if (VEC == 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p + (size_t)c * HW));
      acc[0] = fmaxf(acc[0], v.x); acc[1 % VEC] = fmaxf(acc[1 % VEC], v.y); acc[2 % VEC] = fmaxf(acc[2 % VEC], v.z); acc[3 % VEC] = fmaxf(acc[3 % VEC], v.w);
    } else if (VEC == 2) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(p + (size_t)c * HW));
      acc[0] = fmaxf(acc[0], v.x); acc[1 % VEC] = fmaxf(acc[1 % VEC], v.y);
    } else {
      acc[0] = fmaxf(acc[0], __ldg(p + (size_t)c * HW));
    }
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k) out[(size_t)n * HW + hw + k] = acc[k];
}
// channel-major variant: a CTA owns (n, hw tile) but its warps split the channels; partial maxima in smem
template <int C>
__global__ void chansplit(const float* __restrict__ x, float* __restrict__ out, int HW) {
  __shared__ float4 s[8][32];
  const int n = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hw = (blockIdx.x * 32 + lane) * 4;
  float4 acc = make_float4(0, 0, 0, 0);
  if (hw < HW) {
    const float* p = x + (size_t)n * C * HW + hw;
#pragma unroll 7
    for (int c = warp; c < C; c += 8) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p + (size_t)c * HW));
      acc.x = fmaxf(acc.x, v.x); acc.y = fmaxf(acc.y, v.y); acc.z = fmaxf(acc.z, v.z); acc.w = fmaxf(acc.w, v.w);
    }
  }
  s[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && hw < HW) {
    for (int w = 1; w < 8; ++w) { float4 v = s[w][lane]; acc.x = fmaxf(acc.x, v.x); acc.y = fmaxf(acc.y, v.y); acc.z = fmaxf(acc.z, v.z); acc.w = fmaxf(acc.w, v.w); }
    *reinterpret_cast<float4*>(out + (size_t)n * HW + hw) = acc;
  }
}
__global__ void linear(const float4* __restrict__ x, float* __restrict__ out, size_t n4) {
  float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 8
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x + i);
    acc.x = fmaxf(acc.x, v.x); acc.y = fmaxf(acc.y, v.y); acc.z = fmaxf(acc.z, v.z); acc.w = fmaxf(acc.w, v.w);
  }
  if (acc.x + acc.y + acc.z + acc.w == 12345.f) out[0] = acc.x;
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
  for (int threads : {64, 128, 256}) {
    dim3 g4((HW / 4 + threads - 1) / threads, N), g2((HW / 2 + threads - 1) / threads, N), g1((HW + threads - 1) / threads, N);
    float a = timeit([&] { colwalk<8, C, 4><<<g4, threads>>>(x, out, HW); });
    float b = timeit([&] { colwalk<27, C, 4><<<g4, threads>>>(x, out, HW); });
    float c = timeit([&] { colwalk<8, C, 2><<<g2, threads>>>(x, out, HW); });
    float d = timeit([&] { colwalk<8, C, 1><<<g1, threads>>>(x, out, HW); });
    float e = timeit([&] { colwalk<27, C, 1><<<g1, threads>>>(x, out, HW); });
    printf("threads %3d MB %.0f | vec4 u8 %.1f us (%.0f GB/s) u27 %.1f (%.0f) | vec2 u8 %.1f (%.0f) | vec1 u8 %.1f (%.0f) u27 %.1f (%.0f)\n",
           threads, mb, a * 1e3, mb / a, b * 1e3, mb / b, c * 1e3, mb / c, d * 1e3, mb / d, e * 1e3, mb / e);
  }
  dim3 gs((HW / 4 + 31) / 32, N);
  float s = timeit([&] { chansplit<C><<<gs, 256>>>(x, out, HW); });
  printf("chansplit (8 warps split channels, 512 B rows) %.1f us (%.0f GB/s)\n", s * 1e3, mb / s);
  for (int blocks : {148 * 4, 148 * 8, 148 * 16}) {
    float l = timeit([&] { linear<<<blocks, 256>>>(reinterpret_cast<const float4*>(x), out, elems / 4); });
    printf("linear read, %d blocks: %.1f us (%.0f GB/s)\n", blocks, l * 1e3, mb / l);
  }
  return 0;
}
