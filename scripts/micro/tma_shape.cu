// Dev micro-benchmark: how fast can one CTA per SM stream a [rows x HW] fp32 tensor through a
// shared-memory ring with 2-D TMA boxes of [R rows x T columns]?  (No compute; consumers only
// hand the slot back.)  Also: the same with a TMA store of every slot to a second tensor.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/tma_shape.bin scripts/micro/tma_shape.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t sa(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait_par(unsigned long long* bar, uint32_t par) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(sa(bar)), "r"(par) : "memory");
}

struct Args {
  int hw, rows_per_box, T, boxes_per_row_group, total, stages, stage_bytes, do_store, nload;
};

__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap in, const __grid_constant__ CUtensorMap out, Args a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long full[8], empty[8], done[8];
  const int S = a.stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(&empty[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(sa(&done[s])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const int tiles_x = (a.hw + a.T - 1) / a.T;
  if (warp == 0) {
    if (lane != 0) return;
    int k = 0;
    for (int t = blockIdx.x; t < a.total; t += gridDim.x, ++k) {
      const int slot = k % S;
      const uint32_t ph = (k / S) & 1;
      wait_par(&empty[slot], ph ^ 1);
      const int grp = t / tiles_x, tx = t - grp * tiles_x;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sa(&full[slot])), "r"(a.stage_bytes) : "memory");
      const int rows_per_load = a.rows_per_box / a.nload;
      for (int i = 0; i < a.nload; ++i)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                     ::"r"(sa(smem + (size_t)slot * a.stage_bytes + (size_t)i * rows_per_load * a.T * 4)), "l"(&in), "r"(tx * a.T),
                       "r"(grp * a.rows_per_box + i * rows_per_load), "r"(sa(&full[slot])), "l"(pol) : "memory");
    }
  } else if (warp == 1) {
    int k = 0;
    for (int t = blockIdx.x; t < a.total; t += gridDim.x, ++k) {
      const int slot = k % S;
      const uint32_t ph = (k / S) & 1;
      wait_par(&full[slot], ph);
      // touch the slot so the load is really consumed
      volatile float* p = reinterpret_cast<volatile float*>(smem + (size_t)slot * a.stage_bytes);
      float v = p[lane];
      if (v == 123.456f) p[lane] = 0.f;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(sa(&done[slot])) : "memory");
    }
  } else if (warp == 2) {
    if (lane != 0) return;
    int k = 0;
    for (int t = blockIdx.x; t < a.total; t += gridDim.x, ++k) {
      const int slot = k % S;
      const uint32_t ph = (k / S) & 1;
      wait_par(&done[slot], ph);
      if (a.do_store) {
        const int grp = t / tiles_x, tx = t - grp * tiles_x;
        const int rows_per_load = a.rows_per_box / a.nload;
        for (int i = 0; i < a.nload; ++i)
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;"
                       ::"l"(&out), "r"(tx * a.T), "r"(grp * a.rows_per_box + i * rows_per_load),
                         "r"(sa(smem + (size_t)slot * a.stage_bytes + (size_t)i * rows_per_load * a.T * 4)), "l"(pol) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(sa(&empty[slot])) : "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncFn enc = (EncFn)fp;
  const int hw = 16800, rows = 16 * 148;
  float *in, *out;
  cudaMalloc(&in, (size_t)rows * hw * 4);
  cudaMalloc(&out, (size_t)rows * hw * 4);
  cudaMemset(in, 0, (size_t)rows * hw * 4);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  struct V { int R, T, nload, promo; };
  const V vs[] = {{148, 64, 2, 2}, {148, 64, 2, 0}, {148, 64, 2, 1}, {74, 128, 1, 2}, {148, 128, 2, 2}, {37, 256, 1, 2}, {74, 256, 1, 2}, {148, 32, 2, 2}, {74, 64, 1, 2}};
  for (const V& v : vs) {
    for (int do_store = 0; do_store < 2; ++do_store) {
      for (int S = 2; S <= 6; S += 1) {
        CUtensorMap mi, mo;
        const cuuint64_t dims[2] = {(cuuint64_t)hw, (cuuint64_t)rows};
        const cuuint64_t str[1] = {(cuuint64_t)hw * 4};
        const cuuint32_t box[2] = {(cuuint32_t)v.T, (cuuint32_t)(v.R / v.nload)};
        const cuuint32_t es[2] = {1, 1};
        const CUtensorMapL2promotion pr = v.promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : v.promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
        if (enc(&mi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, in, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ||
            enc(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) {
          printf("encode failed R=%d T=%d\n", v.R, v.T);
          continue;
        }
        Args a;
        a.hw = hw; a.rows_per_box = v.R; a.T = v.T; a.nload = v.nload;
        a.stage_bytes = v.R * v.T * 4;
        if ((size_t)a.stage_bytes * S > 216 * 1024) continue;
        a.stages = S; a.do_store = do_store;
        a.total = (rows / v.R) * ((hw + v.T - 1) / v.T);
        for (int it = 0; it < 3; ++it) {
          if (it == 1) cudaEventRecord(e0);
          stream_kernel<<<148, 128, (size_t)a.stage_bytes * S>>>(mi, mo, a);
        }
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= 2;
        const double bytes = (double)rows * hw * 4 * (do_store ? 2 : 1);
        printf("R=%3d T=%3d nload=%d promo=%d store=%d S=%d stage=%6d B: %.1f us  %.2f TB/s  %s\n", v.R, v.T, v.nload, v.promo, do_store, S, a.stage_bytes,
               ms * 1e3, bytes / (ms * 1e-3) / 1e12, err == cudaSuccess ? "" : cudaGetErrorString(err));
      }
    }
  }
  return 0;
}
