// Elastic Response Selection, part 2: the per-anchor flags and the ordered index lists the
// reference's sel_pos returns (and the teacher NMS consumes).  Part 1 -- the streaming pass over
// the teacher logits, the per-anchor cache and the thresholds -- is teacher.cu.
// Reference: GFLIncrementERD.sel_pos / sel_pos_single
// (mmdet/models/detectors/gfl_increment_erd.py:143-200).
#include <cstdlib>
#include "erd_common.cuh"

namespace erd {

// ----------------------------------------------------------------------------- flags + counts
// The part of the selection the student pass needs -- per-anchor flag byte (bit 0: class-response
// row, bit 1: box candidate) and the per-image counts K_cls / K_bbox -- in one short launch right
// behind the teacher pass; the ordered lists follow off the critical path.
constexpr int kFlagThreads = 256;
constexpr int kFlagPer = 4;

__global__ void __launch_bounds__(kFlagThreads) ers_flags_kernel(Geo g, Workspace ws, int tiles, float* __restrict__ thr_out,
                                                                 uint8_t* __restrict__ sel_flags, int32_t* cls_count,
                                                                 int32_t* box_count) {
  const int n = blockIdx.y;
  // thr = mean + 2 std (unbiased) of the image (gfl_increment_erd.py:149,157) from the teacher pass's
  // per-tile fp64 sums.  Every block of the image reduces them itself (22 KB out of L2) with the same
  // thread -> element mapping and the same tree, so all blocks hold identical bits and the result is
  // reproducible run to run; block 0 publishes it for the lists kernel and the caller.
  __shared__ double s_red[kFlagThreads / 32];
  __shared__ double s_sum[4];
  __shared__ float s_thr[2];
  {
    const double* p = ws.ers_part + (size_t)n * tiles * 4;
    const int total = tiles * 4;
    double acc = 0.0;   // thread t sums elements t, t + 256, ...: all of component t % 4
    for (int e0 = threadIdx.x; e0 < total; e0 += kFlagThreads * 8) {
      double v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int e = e0 + kFlagThreads * j;
        v[j] = e < total ? p[e] : 0.0;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += v[j];
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 8);
    acc += __shfl_xor_sync(0xffffffffu, acc, 16);   // lanes 0..3: the warp's sums of m, m^2, u, u^2
    for (int comp = 0; comp < 4; ++comp) {
      if ((threadIdx.x & 31) == comp) s_red[threadIdx.x >> 5] = acc;
      __syncthreads();
      if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kFlagThreads / 32; ++w) t += s_red[w];
        s_sum[comp] = t;
      }
      __syncthreads();
    }
    if (threadIdx.x < 2) {
      const double s1 = s_sum[threadIdx.x * 2], s2 = s_sum[threadIdx.x * 2 + 1];
      const double An = (double)g.A;
      const double mean = s1 / An;
      double var = (s2 - s1 * s1 / An) / (An - 1.0);   // A == 1 -> NaN, as torch.std
      if (var < 0.0) var = 0.0;
      const float t = __fadd_rn((float)mean, __fmul_rn(2.0f, (float)sqrt(var)));
      s_thr[threadIdx.x] = t;
      if (blockIdx.x == 0) {
        thr_out[n * 2 + threadIdx.x] = t;
        // what the NEXT call's teacher pass stashes by: the lowest mean + 1.7 std over this call's images
        // (kept as ~ordered_bits, so 0 = nothing yet and "lowest" is an atomicMax)
        atomicMax(ws.pthr_state + 2 + threadIdx.x, ~ordered_bits((float)(mean + 1.7 * sqrt(var))));
      }
    }
    __syncthreads();
  }
  const float thr_c = s_thr[0], thr_b = s_thr[1];
  const int a0 = (blockIdx.x * kFlagThreads + threadIdx.x) * kFlagPer;
  const float* m = ws.t_m + (size_t)n * g.A;
  const float* u = ws.t_u + (size_t)n * g.A;
  int nc = 0, nb = 0;
  unsigned packed = 0u;
#pragma unroll
  for (int k = 0; k < kFlagPer; ++k) {
    const bool in = a0 + k < g.A;
    const bool c = in && m[a0 + k] > thr_c, b = in && u[a0 + k] > thr_b;   // strict (gfl_increment_erd.py:150,158)
    nc += c;
    nb += b;
    packed |= (unsigned)((c ? 1 : 0) | (b ? 2 : 0)) << (8 * k);
  }
  uint8_t* out = sel_flags + (size_t)n * g.A + a0;
  if (a0 + kFlagPer <= g.A && (((uintptr_t)out) & 3) == 0) {
    *reinterpret_cast<unsigned*>(out) = packed;
  } else {
    for (int k = 0; k < kFlagPer && a0 + k < g.A; ++k) out[k] = (uint8_t)(packed >> (8 * k));
  }
  __shared__ int s_c[kFlagThreads / 32], s_b[kFlagThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nc += __shfl_xor_sync(0xffffffffu, nc, o);
    nb += __shfl_xor_sync(0xffffffffu, nb, o);
  }
  if ((threadIdx.x & 31) == 0) { s_c[threadIdx.x >> 5] = nc; s_b[threadIdx.x >> 5] = nb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int tc = 0, tb = 0;
    for (int w = 0; w < kFlagThreads / 32; ++w) { tc += s_c[w]; tb += s_b[w]; }
    if (tc) atomicAdd(cls_count + n, tc);
    if (tb) atomicAdd(box_count + n, tb);
    // the grid's last block hands the collected provisional thresholds to the next call and starts a new collection
    __threadfence();
    if (atomicAdd(ws.counters + 4, 1u) == gridDim.x * gridDim.y - 1) {
      __threadfence();
      for (int i = 0; i < 2; ++i) {
        ws.pthr_state[i] = ((volatile unsigned int*)ws.pthr_state)[2 + i];
        ws.pthr_state[2 + i] = 0u;
      }
      ws.counters[4] = 0u;
    }
  }
}

// ----------------------------------------------------------------------------- ordered lists
// Rows strictly above the image's thresholds, in ascending anchor order (what nonzero() yields,
// gfl_increment_erd.py:150-151,158-159).  Off the step's critical path: the student pass derives
// the same flags itself from the teacher pass's cache and thresholds.  Each CTA owns a 2048-anchor chunk
// of one image; it recounts the flags of the anchors before its chunk from the L2-resident
// cache instead of waiting on a cross-CTA prefix, so one launch suffices.
constexpr int kSelThreads = 1024;
constexpr int kSelPer = 2;
constexpr int kSelChunk = kSelThreads * kSelPer;

__global__ void __launch_bounds__(kSelThreads) ers_select_kernel(Geo g, Workspace ws, int32_t* cls_inds,
                                                                 int32_t* cls_count, int32_t* box_inds,
                                                                 int32_t* box_count, const float* __restrict__ thr_in,
                                                                 uint8_t* __restrict__ sel_flags) {
  const int n = blockIdx.y;
  const int chunk0 = blockIdx.x * kSelChunk;
  __shared__ int s_warp[2][kSelThreads / 32];
  __shared__ int s_base[2];
  // the thresholds the teacher pass published (the student pass derives its roles from the same
  // numbers: recomputing them here with another summation order could differ in the last bit)
  const float thr_c = thr_in[n * 2], thr_b = thr_in[n * 2 + 1];
  const float* m = ws.t_m + (size_t)n * g.A;
  const float* u = ws.t_u + (size_t)n * g.A;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // selected rows before this chunk
  int pc = 0, pb = 0;
  if ((((uintptr_t)m | (uintptr_t)u) & 15) == 0) {   // chunk0 is a multiple of 2048: 16 B loads, 4 in flight
    const float4* m4 = reinterpret_cast<const float4*>(m);
    const float4* u4 = reinterpret_cast<const float4*>(u);
#pragma unroll 4
    for (int i = threadIdx.x; i < chunk0 / 4; i += kSelThreads) {
      const float4 a = m4[i], b = u4[i];
      pc += (a.x > thr_c) + (a.y > thr_c) + (a.z > thr_c) + (a.w > thr_c);
      pb += (b.x > thr_b) + (b.y > thr_b) + (b.z > thr_b) + (b.w > thr_b);
    }
  } else {
#pragma unroll 8
    for (int a = threadIdx.x; a < chunk0; a += kSelThreads) {
      pc += m[a] > thr_c;
      pb += u[a] > thr_b;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    pc += __shfl_xor_sync(0xffffffffu, pc, o);
    pb += __shfl_xor_sync(0xffffffffu, pb, o);
  }
  if (lane == 0) { s_warp[0][warp] = pc; s_warp[1][warp] = pb; }
  __syncthreads();
  if (threadIdx.x < 2) {
    int t = 0;
    for (int w = 0; w < kSelThreads / 32; ++w) t += s_warp[threadIdx.x][w];
    s_base[threadIdx.x] = t;
  }
  __syncthreads();
  // this chunk: 8 consecutive anchors per thread, block-wide exclusive scan of the counts
  const int a0 = chunk0 + threadIdx.x * kSelPer;
  unsigned fc = 0, fb = 0;
#pragma unroll
  for (int k = 0; k < kSelPer; ++k) {
    const bool in = a0 + k < g.A;
    const bool c = in && (m[a0 + k] > thr_c), b = in && (u[a0 + k] > thr_b);
    fc |= (unsigned)c << k;
    fb |= (unsigned)b << k;
    if (in) sel_flags[(size_t)n * g.A + a0 + k] = (uint8_t)((c ? 1 : 0) | (b ? 2 : 0));
  }
  const int nc = __popc(fc), nb = __popc(fb);
  int ic = nc, ib = nb;   // inclusive warp scans
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int tc = __shfl_up_sync(0xffffffffu, ic, o);
    const int tb = __shfl_up_sync(0xffffffffu, ib, o);
    if (lane >= o) { ic += tc; ib += tb; }
  }
  __syncthreads();
  if (lane == 31) { s_warp[0][warp] = ic; s_warp[1][warp] = ib; }
  __syncthreads();
  int oc = s_base[0] + ic - nc, ob = s_base[1] + ib - nb;
  int totc = 0, totb = 0;
  for (int w = 0; w < kSelThreads / 32; ++w) {
    if (w < warp) { oc += s_warp[0][w]; ob += s_warp[1][w]; }
    totc += s_warp[0][w];
    totb += s_warp[1][w];
  }
  int32_t* out_c = cls_inds + (size_t)n * g.sel_cap;
  int32_t* out_b = box_inds + (size_t)n * g.sel_cap;
#pragma unroll
  for (int k = 0; k < kSelPer; ++k) {
    if (fc & (1u << k)) out_c[oc++] = a0 + k;
    if (fb & (1u << k)) out_b[ob++] = a0 + k;
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
    cls_count[n] = s_base[0] + totc;
    box_count[n] = s_base[1] + totb;
  }
}

cudaError_t launch_ers_flags(const Geo& g, const Workspace& ws, int tiles_per_img, float* thr, uint8_t* sel_flags,
                             int32_t* cls_count, int32_t* box_count, cudaStream_t st) {
  const int per_block = kFlagThreads * kFlagPer;
  ERD_LAUNCH(kKErsFlags, st,
             (ers_flags_kernel<<<dim3((g.A + per_block - 1) / per_block, g.n_img), kFlagThreads, 0, st>>>(
                 g, ws, tiles_per_img, thr, sel_flags, cls_count, box_count)));
  return cudaGetLastError();
}

cudaError_t launch_ers_lists(const Geo& g, const Workspace& ws, int32_t* cls_inds, int32_t* cls_count, int32_t* box_inds,
                             int32_t* box_count, const float* thr, uint8_t* sel_flags, cudaStream_t st) {
  ERD_LAUNCH(kKErsSelect, st,
             (ers_select_kernel<<<dim3((g.A + kSelChunk - 1) / kSelChunk, g.n_img), kSelThreads, 0, st>>>(
                 g, ws, cls_inds, cls_count, box_inds, box_count, thr, sel_flags)));
  return cudaGetLastError();
}

cudaError_t launch_ers(const Geo& g, const Workspace& ws, const Ptr5& t_cls, const Ptr5& t_box, int32_t* cls_inds,
                       int32_t* cls_count, int32_t* box_inds, int32_t* box_count, float* thr, uint8_t* sel_flags,
                       cudaStream_t st) {
  int tiles = 0;
  cudaError_t e = launch_teacher_pass(g, ws, t_cls, t_box, cls_count, box_count, &tiles, st);
  if (e == cudaSuccess) e = launch_ers_flags(g, ws, tiles, thr, sel_flags, cls_count, box_count, st);
  if (e != cudaSuccess) return e;
  return launch_ers_lists(g, ws, cls_inds, cls_count, box_inds, box_count, thr, sel_flags, st);
}

}  // namespace erd
