// Analytic anchors + valid flags + ATSS assignment + pseudo sampling, without ever
// materialising an A x G matrix.
// Reference: ATSSAssigner.assign (task_modules/assigners/atss_assigner.py:74-254),
// bbox_center_distance (:15-36), bbox_overlaps pairwise branch
// (structures/bbox/bbox_overlaps.py:151-193), AnchorGenerator
// (task_modules/prior_generators/anchor_generator.py:161-205,266-301,415-476),
// GFLHead._get_targets_single (dense_heads/gfl_head.py:562-669).
//
// Integer-critical arithmetic uses explicit round-to-nearest intrinsics (no FMA
// contraction) in the reference's operation order.
#include "erd_common.cuh"

namespace erd {

__device__ __forceinline__ float center_distance(float pcx, float pcy, float gcx, float gcy) {
  const float dx = __fsub_rn(pcx, gcx), dy = __fsub_rn(pcy, gcy);
  return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));   // atss_assigner.py:33-34
}

__device__ __forceinline__ float anchor_gt_iou(float ax1, float ay1, float ax2, float ay2, float gx1, float gy1,
                                               float gx2, float gy2) {
  const float area_a = __fmul_rn(__fsub_rn(ax2, ax1), __fsub_rn(ay2, ay1));
  const float area_g = __fmul_rn(__fsub_rn(gx2, gx1), __fsub_rn(gy2, gy1));
  const float w = fmaxf(__fsub_rn(fminf(ax2, gx2), fmaxf(ax1, gx1)), 0.f);
  const float h = fmaxf(__fsub_rn(fminf(ay2, gy2), fmaxf(ay1, gy1)), 0.f);
  const float inter = __fmul_rn(w, h);
  const float uni = fmaxf(__fsub_rn(__fadd_rn(area_a, area_g), inter), 1e-6f);
  return __fdiv_rn(inter, uni);
}

// (distance bits, index in level): ascending order = nearest first, lowest index on ties
__device__ __forceinline__ unsigned long long distance_key(const LevelView& v, int x, int y, float gcx, float gcy) {
  const float d = center_distance((float)(x * v.stride), (float)(y * v.stride), gcx, gcy);
  return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned int)(y * v.W + x);
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t < v ? t : v;
  }
  return v;
}

// One CTA per GT box, one warp per pyramid level.  Each warp selects the (up to) nine
// valid anchors of its level closest to the GT centre, by fp32 distance then lowest index.
// Search runs over a clipped window around the nearest lattice point; a conservative bound
// on everything outside the window either proves the window sufficient or widens the search
// to the whole level.
constexpr int kCandThreads = 32 * kLevels;
constexpr int kWinR = 3;

__global__ void __launch_bounds__(kCandThreads) atss_candidates_kernel(Geo g, Workspace ws,
                                                                        const float* __restrict__ gt_boxes,
                                                                        const int32_t* __restrict__ gt_offsets,
                                                                        const int32_t* __restrict__ pad_hw) {
  const int gid = blockIdx.x;
  // the grid covers shape.total_gt rows, which may be a fixed capacity (a CUDA graph of the step then stays
  // valid from batch to batch); the rows in use are gt_offsets[N]
  if (gid >= gt_offsets[g.n_img]) return;
  const int lane = threadIdx.x & 31, l = threadIdx.x >> 5;
  __shared__ int s_img, s_first, s_pad[2];
  __shared__ int s_idx[kLevels * kTopK];
  __shared__ float s_iou[kLevels * kTopK];
  __shared__ float s_cx[kLevels * kTopK], s_cy[kLevels * kTopK];
  // every load of the prologue is independent, so the CTA pays one memory round trip: each
  // thread tests one image's GT range (the image owning gid publishes itself and its padding)
  const float gx1 = gt_boxes[gid * 4 + 0], gy1 = gt_boxes[gid * 4 + 1];
  const float gx2 = gt_boxes[gid * 4 + 2], gy2 = gt_boxes[gid * 4 + 3];
  for (int i = threadIdx.x; i < g.n_img; i += kCandThreads) {
    const int lo = gt_offsets[i], hi = gt_offsets[i + 1];
    const int ph = pad_hw[i * 2], pw = pad_hw[i * 2 + 1];
    if (lo <= gid && gid < hi) {
      s_img = i;
      s_first = lo;
      s_pad[0] = ph;
      s_pad[1] = pw;
    }
  }
  if (threadIdx.x < kLevels * kTopK) s_idx[threadIdx.x] = -1;
  __syncthreads();
  const int n = s_img;
  const float gcx = __fmul_rn(__fadd_rn(gx1, gx2), 0.5f);   // (x1 + x2) / 2.0, atss_assigner.py:25-26
  const float gcy = __fmul_rn(__fadd_rn(gy1, gy2), 0.5f);
  const LevelView v = level_view(g, l, s_pad[0], s_pad[1]);
  const int nvalid = v.vw * v.vh;
  const int ksel = min(kTopK, nvalid);                      // atss_assigner.py:198
  if (ksel > 0) {
    const float fs = (float)v.stride;
    int x0 = 0, y0 = 0, wx = v.vw, wy = v.vh;
    float outside = INFINITY;   // lower bound of the distance of any valid anchor outside the window
    if (nvalid > 64) {
      const int ix = min(max(__float2int_rn(gcx / fs), 0), v.vw - 1);
      const int iy = min(max(__float2int_rn(gcy / fs), 0), v.vh - 1);
      x0 = max(ix - kWinR, 0);
      y0 = max(iy - kWinR, 0);
      const int x1 = min(ix + kWinR, v.vw - 1), y1 = min(iy + kWinR, v.vh - 1);
      wx = x1 - x0 + 1;
      wy = y1 - y0 + 1;
      if (x0 > 0) outside = fminf(outside, gcx - (float)(x0 - 1) * fs);
      if (x1 < v.vw - 1) outside = fminf(outside, (float)(x1 + 1) * fs - gcx);
      if (y0 > 0) outside = fminf(outside, gcy - (float)(y0 - 1) * fs);
      if (y1 < v.vh - 1) outside = fminf(outside, (float)(y1 + 1) * fs - gcy);
    }
    for (int pass = 0; pass < 2; ++pass) {
      const int npts = wx * wy;
      unsigned long long prev = 0ull;
      if (npts <= 64) {   // the usual case (7x7 window): two keys per lane, computed once
        unsigned long long k0 = ~0ull, k1 = ~0ull;
        if (lane < npts) k0 = distance_key(v, x0 + lane % wx, y0 + lane / wx, gcx, gcy);
        if (lane + 32 < npts) k1 = distance_key(v, x0 + (lane + 32) % wx, y0 + (lane + 32) / wx, gcx, gcy);
        if (k1 < k0) { const unsigned long long t = k0; k0 = k1; k1 = t; }
        for (int r = 0; r < ksel; ++r) {
          const unsigned long long best = warp_min_u64(k0);
          if (k0 == best) { k0 = k1; k1 = ~0ull; }   // keys are unique (they embed the index)
          prev = best;
          if (lane == 0) s_idx[l * kTopK + r] = (int)(unsigned int)(best & 0xffffffffull);
        }
      } else {
        for (int r = 0; r < ksel; ++r) {
          unsigned long long best = ~0ull;
          for (int p = lane; p < npts; p += 32) {
            const unsigned long long key = distance_key(v, x0 + p % wx, y0 + p / wx, gcx, gcy);
            if ((r == 0 || key > prev) && key < best) best = key;
          }
          best = warp_min_u64(best);
          prev = best;
          if (lane == 0) s_idx[l * kTopK + r] = (int)(unsigned int)(best & 0xffffffffull);
        }
      }
      // the window is sufficient when the k-th distance is safely below every outside anchor
      const float dk = __uint_as_float((unsigned int)(prev >> 32));
      if (pass == 1 || dk < outside * 0.9999f) break;
      x0 = 0; y0 = 0; wx = v.vw; wy = v.vh;
    }
    __syncwarp();
    if (lane < ksel) {
      const int idx = s_idx[l * kTopK + lane];
      const int x = idx % v.W, y = idx / v.W;
      const float cx = (float)(x * v.stride), cy = (float)(y * v.stride);
      s_iou[l * kTopK + lane] = anchor_gt_iou(cx - v.half, cy - v.half, cx + v.half, cy + v.half, gx1, gy1, gx2, gy2);
      s_cx[l * kTopK + lane] = cx;
      s_cy[l * kTopK + lane] = cy;
      s_idx[l * kTopK + lane] = v.start + idx;
    }
  }
  __syncthreads();
  if (l != 0) return;
  // mean + std (unbiased) of the candidate IoUs in fp64, rounded once each (atss_assigner.py:207-210)
  const int t0 = lane, t1 = lane + 32;
  const bool h0 = s_idx[t0] >= 0, h1 = t1 < kLevels * kTopK && s_idx[t1] >= 0;
  const double v0 = h0 ? (double)s_iou[t0] : 0.0, v1 = h1 ? (double)s_iou[t1] : 0.0;
  const double cnt = warp_sum((double)(h0 + h1));
  const double mean = warp_sum(v0 + v1) / cnt;
  const double d0 = h0 ? v0 - mean : 0.0, d1 = h1 ? v1 - mean : 0.0;
  const double var = warp_sum(d0 * d0 + d1 * d1) / (cnt - 1.0);   // one candidate -> NaN, as torch.std
  const float thr = __fadd_rn((float)mean, (float)sqrt(var));
  const int glocal = gid - s_first;
#pragma unroll
  for (int rep = 0; rep < 2; ++rep) {
    const int t = rep ? t1 : t0;
    if (!(rep ? h1 : h0)) continue;
    const float iou = s_iou[t];
    const float cx = s_cx[t], cy = s_cy[t];
    const float inside = fminf(fminf(__fsub_rn(cx, gx1), __fsub_rn(cy, gy1)),
                               fminf(__fsub_rn(gx2, cx), __fsub_rn(gy2, cy)));   // :227-231
    if (iou >= thr && inside > 0.01f) {
      // highest IoU wins, first GT on ties (:243 torch.max returns the first maximum)
      const unsigned long long key =
          ((unsigned long long)__float_as_uint(iou) << 32) | (0xffffffffu - (unsigned int)glocal);
      atomicMax(ws.atss_key + (size_t)n * g.A + s_idx[t], key);
    }
  }
}

// Per anchor: decode the argmax table into gt_inds and append positives to the image's list
// (atss_decode_anchor, erd_common.cuh).
__global__ void __launch_bounds__(256) atss_finalize_kernel(Geo g, Workspace ws, const int32_t* __restrict__ pad_hw,
                                                            const int32_t* __restrict__ gt_offsets,
                                                            int32_t* __restrict__ gt_inds,
                                                            int32_t* __restrict__ num_pos) {
  const int n = blockIdx.y;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < g.A) atss_decode_anchor(g, ws, pad_hw, gt_offsets, gt_inds, n, a);
  // The last block to finish publishes the per-image positive counts and re-zeroes the state
  // the step accumulates into, so no memset sits on the step's critical path (the workspace is
  // zero-initialised once, erd_workspace_init).
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(ws.counters + 3, 1u) == gridDim.x * gridDim.y - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int i = threadIdx.x; i < g.n_img; i += blockDim.x) {
    num_pos[i] = ((volatile int*)ws.pos_counter)[i];
    ws.pos_counter[i] = 0;
  }
  if (threadIdx.x == 0) ws.counters[3] = 0u;
}

cudaError_t launch_atss_candidates(const Geo& g, const Workspace& ws, const float* gt_boxes, const int32_t* gt_offsets,
                                   const int32_t* pad_hw, cudaStream_t st) {
  if (g.total_gt > 0)
    ERD_LAUNCH(kKAtssCand, st,
               (atss_candidates_kernel<<<g.total_gt, kCandThreads, 0, st>>>(g, ws, gt_boxes, gt_offsets, pad_hw)));
  return cudaGetLastError();
}

cudaError_t launch_atss(const Geo& g, const Workspace& ws, const float* gt_boxes, const int64_t* gt_labels,
                        const int32_t* gt_offsets, const int32_t* pad_hw, int32_t* gt_inds, int32_t* num_pos,
                        cudaStream_t st) {
  (void)gt_labels;
  cudaError_t e = launch_atss_candidates(g, ws, gt_boxes, gt_offsets, pad_hw, st);
  if (e != cudaSuccess) return e;
  ERD_LAUNCH(kKAtssFin, st,
             (atss_finalize_kernel<<<dim3((g.A + 255) / 256, g.n_img), 256, 0, st>>>(g, ws, pad_hw, gt_offsets, gt_inds, num_pos)));
  return cudaGetLastError();
}

}  // namespace erd
