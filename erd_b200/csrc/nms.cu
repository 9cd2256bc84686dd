// Teacher box decode + class-aware greedy IoU-NMS over the ERS-selected rows of each image.
// Reference call site: GFLHeadIncrementERD.distill_loss_by_image_single
// (dense_heads/gfl_head_increment_erd.py:189-202).  mmcv.ops.batched_nms is third-party
// (mmcv>=2.0.0rc4,<2.1.0, not in the reference tree); its published semantics are followed:
// boxes_for_nms = boxes + class_id * (boxes.max() + 1); visit by descending score; area
// (x2-x1)*(y2-y1); suppress when inter / (area_i + area_j - inter) > thr; kept indices are
// returned in visiting order.  All box arithmetic is IEEE fp32 without contraction.
#include "erd_common.cuh"

namespace erd {

constexpr int kSortThreads = 1024;

// One CTA per image: build boxes, find the coordinate maximum, sort by (score desc, list
// position asc), apply the class offset, write boxes in score order.
__global__ void __launch_bounds__(kSortThreads) nms_sort_kernel(Geo g, Workspace ws,
                                                                const int32_t* __restrict__ box_inds,
                                                                const int32_t* __restrict__ box_count,
                                                                const int32_t* __restrict__ pad_hw, int pow2_cap) {
  extern __shared__ unsigned long long s_key[];
  __shared__ float s_max[kSortThreads / 32];
  const int n = blockIdx.x;
  const int K = box_count[n];
  if (K == 0) return;
  int P = 1;
  while (P < K) P <<= 1;
  const int pad_h = pad_hw[n * 2], pad_w = pad_hw[n * 2 + 1];
  const int32_t* list = box_inds + (size_t)n * g.sel_cap;
  float4* raw = ws.nms_raw + (size_t)n * g.sel_cap;
  int* cls = ws.nms_cls + (size_t)n * g.sel_cap;
  float mx = -INFINITY;
  for (int r = threadIdx.x; r < P; r += kSortThreads) {
    unsigned long long key = ~0ull;
    if (r < K) {
      const int a = list[r];
      const size_t ga = (size_t)n * g.A + a;
      const int l = level_of_anchor(g, a);
      const int rel = a - g.start[l];
      const int x = rel % g.w[l], y = rel / g.w[l];
      const int s = g.stride[l];
      // anchors handed to the distillation step were unmap()ed with fill 0
      // (gfl_head.py:660): anchors outside pad_shape sit at the origin.
      const bool valid = x < min((pad_w + s - 1) / s, g.w[l]) && y < min((pad_h + s - 1) / s, g.h[l]);
      const float cx = valid ? (float)(x * s) : 0.f, cy = valid ? (float)(y * s) : 0.f;
      const float4 d = ws.t_dist[ga];   // bin units used as pixels (no * stride), :189-192
      const float4 b = make_float4(__fsub_rn(cx, d.x), __fsub_rn(cy, d.y), __fadd_rn(cx, d.z), __fadd_rn(cy, d.w));
      raw[r] = b;
      cls[r] = ws.t_arg[ga];
      mx = fmaxf(mx, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
      key = ((unsigned long long)(~__float_as_uint(ws.t_m[ga])) << 32) | (unsigned int)r;
    }
    s_key[r] = key;
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
  __syncthreads();
  // bitonic sort, ascending on key = descending score, ties by list position
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += kSortThreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = s_key[i], b = s_key[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { s_key[i] = b; s_key[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  float maxc = s_max[0];
  for (int w = 1; w < kSortThreads / 32; ++w) maxc = fmaxf(maxc, s_max[w]);
  const float unit = __fadd_rn(maxc, 1.0f);
  float4* sorted = ws.nms_box + (size_t)n * g.sel_cap;
  int* order = ws.nms_order + (size_t)n * g.sel_cap;
  for (int i = threadIdx.x; i < K; i += kSortThreads) {
    const int r = (int)(unsigned int)(s_key[i] & 0xffffffffull);
    const float4 b = raw[r];
    const float off = __fmul_rn((float)cls[r], unit);
    sorted[i] = make_float4(__fadd_rn(b.x, off), __fadd_rn(b.y, off), __fadd_rn(b.z, off), __fadd_rn(b.w, off));
    order[i] = r;
  }
  (void)pow2_cap;
}

// Suppression bit matrix over score-ordered boxes: bit j of mask[i][cb] is set when box
// cb*64+j (j > i) overlaps box i above the threshold.  Only tiles with cb >= rb are written.
__global__ void __launch_bounds__(64) nms_mask_kernel(Geo g, Workspace ws, const int32_t* __restrict__ box_count,
                                                      float iou_thr) {
  const int n = blockIdx.y;
  const int K = box_count[n];
  const int W = (K + 63) >> 6;
  const int Wcap = nms_words(g.sel_cap);
  const float4* boxes = ws.nms_box + (size_t)n * g.sel_cap;
  unsigned long long* mask = ws.nms_mask + (size_t)n * g.sel_cap * Wcap;
  __shared__ float4 s_col[64];
  __shared__ float s_area[64];
  const int ntile = W * (W + 1) / 2;
  for (int t = blockIdx.x; t < ntile; t += gridDim.x) {
    // t -> (rb, cb) with cb >= rb, row-major over the upper triangle
    int rb = 0, rem = t;
    while (rem >= W - rb) { rem -= W - rb; ++rb; }
    const int cb = rb + rem;
    __syncthreads();
    const int cj = cb * 64 + threadIdx.x;
    if (cj < K) {
      const float4 b = boxes[cj];
      s_col[threadIdx.x] = b;
      s_area[threadIdx.x] = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    }
    __syncthreads();
    const int i = rb * 64 + threadIdx.x;
    if (i < K) {
      const float4 a = boxes[i];
      const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
      unsigned long long bits = 0ull;
      const int ncol = min(64, K - cb * 64);
      const int j0 = (rb == cb) ? threadIdx.x + 1 : 0;
      for (int j = j0; j < ncol; ++j) {
        const float4 b = s_col[j];
        const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
        const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
        const float inter = __fmul_rn(w, h);
        const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, s_area[j]), inter));
        if (ovr > iou_thr) bits |= 1ull << j;
      }
      mask[(size_t)i * Wcap + cb] = bits;
    }
  }
}

// Greedy pass, one warp per image.  64-box chunks: the in-chunk dependency is resolved on
// the diagonal words held in registers; rows of surviving boxes are then OR-ed into the
// running removal mask of the later chunks (lane-strided words in shared memory).
__global__ void __launch_bounds__(32) nms_scan_kernel(Geo g, Workspace ws, const int32_t* __restrict__ box_count,
                                                      int32_t* __restrict__ keep, int32_t* __restrict__ keep_count) {
  extern __shared__ unsigned long long s_remv[];
  const int n = blockIdx.x;
  const int lane = threadIdx.x;
  const int K = box_count[n];
  const int W = (K + 63) >> 6;
  const int Wcap = nms_words(g.sel_cap);
  const unsigned long long* mask = ws.nms_mask + (size_t)n * g.sel_cap * Wcap;
  const int* order = ws.nms_order + (size_t)n * g.sel_cap;
  int32_t* out = keep + (size_t)n * g.sel_cap;
  for (int w = lane; w < W; w += 32) s_remv[w] = 0ull;
  __syncwarp();
  int nkeep = 0;
  for (int c = 0; c < W; ++c) {
    const int rows = min(64, K - c * 64);
    // diagonal words of this chunk: lane holds rows lane and lane + 32
    const unsigned long long d_lo = lane < rows ? mask[(size_t)(c * 64 + lane) * Wcap + c] : 0ull;
    const unsigned long long d_hi = lane + 32 < rows ? mask[(size_t)(c * 64 + lane + 32) * Wcap + c] : 0ull;
    unsigned long long alive = ~s_remv[c];
    if (rows < 64) alive &= (1ull << rows) - 1ull;
#pragma unroll 4
    for (int t = 0; t < 64; ++t) {
      const unsigned long long d = __shfl_sync(0xffffffffu, t < 32 ? d_lo : d_hi, t & 31);
      if ((alive >> t) & 1ull) alive &= ~d;
    }
    // emit survivors in score order
    const unsigned long long below_lo = alive & ((1ull << lane) - 1ull);
    const unsigned long long below_hi = alive & ((1ull << (lane + 32)) - 1ull);
    if ((alive >> lane) & 1ull) out[nkeep + __popcll(below_lo)] = order[c * 64 + lane];
    if ((alive >> (lane + 32)) & 1ull) out[nkeep + __popcll(below_hi)] = order[c * 64 + lane + 32];
    nkeep += __popcll(alive);
    // fold the survivors' rows into the removal mask of later chunks
    unsigned long long todo = alive;
    while (todo) {
      const int t = __ffsll((long long)todo) - 1;
      todo &= todo - 1ull;
      const unsigned long long* row = mask + (size_t)(c * 64 + t) * Wcap;
      for (int w = c + 1 + lane; w < W; w += 32) s_remv[w] |= row[w];
    }
    __syncwarp();
  }
  if (lane == 0) keep_count[n] = nkeep;
}

cudaError_t launch_nms(const Geo& g, const Workspace& ws, const int32_t* box_inds, const int32_t* box_count,
                       const int32_t* pad_hw, float iou_thr, int32_t* keep, int32_t* keep_count, cudaStream_t st) {
  int P = 1;
  while (P < g.sel_cap) P <<= 1;
  const size_t sort_smem = sizeof(unsigned long long) * (size_t)P;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(nms_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done = true;
  }
  if (sort_smem > 200 * 1024) return cudaErrorInvalidValue;
  ERD_LAUNCH(kKNmsSort, st,
             (nms_sort_kernel<<<g.n_img, kSortThreads, sort_smem, st>>>(g, ws, box_inds, box_count, pad_hw, P)));
  ERD_LAUNCH(kKNmsMask, st, (nms_mask_kernel<<<dim3(64, g.n_img), 64, 0, st>>>(g, ws, box_count, iou_thr)));
  const size_t scan_smem = sizeof(unsigned long long) * (size_t)nms_words(g.sel_cap);
  ERD_LAUNCH(kKNmsScan, st, (nms_scan_kernel<<<g.n_img, 32, scan_smem, st>>>(g, ws, box_count, keep, keep_count)));
  return cudaGetLastError();
}

}  // namespace erd
