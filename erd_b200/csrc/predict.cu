// Inference post-process of the GFL head on the device: per level sigmoid scores above score_thr,
// the nms_pre best of them, softmax-integral box decode (x stride) clamped to the image, the
// min_bbox_size filter, class-aware IoU-NMS and the max_per_img best survivors.
// Reference: GFLHead._predict_by_feat_single (mmdet/models/dense_heads/gfl_head.py:408-502),
// filter_scores_and_topk (mmdet/models/utils/misc.py:308-354), BaseDenseHead._bbox_post_process
// (mmdet/models/dense_heads/base_dense_head.py:424-486), distance2bbox with max_shape
// (mmdet/structures/bbox/transforms.py:147-182); mmcv.ops.batched_nms as in nms.cu.
//
// Kernels
//   predict_collect   dense pass over the class logits (coalesced, one anchor per thread): every
//                     (anchor, class) whose sigmoid clears score_thr is appended to the list of its
//                     (image, level) as a 64-bit key  ordered(logit) << 32 | ~flat_index .  sigmoid is
//                     monotone, so sorting keys descending is "descending score, earlier flat index
//                     (anchor * C + class, the reference's nonzero() order) first on ties".
//   predict_topk      one CTA per (image, level): the nms_pre largest keys -- a shared-memory bitonic
//                     sort when the list fits (<= 4096), else an 11-bit radix select over the list
//                     first -- then the decode of those candidates in score order.
//   predict_merge     one CTA per image: ordered concatenation of the five level lists without the
//                     boxes the min_bbox_size filter drops, the class offsets of batched_nms
//                     (label * (max coordinate + 1)), clean predecessor rows for the NMS.
//   nms_mask / nms_resolve / nms_order   the loss path's NMS kernels (nms.cu), unchanged.
//   predict_output    the first max_per_img survivors in score order -> dets (x1, y1, x2, y2, score), labels.
#include <string.h>

#include "erd_common.cuh"

namespace erd {

cudaError_t launch_nms_prepared(const Geo& g, const Workspace& ws, const int32_t* box_inds, const int32_t* box_count,
                                float iou_thr, int32_t* keep, int32_t* keep_count, uint8_t* sel_flags, cudaStream_t st,
                                cudaEvent_t resolved);   // nms.cu

constexpr int kSortCap = 4096;   // keys one CTA sorts in shared memory
constexpr int kRadixBits = 11;   // bits per round of the radix select in front of the sort

struct PredictWs {
  unsigned long long* cand;   // [N][A * C] candidate keys, level l of image n at (n * A + start[l]) * C
  int* cand_count;            // [N][L] (zeroed at the start of every call)
  float4* lvl_box;            // [N][L][nms_pre] decoded, clamped boxes in score order
  float* lvl_score;           // [N][L][nms_pre]
  int* lvl_label;             // [N][L][nms_pre]
  int* lvl_count;             // [N][L] candidates kept per level = min(nms_pre, count)
  float4* out_box;            // [N][cap] boxes handed to the NMS, without the class offsets
  int* out_label;             // [N][cap]
  int32_t* box_inds;          // [N][cap] identity (the NMS kernels address flags through it)
  int32_t* box_count;         // [N]
  int32_t* keep;              // [N][cap] survivors as list positions, score order
  int32_t* keep_count;        // [N]
  uint8_t* flags;             // [N][cap] scratch for the NMS kernels' survivor marks
};

struct PredictArgs {
  Ptr5 cls, box;
  const int32_t* img_hw;       // (N,2) clamp limits (h, w)
  const float* inv_scale;      // (N,2) 1 / scale_factor (w, h) or NULL: rescale=False
  int nms_pre, max_per_img;
  float score_thr, min_size;   // min_size < 0: no size filter
  int lvl_tile_start[kLevels + 1];
};

// ----------------------------------------------------------------------------- collect
constexpr int kCollectThreads = 128;

__global__ void __launch_bounds__(kCollectThreads) predict_collect_kernel(Geo g, PredictWs pw, PredictArgs A) {
  const int n = blockIdx.y;
  int l = 0;
#pragma unroll
  for (int i = 1; i < kLevels; ++i) l += ((int)blockIdx.x >= A.lvl_tile_start[i]) ? 1 : 0;
  const int hw = ((int)blockIdx.x - A.lvl_tile_start[l]) * kCollectThreads + (int)threadIdx.x;
  const int HW = g.hw[l], C = g.C;
  const bool in = hw < HW;
  const float* p = A.cls.p[l] + (size_t)n * C * HW + hw;
  unsigned long long* list = pw.cand + ((size_t)n * g.A + g.start[l]) * C;
  int* count = pw.cand_count + n * kLevels + l;
  const int lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < C; c0 += 8) {
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = (in && c0 + j < C) ? __ldg(p + (size_t)(c0 + j) * HW) : -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool hit = sigmoid_ref(x[j]) > A.score_thr;                              // misc.py:333 (strict)
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m == 0u) continue;   // warp-uniform
      int base = 0;
      if (lane == 0) base = atomicAdd(count, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (hit) {
        const unsigned int flat = (unsigned int)hw * (unsigned int)C + (unsigned int)(c0 + j);   // row-major (anchor, class)
        list[base + __popc(m & ((1u << lane) - 1u))] = ((unsigned long long)ordered_bits(x[j]) << 32) | (0xffffffffu - flat);
      }
    }
  }
}

// ----------------------------------------------------------------------------- top-k + decode
constexpr int kTopkThreads = 1024;

__device__ __forceinline__ void bitonic_desc(unsigned long long* s, int P) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += kTopkThreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = s[i], b = s[ixj];
          const bool up = (i & k) == 0;
          if ((a < b) == up) { s[i] = b; s[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(kTopkThreads) predict_topk_kernel(Geo g, PredictWs pw, PredictArgs A) {
  const int l = blockIdx.x, n = blockIdx.y;
  const int C = g.C, HW = g.hw[l];
  __shared__ unsigned long long s_key[kSortCap];
  __shared__ int s_hist[1 << kRadixBits];
  __shared__ unsigned long long s_lower;
  __shared__ int s_take;
  const int count = pw.cand_count[n * kLevels + l];
  const unsigned long long* list = pw.cand + ((size_t)n * g.A + g.start[l]) * C;
  const int K = min(A.nms_pre, count);                                                // misc.py:338
  // ---- the keys that can be among the K largest, at most kSortCap of them, into shared memory
  unsigned long long lower = 0ull;   // keys >= lower are gathered
  if (count > kSortCap) {
    // radix select, kRadixBits per round from the top: `prefix` is the value of the bits fixed so far, `above`
    // the number of keys above the prefix's range.  Stops as soon as the keys >= the bound fit the sort.
    unsigned long long prefix = 0ull;
    int bits = 0, above = 0;
    for (;;) {
      const int width = min(kRadixBits, 64 - bits);
      const int shift = 64 - bits - width;
      for (int i = threadIdx.x; i < (1 << kRadixBits); i += kTopkThreads) s_hist[i] = 0;
      __syncthreads();
      for (int i = threadIdx.x; i < count; i += kTopkThreads) {
        const unsigned long long key = list[i];
        if (bits == 0 || (key >> (64 - bits)) == prefix) atomicAdd(&s_hist[(int)((key >> shift) & ((1ull << width) - 1ull))], 1);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int cum = above, b = (1 << width) - 1;
        for (; b > 0; --b) {   // the bin in which the K-th largest key lies
          if (cum + s_hist[b] >= K) break;
          cum += s_hist[b];
        }
        s_lower = ((prefix << width) | (unsigned long long)b) << shift;
        s_take = cum + s_hist[b];   // keys >= s_lower
        s_hist[0] = cum;            // keys above bin b (read back below)
      }
      __syncthreads();
      lower = s_lower;
      const int ge = s_take;
      if (ge <= kSortCap || bits + width >= 64) break;
      above = s_hist[0];
      prefix = lower >> shift;
      bits += width;
      __syncthreads();
    }
  }
  __shared__ int s_n;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < count; i += kTopkThreads) {
    const unsigned long long key = list[i];
    if (key >= lower) {
      const int pos = atomicAdd(&s_n, 1);
      if (pos < kSortCap) s_key[pos] = key;
    }
  }
  __syncthreads();
  const int got = min(s_n, kSortCap);
  int P = 1;
  while (P < got) P <<= 1;
  for (int i = got + threadIdx.x; i < P; i += kTopkThreads) s_key[i] = 0ull;
  __syncthreads();
  bitonic_desc(s_key, P);
  if (threadIdx.x == 0) pw.lvl_count[n * kLevels + l] = K;
  // ---- decode the K best in score order (gfl_head.py:466-467 via transforms.py:147-182)
  const float fs = (float)g.stride[l];
  const float lim_h = (float)A.img_hw[n * 2], lim_w = (float)A.img_hw[n * 2 + 1];
  const size_t o = ((size_t)n * kLevels + l) * A.nms_pre;
  for (int k = threadIdx.x; k < K; k += kTopkThreads) {
    const unsigned long long key = s_key[k];
    const unsigned int flat = 0xffffffffu - (unsigned int)(key & 0xffffffffull);
    const int a = (int)(flat / (unsigned int)C), label = (int)(flat % (unsigned int)C);
    const float score = sigmoid_ref(from_ordered_bits((unsigned int)(key >> 32)));
    const float* bp = A.box.p[l] + (size_t)n * kBoxCh * HW + a;
    float d[4];
#pragma unroll
    for (int sd = 0; sd < 4; ++sd) {
      float z[kBins];
#pragma unroll
      for (int j = 0; j < kBins; ++j) z[j] = __ldg(bp + (size_t)(sd * kBins + j) * HW);
      float mx = z[0];
#pragma unroll
      for (int j = 1; j < kBins; ++j) mx = fmaxf(mx, z[j]);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < kBins; ++j) {
        z[j] = expf(z[j] - mx);
        sum += z[j];
      }
      float e = 0.f;
#pragma unroll
      for (int j = 0; j < kBins; ++j) e = fmaf(__fdiv_rn(z[j], sum), (float)j, e);   // softmax, then the dot with 0..reg_max (:48-62)
      d[sd] = e * fs;
    }
    const float px = (float)((a % g.w[l]) * g.stride[l]), py = (float)((a / g.w[l]) * g.stride[l]);
    float4 b = make_float4(px - d[0], py - d[1], px + d[2], py + d[3]);
    b.x = fminf(fmaxf(b.x, 0.f), lim_w);   // clamp(min=0, max=img_shape) (transforms.py:180-181)
    b.z = fminf(fmaxf(b.z, 0.f), lim_w);
    b.y = fminf(fmaxf(b.y, 0.f), lim_h);
    b.w = fminf(fmaxf(b.w, 0.f), lim_h);
    if (A.inv_scale) {   // rescale=True: boxes * (1 / scale_factor) (base_dense_head.py:458-461)
      const float sx = A.inv_scale[n * 2], sy = A.inv_scale[n * 2 + 1];
      b.x *= sx; b.z *= sx; b.y *= sy; b.w *= sy;
    }
    pw.lvl_box[o + k] = b;
    pw.lvl_score[o + k] = score;
    pw.lvl_label[o + k] = label;
  }
}

// ----------------------------------------------------------------------------- merge the levels
constexpr int kMergeThreads = 1024;

__global__ void __launch_bounds__(kMergeThreads) predict_merge_kernel(Geo g, Workspace ws, PredictWs pw, PredictArgs A) {
  const int n = blockIdx.x;
  __shared__ int s_warp[kMergeThreads / 32];
  __shared__ float s_max[kMergeThreads / 32];
  __shared__ int s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cap = g.sel_cap;
  float4* nbox = ws.nms_box + (size_t)n * cap;
  float* nscore = ws.nms_score + (size_t)n * cap;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  float mx = -INFINITY;
  // ordered concatenation (level order, score order inside a level) of the boxes that pass the size filter
  for (int l = 0; l < kLevels; ++l) {
    const int K = pw.lvl_count[n * kLevels + l];
    const size_t o = ((size_t)n * kLevels + l) * A.nms_pre;
    for (int k0 = 0; k0 < K; k0 += kMergeThreads) {
      const int k = k0 + threadIdx.x;
      bool ok = false;
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < K) {
        b = pw.lvl_box[o + k];
        ok = A.min_size < 0.f || ((b.z - b.x) > A.min_size && (b.w - b.y) > A.min_size);   // base_dense_head.py:470-474
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) s_warp[warp] = __popc(m);
      __syncthreads();
      int before = s_base;
      for (int w = 0; w < warp; ++w) before += s_warp[w];
      if (ok) {
        const int pos = before + __popc(m & ((1u << lane) - 1u));
        pw.out_box[(size_t)n * cap + pos] = b;
        pw.out_label[(size_t)n * cap + pos] = pw.lvl_label[o + k];
        nscore[pos] = pw.lvl_score[o + k];
        pw.box_inds[(size_t)n * cap + pos] = pos;
        mx = fmaxf(mx, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < kMergeThreads / 32; ++w) tot += s_warp[w];
        s_base += tot;
      }
      __syncthreads();
    }
  }
  const int K = s_base;
  if (threadIdx.x == 0) pw.box_count[n] = K;
  // class offsets of batched_nms: boxes + label * (boxes.max() + 1)
  mx = warp_max(mx);
  if (lane == 0) s_max[warp] = mx;
  __syncthreads();
  float maxc = s_max[0];
  for (int w = 1; w < kMergeThreads / 32; ++w) maxc = fmaxf(maxc, s_max[w]);
  const float unit = __fadd_rn(maxc, 1.0f);
  for (int r = threadIdx.x; r < K; r += kMergeThreads) {
    const float4 b = pw.out_box[(size_t)n * cap + r];
    const float off = __fmul_rn((float)pw.out_label[(size_t)n * cap + r], unit);
    nbox[r] = make_float4(__fadd_rn(b.x, off), __fadd_rn(b.y, off), __fadd_rn(b.z, off), __fadd_rn(b.w, off));
  }
  // clean predecessor words / map of the K rows in use (what nms_prep does for the loss path)
  const int W = (K + 63) >> 6;
  const int Wcap = nms_words(cap), NZ = nms_nz_words(cap);
  unsigned long long* pred = ws.nms_mask + (size_t)n * cap * Wcap;
  unsigned long long* nz = ws.nms_nz + (size_t)n * cap * NZ;
  for (int i = threadIdx.x; i < K * W; i += kMergeThreads) pred[(size_t)(i / W) * Wcap + (i % W)] = 0ull;
  for (int i = threadIdx.x; i < K * NZ; i += kMergeThreads) nz[i] = 0ull;
}

// ----------------------------------------------------------------------------- output
__global__ void predict_output_kernel(Geo g, Workspace ws, PredictWs pw, PredictArgs A, float* __restrict__ dets,
                                      int32_t* __restrict__ labels, int32_t* __restrict__ num_dets) {
  const int n = blockIdx.x;
  const int cap = g.sel_cap;
  const int m = min(pw.keep_count[n], A.max_per_img);                                 // base_dense_head.py:484
  for (int i = threadIdx.x; i < A.max_per_img; i += blockDim.x) {
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    float s = 0.f;
    int lab = -1;
    if (i < m) {
      const int pos = pw.keep[(size_t)n * cap + i];
      b = pw.out_box[(size_t)n * cap + pos];
      s = ws.nms_score[(size_t)n * cap + pos];
      lab = pw.out_label[(size_t)n * cap + pos];
    }
    float* d = dets + ((size_t)n * A.max_per_img + i) * 5;
    d[0] = b.x; d[1] = b.y; d[2] = b.z; d[3] = b.w; d[4] = s;
    labels[(size_t)n * A.max_per_img + i] = lab;
  }
  if (threadIdx.x == 0) num_dets[n] = m;
}

// ----------------------------------------------------------------------------- host side
static size_t p_align(size_t x) { return (x + 255) & ~(size_t)255; }

// carve the predict workspace; returns its size.  `g.sel_cap` must already be levels * nms_pre.
static size_t carve_predict(const Geo& g, int nms_pre, void* base, PredictWs* pw, Workspace* ws) {
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += p_align(bytes);
    return r;
  };
  const size_t N = g.n_img, cap = g.sel_cap, NL = N * kLevels;
  pw->cand_count = (int*)take(NL * 4);
  pw->cand = (unsigned long long*)take(N * (size_t)g.A * g.C * 8);
  pw->lvl_box = (float4*)take(NL * nms_pre * 16);
  pw->lvl_score = (float*)take(NL * nms_pre * 4);
  pw->lvl_label = (int*)take(NL * nms_pre * 4);
  pw->lvl_count = (int*)take(NL * 4);
  pw->out_box = (float4*)take(N * cap * 16);
  pw->out_label = (int*)take(N * cap * 4);
  pw->box_inds = (int32_t*)take(N * cap * 4);
  pw->box_count = (int32_t*)take(N * 4);
  pw->keep = (int32_t*)take(N * cap * 4);
  pw->keep_count = (int32_t*)take(N * 4);
  pw->flags = (uint8_t*)take(N * cap);
  memset(ws, 0, sizeof(*ws));
  ws->nms_box = (float4*)take(N * cap * 16);
  ws->nms_score = (float*)take(N * cap * 4);
  ws->nms_mask = (unsigned long long*)take(N * cap * nms_words((int)cap) * 8);
  ws->nms_nz = (unsigned long long*)take(N * cap * nms_nz_words((int)cap) * 8);
  ws->keep_raw = (int*)take(N * cap * 4);
  return off;
}

}  // namespace erd

using namespace erd;

// make_geo of api.cu, for the fields this path reads
static int predict_geo(const ErdShape* s, const ErdPredictConfig* cfg, Geo* g) {
  if (!s || !cfg) return (int)ERD_ERR_NULL;
  if (s->num_levels != kLevels || s->reg_max != kBins - 1 || s->num_imgs < 1 || s->num_classes < 1) return (int)ERD_ERR_BAD_SHAPE;
  if (cfg->nms_pre < 1 || cfg->nms_pre > kSortCap || cfg->max_per_img < 1 || kLevels * cfg->nms_pre > 16384)
    return (int)ERD_ERR_BAD_SHAPE;
  memset(g, 0, sizeof(*g));
  g->n_img = s->num_imgs;
  g->C = s->num_classes;
  int a = 0;
  for (int l = 0; l < kLevels; ++l) {
    if (s->level_h[l] < 1 || s->level_w[l] < 1 || s->stride[l] < 1) return (int)ERD_ERR_BAD_SHAPE;
    g->h[l] = s->level_h[l];
    g->w[l] = s->level_w[l];
    g->hw[l] = s->level_h[l] * s->level_w[l];
    g->stride[l] = s->stride[l];
    g->start[l] = a;
    a += g->hw[l];
  }
  if ((long long)a * s->num_classes >= (1ll << 32)) return (int)ERD_ERR_BAD_SHAPE;   // flat (anchor, class) index is 32 bit
  g->sel_cap = kLevels * cfg->nms_pre;
  g->A = g->sel_cap;   // the NMS kernels address their survivor flags as n * A + box_inds[..]: list positions here
  return (int)ERD_OK;
}

extern "C" {

int erd_predict_workspace_bytes(const ErdShape* shape, const ErdPredictConfig* cfg, size_t* bytes) {
  Geo g;
  int rc = predict_geo(shape, cfg, &g);
  if (rc) return rc;
  if (!bytes) return (int)ERD_ERR_NULL;
  Geo ga = g;
  ga.A = 0;
  for (int l = 0; l < kLevels; ++l) ga.A += g.hw[l];
  PredictWs pw;
  Workspace ws;
  *bytes = carve_predict(ga, cfg->nms_pre, nullptr, &pw, &ws);
  return (int)ERD_OK;
}

int erd_predict(const ErdShape* shape, const ErdPredictConfig* cfg, const float* const* cls_scores,
                const float* const* bbox_preds, const int32_t* img_hw, const float* inv_scale, float* dets,
                int32_t* labels, int32_t* num_dets, void* workspace, void* stream) {
  Geo gn;
  int rc = predict_geo(shape, cfg, &gn);
  if (rc) return rc;
  if (!cls_scores || !bbox_preds || !img_hw || !dets || !labels || !num_dets || !workspace) return (int)ERD_ERR_NULL;
  for (int l = 0; l < kLevels; ++l)
    if (!cls_scores[l] || !bbox_preds[l]) return (int)ERD_ERR_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  Geo ga = gn;   // anchor geometry (collect / top-k); gn: the NMS kernels' view (A = list capacity)
  ga.A = 0;
  for (int l = 0; l < kLevels; ++l) ga.A += gn.hw[l];
  PredictWs pw;
  Workspace ws;
  carve_predict(ga, cfg->nms_pre, workspace, &pw, &ws);
  PredictArgs A;
  for (int l = 0; l < kLevels; ++l) {
    A.cls.p[l] = cls_scores[l];
    A.box.p[l] = bbox_preds[l];
  }
  A.img_hw = img_hw;
  A.inv_scale = inv_scale;
  A.nms_pre = cfg->nms_pre;
  A.max_per_img = cfg->max_per_img;
  A.score_thr = cfg->score_thr;
  A.min_size = cfg->min_bbox_size;
  int tiles = 0;
  for (int l = 0; l < kLevels; ++l) {
    A.lvl_tile_start[l] = tiles;
    tiles += (ga.hw[l] + kCollectThreads - 1) / kCollectThreads;
  }
  A.lvl_tile_start[kLevels] = tiles;
  cudaError_t e = cudaMemsetAsync(pw.cand_count, 0, (size_t)ga.n_img * kLevels * sizeof(int), st);
  if (e != cudaSuccess) return (int)ERD_ERR_CUDA;
  predict_collect_kernel<<<dim3(tiles, ga.n_img), kCollectThreads, 0, st>>>(ga, pw, A);
  predict_topk_kernel<<<dim3(kLevels, ga.n_img), kTopkThreads, 0, st>>>(ga, pw, A);
  predict_merge_kernel<<<gn.n_img, kMergeThreads, 0, st>>>(gn, ws, pw, A);
  e = cudaGetLastError();
  if (e == cudaSuccess)
    e = launch_nms_prepared(gn, ws, pw.box_inds, pw.box_count, cfg->iou_threshold, pw.keep, pw.keep_count, pw.flags, st,
                            nullptr);
  if (e == cudaSuccess) {
    predict_output_kernel<<<gn.n_img, 128, 0, st>>>(gn, ws, pw, A, dets, labels, num_dets);
    e = cudaGetLastError();
  }
  return e == cudaSuccess ? (int)ERD_OK : (int)ERD_ERR_CUDA;
}

}  // extern "C"
