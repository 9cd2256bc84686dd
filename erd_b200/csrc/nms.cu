// Teacher box decode + class-aware greedy IoU-NMS over the ERS-selected rows of each image.
// Reference call site: GFLHeadIncrementERD.distill_loss_by_image_single
// (dense_heads/gfl_head_increment_erd.py:189-202).  mmcv.ops.batched_nms is third-party
// (mmcv>=2.0.0rc4,<2.1.0, not in the reference tree); its published semantics are followed:
// boxes_for_nms = boxes + class_id * (boxes.max() + 1); visit by descending score; area
// (x2-x1)*(y2-y1); suppress when inter / (area_i + area_j - inter) > thr; kept indices are
// returned in visiting order.  All box arithmetic is IEEE fp32 without contraction.
// Pairs with an empty intersection are decided without the division (ovr = 0 <= thr).
//
// Launches: prep (one CTA per image: decode, class offsets), predecessor bit matrix (many CTAs
// per image), resolve (one CTA per image; also marks the survivors for the box sweep), and --
// off the critical path -- the ordering of the survivors by score that batched_nms returns.
// Nothing before the resolve pass needs the boxes sorted: "i is visited before j" is evaluated
// per pair as (score_i > score_j) or (equal scores and i earlier in the list).
#include <cstdlib>

#include "erd_common.cuh"

namespace erd {

constexpr int kPrepThreads = 1024;
constexpr int kClsBuckets = 1024;   // classes >= 1023 share the last bucket (grouping is then coarser, never wrong)

// One CTA per image: teacher boxes of the selected rows, the coordinate maximum, class offsets; the rows are then
// GROUPED BY CLASS (counting sort; the order inside a class is arbitrary -- "visited first" is decided per pair from
// scores and list positions): boxes of different classes never overlap after the class offset, so the pair kernel
// only has to look at the tiles whose row and column blocks share a class.
// Zeroes the image's slice of the predecessor matrix and map.
__global__ void __launch_bounds__(kPrepThreads) nms_prep_kernel(Geo g, Workspace ws,
                                                                const int32_t* __restrict__ box_inds,
                                                                const int32_t* __restrict__ box_count,
                                                                const int32_t* __restrict__ pad_hw) {
  __shared__ float s_max[kPrepThreads / 32];
  __shared__ int s_cur[kClsBuckets];
  __shared__ int s_wsum[kPrepThreads / 32];
  const int n = blockIdx.x;
  const int K = box_count[n];
  const int pad_h = pad_hw[n * 2], pad_w = pad_hw[n * 2 + 1];
  const int32_t* list = box_inds + (size_t)n * g.sel_cap;
  float4* tbox = ws.nms_tbox + (size_t)n * g.sel_cap;
  float* tscore = ws.nms_tscore + (size_t)n * g.sel_cap;
  int* tcls = ws.nms_tcls + (size_t)n * g.sel_cap;
  for (int i = threadIdx.x; i < kClsBuckets; i += kPrepThreads) s_cur[i] = 0;
  __syncthreads();
  float mx = -INFINITY;
  for (int r = threadIdx.x; r < K; r += kPrepThreads) {
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
    const int c = ws.t_arg[ga];
    tbox[r] = b;
    tcls[r] = c;
    tscore[r] = ws.t_m[ga];
    atomicAdd(&s_cur[min(c, kClsBuckets - 1)], 1);
    mx = fmaxf(mx, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
  }
  // zero the predecessor words / map of the K rows in use
  const int W = (K + 63) >> 6;
  const int Wcap = nms_words(g.sel_cap), NZ = nms_nz_words(g.sel_cap);
  unsigned long long* pred = ws.nms_mask + (size_t)n * g.sel_cap * Wcap;
  unsigned long long* nz = ws.nms_nz + (size_t)n * g.sel_cap * NZ;
  for (int i = threadIdx.x; i < K * W; i += kPrepThreads) pred[(size_t)(i / W) * Wcap + (i % W)] = 0ull;
  for (int i = threadIdx.x; i < K * NZ; i += kPrepThreads) nz[i] = 0ull;
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
  __syncthreads();
  // exclusive prefix of the class histogram (thread t owns bucket t): the first grouped position of every class
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int v = s_cur[threadIdx.x];
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += s_wsum[w];
    s_cur[threadIdx.x] = base + inc - v;
  }
  __syncthreads();
  float maxc = s_max[0];
  for (int w = 1; w < kPrepThreads / 32; ++w) maxc = fmaxf(maxc, s_max[w]);
  const float unit = __fadd_rn(maxc, 1.0f);   // boxes.max() + 1
  float4* boxes = ws.nms_box + (size_t)n * g.sel_cap;
  float* score = ws.nms_score + (size_t)n * g.sel_cap;
  int* cls = ws.nms_cls + (size_t)n * g.sel_cap;
  int* orig = ws.nms_orig + (size_t)n * g.sel_cap;
  for (int r = threadIdx.x; r < K; r += kPrepThreads) {   // each thread re-reads what it wrote
    const float4 b = tbox[r];
    const int c = tcls[r];
    const float off = __fmul_rn((float)c, unit);
    const int bucket = min(c, kClsBuckets - 1);
    const int p = atomicAdd(&s_cur[bucket], 1);
    boxes[p] = make_float4(__fadd_rn(b.x, off), __fadd_rn(b.y, off), __fadd_rn(b.z, off), __fadd_rn(b.w, off));
    score[p] = tscore[r];
    cls[p] = bucket;
    orig[p] = r;
  }
}

// Predecessor bit matrix in list order: bit i of pred[j][i / 64] is set when box i is visited
// before box j by the greedy pass and overlaps it above the threshold.  Overlaps are rare, so
// bits are set with atomics and the per-box map nz[j] records which words are non-zero.
// 256 threads per 64x64 tile (rb <= cb): four threads share a column, 16 rows each.
constexpr int kMaskThreads = 256;
constexpr int kMaxMaskWords = 264;   // 64-box blocks of an image (sel_cap <= 16 384 + slack); beyond it: no tile skipping

__device__ __forceinline__ bool nms_overlaps(const float4& a, float area_a, const float4& b, float area_b,
                                             float iou_thr) {
  const float w = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
  const float h = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
  if (!(w > 0.f) || !(h > 0.f)) return 0.f > iou_thr;   // inter == 0: ovr is 0 (or NaN for empty boxes)
  const float inter = __fmul_rn(w, h);
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
  return ovr > iou_thr;
}

__device__ __forceinline__ void nms_tile_of(int t, int W, int& rb, int& cb) {
  rb = 0;   // t -> (rb, cb), cb >= rb, row-major over the upper triangle
  while (t >= W - rb) { t -= W - rb; ++rb; }
  cb = rb + t;
}

__global__ void __launch_bounds__(kMaskThreads) nms_mask_kernel(Geo g, Workspace ws,
                                                                const int32_t* __restrict__ box_count,
                                                                float iou_thr) {
  const int n = blockIdx.y;
  const int K = box_count[n];
  const int W = (K + 63) >> 6;
  const int Wcap = nms_words(g.sel_cap);
  const int NZ = nms_nz_words(g.sel_cap);
  const float4* boxes = ws.nms_box + (size_t)n * g.sel_cap;
  const float* score = ws.nms_score + (size_t)n * g.sel_cap;
  unsigned long long* pred = ws.nms_mask + (size_t)n * g.sel_cap * Wcap;
  unsigned long long* nz = ws.nms_nz + (size_t)n * g.sel_cap * NZ;
  __shared__ float4 s_row[2][64];
  __shared__ float2 s_rx[2][64];   // (x1, x2) of the row boxes: most pairs are rejected on x alone
  __shared__ float s_rarea[2][64];
  __shared__ float s_rscore[2][64];
  __shared__ int s_rorig[2][64];
  __shared__ int s_cfirst[kMaxMaskWords], s_clast[kMaxMaskWords];   // first / last class of every 64-box block
  const int col = threadIdx.x >> 2, part = threadIdx.x & 3;
  const int ntile = W * (W + 1) / 2;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // Grouped by class (loss path): list positions for the "visited first" rule, and the class range of every block --
  // a tile whose row block ends below the class its column block starts with holds no pair of one class.
  const int* orig = ws.nms_orig ? ws.nms_orig + (size_t)n * g.sel_cap : nullptr;
  const bool grouped = orig != nullptr && W <= kMaxMaskWords;
  if (grouped) {
    const int* cls = ws.nms_cls + (size_t)n * g.sel_cap;
    for (int b = threadIdx.x; b < W; b += kMaskThreads) {
      s_cfirst[b] = cls[b * 64];
      s_clast[b] = cls[min(K - 1, b * 64 + 63)];
    }
    __syncthreads();
  }
  // tiles of this CTA: the needed ones among blockIdx.x, blockIdx.x + gridDim.x, ...; the boxes of the next tile
  // are fetched while the current one is evaluated
  int rb = 0, cb = 0;
  auto next_needed = [&](int t) {   // first needed tile at or after t in this CTA's sequence (ntile: none)
    for (; t < ntile; t += gridDim.x) {
      nms_tile_of(t, W, rb, cb);
      if (!grouped || s_clast[rb] >= s_cfirst[cb]) break;
    }
    return t;
  };
  int t = next_needed(blockIdx.x);
  if (t >= ntile) return;
  const bool rthread = threadIdx.x < 64;
  auto oidx = [&](int p) { return orig ? orig[p] : p; };
  float4 next_row = (rthread && rb * 64 + threadIdx.x < K) ? boxes[rb * 64 + threadIdx.x] : zero4;
  float next_rs = (rthread && rb * 64 + threadIdx.x < K) ? score[rb * 64 + threadIdx.x] : 0.f;
  int next_ro = (rthread && rb * 64 + threadIdx.x < K) ? oidx(rb * 64 + threadIdx.x) : 0;
  float4 next_col = cb * 64 + col < K ? boxes[cb * 64 + col] : zero4;
  float next_cs = cb * 64 + col < K ? score[cb * 64 + col] : 0.f;
  int next_co = cb * 64 + col < K ? oidx(cb * 64 + col) : 0;
  int buf = 0;
  while (t < ntile) {
    if (rthread) {
      s_row[buf][threadIdx.x] = next_row;
      s_rx[buf][threadIdx.x] = make_float2(next_row.x, next_row.z);
      s_rarea[buf][threadIdx.x] = __fmul_rn(__fsub_rn(next_row.z, next_row.x), __fsub_rn(next_row.w, next_row.y));
      s_rscore[buf][threadIdx.x] = next_rs;
      s_rorig[buf][threadIdx.x] = next_ro;
    }
    const float4 a = next_col;
    const float sa = next_cs;
    const int oa = next_co;
    const int crb = rb, ccb = cb;
    const int tn = next_needed(t + gridDim.x);
    if (tn < ntile) {
      next_row = (rthread && rb * 64 + threadIdx.x < K) ? boxes[rb * 64 + threadIdx.x] : zero4;
      next_rs = (rthread && rb * 64 + threadIdx.x < K) ? score[rb * 64 + threadIdx.x] : 0.f;
      next_ro = (rthread && rb * 64 + threadIdx.x < K) ? oidx(rb * 64 + threadIdx.x) : 0;
      next_col = cb * 64 + col < K ? boxes[cb * 64 + col] : zero4;
      next_cs = cb * 64 + col < K ? score[cb * 64 + col] : 0.f;
      next_co = cb * 64 + col < K ? oidx(cb * 64 + col) : 0;
    }
    __syncthreads();
    const int j = ccb * 64 + col;
    if (j < K) {
      const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
      const int rmax = (crb == ccb) ? col : min(64, K - crb * 64);   // pairs i < j, once each
      const bool neg_thr = 0.f > iou_thr;   // then even disjoint boxes "overlap": no shortcut
#pragma unroll 4
      for (int r = part * 16; r < part * 16 + 16; ++r) {
        if (r >= rmax) break;
        const float2 bx = s_rx[buf][r];
        if (!neg_thr && !(fminf(a.z, bx.y) > fmaxf(a.x, bx.x))) continue;   // w <= 0: ovr is 0 (or NaN)
        const float4 b = s_row[buf][r];
        if (!nms_overlaps(b, s_rarea[buf][r], a, area_a, iou_thr)) continue;
        const int i = crb * 64 + r;
        const float si = s_rscore[buf][r];
        if (si > sa || (si == sa && s_rorig[buf][r] < oa)) {   // i is visited first: higher score, on equal scores the earlier list position
          atomicOr(pred + (size_t)j * Wcap + crb, 1ull << r);
          atomicOr(nz + (size_t)j * NZ + (crb >> 6), 1ull << (crb & 63));
        } else {
          atomicOr(pred + (size_t)i * Wcap + ccb, 1ull << col);
          atomicOr(nz + (size_t)i * NZ + (ccb >> 6), 1ull << (ccb & 63));
        }
      }
    }
    buf ^= 1;
    t = tn;
  }
}

// Greedy NMS as rounds over the predecessor matrix, one CTA per image: a box is suppressed
// as soon as a kept predecessor overlaps it and kept once every overlapping predecessor is
// decided and none is kept -- exactly the sequential greedy result, in as many rounds as the
// longest overlap chain.  Each round reads the previous round's state (double buffered).
// Epilogue: survivors in score order; each survivor is flagged (bit 2 of its anchor's flag
// byte) and its list position recorded so the box sweep finds its distillation row.
constexpr int kResThreads = 512;

__global__ void __launch_bounds__(kResThreads) nms_resolve_kernel(Geo g, Workspace ws,
                                                                  const int32_t* __restrict__ box_inds,
                                                                  const int32_t* __restrict__ box_count,
                                                                  int32_t* __restrict__ keep,
                                                                  int32_t* __restrict__ keep_count,
                                                                  uint8_t* __restrict__ sel_flags) {
  extern __shared__ unsigned long long s_state[];   // kept[W] decided[W] kept_next[W] decided_next[W]
  const int n = blockIdx.x;
  const int K = box_count[n];
  const int W = (K + 63) >> 6;
  const int Wcap = nms_words(g.sel_cap);
  const int NZ = nms_nz_words(g.sel_cap);
  unsigned long long* kept = s_state;
  unsigned long long* dec = s_state + Wcap;
  unsigned long long* kept_n = s_state + 2 * Wcap;
  unsigned long long* dec_n = s_state + 3 * Wcap;
  const unsigned long long* pred = ws.nms_mask + (size_t)n * g.sel_cap * Wcap;
  const unsigned long long* nz = ws.nms_nz + (size_t)n * g.sel_cap * NZ;
  for (int w = threadIdx.x; w < 4 * Wcap; w += kResThreads) s_state[w] = 0ull;
  __syncthreads();
  // Each thread owns boxes j = tid, tid + 512, ...; the (few) non-zero predecessor words of its
  // first kOwn boxes are fetched once into registers, so a round costs shared-memory traffic
  // only.  Boxes beyond kOwn per thread or with more than kCache non-zero words re-read global.
  constexpr int kOwn = 3, kCache = 4;
  unsigned long long cbits[kOwn][kCache];
  int cword[kOwn][kCache];
  int cnum[kOwn];
#pragma unroll
  for (int o = 0; o < kOwn; ++o) {
    cnum[o] = 0;
    const int j = threadIdx.x + o * kResThreads;
    if (j >= K) continue;
    int cnt = 0;
    for (int zw = 0; zw < NZ; ++zw) {
      unsigned long long m = nz[(size_t)j * NZ + zw];
      while (m) {
        const int w = (zw << 6) + __ffsll((long long)m) - 1;
        m &= m - 1ull;
        if (cnt < kCache) {
          cword[o][cnt] = w;
          cbits[o][cnt] = pred[(size_t)j * Wcap + w];
        }
        ++cnt;
      }
    }
    cnum[o] = cnt;   // > kCache: not fully cached
  }
  int pending = K > 0;
  while (pending) {
    int undecided = 0;
    bool now_dec[kOwn], now_kept[kOwn];
#pragma unroll
    for (int o = 0; o < kOwn; ++o) {
      now_dec[o] = now_kept[o] = false;
      const int j = threadIdx.x + o * kResThreads;
      if (j >= K || cnum[o] > kCache) continue;
      const int wj = j >> 6;
      const unsigned long long bit = 1ull << (j & 63);
      if (dec[wj] & bit) continue;
      bool sup = false, wait = false;
#pragma unroll
      for (int c = 0; c < kCache; ++c) {
        if (c >= cnum[o]) break;
        const unsigned long long pr = cbits[o][c];
        if (pr & kept[cword[o][c]]) sup = true;
        if (pr & ~dec[cword[o][c]]) wait = true;
      }
      now_dec[o] = sup || !wait;
      now_kept[o] = !sup && !wait;
      if (wait && !sup) undecided = 1;
    }
    // a warp owns the 32 consecutive boxes of each of its slots = one 32-bit half of a state
    // word: publish the round's decisions with a ballot and a plain store, no atomics
#pragma unroll
    for (int o = 0; o < kOwn; ++o) {
      const unsigned kd_bits = __ballot_sync(0xffffffffu, now_dec[o]);
      const unsigned kk_bits = __ballot_sync(0xffffffffu, now_kept[o]);
      const int half = (threadIdx.x + o * kResThreads) >> 5;   // warp-uniform
      if ((threadIdx.x & 31) == 0 && kd_bits) {
        reinterpret_cast<unsigned*>(dec_n)[half] |= kd_bits;
        reinterpret_cast<unsigned*>(kept_n)[half] |= kk_bits;
      }
    }
    for (int j = threadIdx.x; j < K; j += kResThreads) {   // uncached boxes
      const int o = j / kResThreads;
      if (o < kOwn && cnum[o < kOwn ? o : 0] <= kCache) continue;
      const int wj = j >> 6;
      const unsigned long long bit = 1ull << (j & 63);
      if (dec[wj] & bit) continue;
      bool sup = false, wait = false;
      for (int zw = 0; zw < NZ && !sup; ++zw) {
        unsigned long long m = nz[(size_t)j * NZ + zw];
        while (m) {
          const int w = (zw << 6) + __ffsll((long long)m) - 1;
          m &= m - 1ull;
          const unsigned long long pr = pred[(size_t)j * Wcap + w];
          if (pr & kept[w]) { sup = true; break; }
          if (pr & ~dec[w]) wait = true;
        }
      }
      const unsigned b32 = 1u << (j & 31);
      if (sup) {
        atomicOr(reinterpret_cast<unsigned*>(dec_n) + (j >> 5), b32);
      } else if (!wait) {
        atomicOr(reinterpret_cast<unsigned*>(kept_n) + (j >> 5), b32);
        atomicOr(reinterpret_cast<unsigned*>(dec_n) + (j >> 5), b32);
      } else {
        undecided = 1;
      }
    }
    pending = __syncthreads_or(undecided);
    for (int w = threadIdx.x; w < W; w += kResThreads) { kept[w] = kept_n[w]; dec[w] = dec_n[w]; }
    __syncthreads();
  }
  // survivors: count, unordered compact list (ordered by score later, off the critical path),
  // and the marks the box sweep needs: flag bit 2 and the row of the distillation gradient
  if (threadIdx.x == 0) {
    int run = 0;
    for (int w = 0; w < W; ++w) {
      const int c = __popcll(kept[w]);
      dec[w] = (unsigned long long)run;
      run += c;
    }
    keep_count[n] = run;
  }
  __syncthreads();
  int* out = ws.keep_raw + (size_t)n * g.sel_cap;   // score order comes later (nms_order_kernel)
  const int32_t* list = box_inds + (size_t)n * g.sel_cap;
  const int* orig = ws.nms_orig ? ws.nms_orig + (size_t)n * g.sel_cap : nullptr;
  for (int j = threadIdx.x; j < K; j += kResThreads) {
    const int wj = j >> 6;
    const unsigned long long kw = kept[wj];
    if (!(kw & (1ull << (j & 63)))) continue;
    out[(int)dec[wj] + __popcll(kw & ((1ull << (j & 63)) - 1ull))] = j;   // position in nms_box / nms_score order
    sel_flags[(size_t)n * g.A + list[orig ? orig[j] : j]] |= 4;
  }
}

// keep list in descending-score order (ties: earlier list position first), what batched_nms
// returns.  One CTA per image, bitonic sort of the survivors in shared memory.
constexpr int kSortThreads = 1024;

__global__ void __launch_bounds__(kSortThreads) nms_order_kernel(Geo g, Workspace ws, int32_t* __restrict__ keep,
                                                                 const int32_t* __restrict__ keep_count) {
  extern __shared__ unsigned long long s_key[];
  const int n = blockIdx.x;
  const int M = keep_count[n];
  int32_t* out = keep + (size_t)n * g.sel_cap;
  const int* raw = ws.keep_raw + (size_t)n * g.sel_cap;
  const int* orig = ws.nms_orig ? ws.nms_orig + (size_t)n * g.sel_cap : nullptr;   // grouped position -> list position
  if (M < 2) {
    if (M == 1 && threadIdx.x == 0) out[0] = orig ? orig[raw[0]] : raw[0];
    return;
  }
  int P = 1;
  while (P < M) P <<= 1;
  const float* score = ws.nms_score + (size_t)n * g.sel_cap;
  for (int r = threadIdx.x; r < P; r += kSortThreads) {
    unsigned long long key = ~0ull;
    if (r < M) {
      const int pos = raw[r];
      key = ((unsigned long long)(~__float_as_uint(score[pos])) << 32) | (unsigned int)(orig ? orig[pos] : pos);
    }
    s_key[r] = key;
  }
  __syncthreads();
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
  for (int r = threadIdx.x; r < M; r += kSortThreads) out[r] = (int)(unsigned int)(s_key[r] & 0xffffffffull);
}

// mask -> resolve -> order over boxes / scores already in ws.nms_box / ws.nms_score with clean predecessor rows
// (the loss path's nms_prep or the inference path's merge kernel put them there)
cudaError_t launch_nms_prepared(const Geo& g, const Workspace& ws, const int32_t* box_inds, const int32_t* box_count,
                                float iou_thr, int32_t* keep, int32_t* keep_count, uint8_t* sel_flags, cudaStream_t st,
                                cudaEvent_t resolved) {
  // The chain runs beside the student pass on the few SMs that pass leaves free, so the grid is sized for
  // those: each CTA loops over the image's tiles (K = 500 candidates are 36 tiles), and a launch of
  // thousands of CTAs that mostly exit at once would queue behind each other there.
  static const int mask_ctas = [] {
    const char* e = getenv("ERD_NMS_MASK_CTAS");
    const int v = e ? atoi(e) : 32;
    return v < 1 ? 1 : v;
  }();
  ERD_LAUNCH(kKNmsMask, st,
             (nms_mask_kernel<<<dim3(mask_ctas, g.n_img), kMaskThreads, 0, st>>>(g, ws, box_count, iou_thr)));
  const size_t res_smem = sizeof(unsigned long long) * 4 * (size_t)nms_words(g.sel_cap);
  ERD_LAUNCH(kKNmsScan, st,
             (nms_resolve_kernel<<<g.n_img, kResThreads, res_smem, st>>>(g, ws, box_inds, box_count, keep,
                                                                         keep_count, sel_flags)));
  // everything the loss needs exists now; the ordering of the keep list is only an output
  if (resolved) {
    cudaError_t e = cudaEventRecord(resolved, st);
    if (e != cudaSuccess) return e;
  }
  int P = 1;
  while (P < g.sel_cap) P <<= 1;
  const size_t sort_smem = sizeof(unsigned long long) * (size_t)P;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(nms_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done = true;
  }
  if (sort_smem > 200 * 1024) return cudaErrorInvalidValue;
  ERD_LAUNCH(kKNmsOrder, st, (nms_order_kernel<<<g.n_img, kSortThreads, sort_smem, st>>>(g, ws, keep, keep_count)));
  return cudaGetLastError();
}

cudaError_t launch_nms(const Geo& g, const Workspace& ws, const int32_t* box_inds, const int32_t* box_count,
                       const int32_t* pad_hw, float iou_thr, int32_t* keep, int32_t* keep_count, uint8_t* sel_flags,
                       cudaStream_t st, cudaEvent_t prepped, cudaEvent_t resolved) {
  ERD_LAUNCH(kKNmsSort, st, (nms_prep_kernel<<<g.n_img, kPrepThreads, 0, st>>>(g, ws, box_inds, box_count, pad_hw)));
  if (prepped) {   // lets the caller start DRAM-heavy work only after the latency-bound gathers
    cudaError_t e = cudaEventRecord(prepped, st);
    if (e != cudaSuccess) return e;
  }
  return launch_nms_prepared(g, ws, box_inds, box_count, iou_thr, keep, keep_count, sel_flags, st, resolved);
}

}  // namespace erd
