// Shared device/host definitions for the erd_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/erd_b200.h"

namespace erd {

constexpr int kLevels = 5;
constexpr int kBins = 17;             // reg_max + 1
constexpr int kBoxCh = 4 * kBins;     // 68 box-distribution channels
constexpr int kTile = 512;            // anchors per CTA in the streaming kernels
constexpr int kTileThreads = 128;     // 4 anchors per thread
constexpr int kTopK = 9;              // ATSSAssigner topk (config train_cfg.assigner)
constexpr float kEps32 = 1.1920928955078125e-07f;  // torch.finfo(float32).eps

// Per-anchor record of an assigned (positive) anchor, written by the positives' prepass and
// fetched by the student pass with two 16-byte asynchronous copies.
struct __align__(32) PosRec {
  float4 gt;      // assigned GT box, px
  int label;      // new-class label, -1 when the GT's label lies outside the new-class range
  float score;    // IoU quality score of the decoded box (detached)   gfl_head_increment_erd.py:289-292
  float w;        // weight_targets: max_c sigmoid(new-class logits)  :283-284
  int pslot;      // index in the image's positives list / row of ws.pos_rows
};

struct Ptr5 {
  const float* p[kLevels];
};
struct MPtr5 {
  float* p[kLevels];
};

// Static geometry of a batch, passed by value to every kernel.
struct Geo {
  int n_img, C, ori, cn, A;
  int h[kLevels], w[kLevels], hw[kLevels], stride[kLevels], start[kLevels];
  int tile_start[kLevels + 1];  // prefix sum of ceil(hw/kTile): CTA index -> level
  int vec[kLevels];             // 1 when hw % 4 == 0 and all level pointers are 16 B aligned
  int vec2[kLevels];            // 1 when hw % 2 == 0 and all level pointers are 8 B aligned
  float half[kLevels];          // anchor half size = 0.5 * stride * scale
  int sel_cap;                  // A/5 + 1
  int pos_cap;                  // min(A, 45 * max_gt_per_img): rows of the positives' gradient buffer
  int total_gt;
  float w_cls, w_bbox, w_dfl, w_ld, T;   // loss weights and KD temperature
};

// Device workspace carved from the caller's buffer (see workspace.cu).
struct Workspace {
  float* t_m;                     // [N][A] teacher max_c sigmoid(cls)
  int* t_arg;                     // [N][A] teacher argmax class
  float* t_u;                     // [N][A] teacher max raw box logit
  float4* t_dist;                 // [N][A] teacher softmax-integral distances (l,t,r,b), bin units
  double* ers_part;               // [N][tiles of 32 anchors][4] per-tile sums: m, m^2, u, u^2
  // The stash: the teacher's logit column [ori + 68, padded to 4] of every anchor that clears the
  // PROVISIONAL thresholds (the lowest mean + 1.7 std any image of the PREVIOUS call had), written by
  // the teacher pass while the tile is in shared memory, so that the student pass does not gather those
  // columns from the NCHW tensors again (64 B of DRAM per 4 B used).  Region [n][cta][kStashPerCta].
  float* t_stash;
  unsigned short* t_slot;         // [N][A] stash row of the anchor + 1, 0: not stashed (read the tensors)
  unsigned int* pthr_state;       // [4] provisional thresholds as ~ordered_bits(float), 0 = none yet: [0..1] in use by the
                                  // teacher pass (class response, box), [2..3] being collected by the flags kernel
  unsigned long long* atss_key;   // [N][A] packed (iou bits << 32 | ~gt) argmax table
  int2* pos_list;                 // [N][A] (anchor, global GT row) of the assigned anchors (unordered)
  int* pos_counter;               // [N] running length of pos_list (zero between steps)
  struct PosRec* pos_rec;         // [N][A] what the student pass needs of a positive, defined at positives only
  double* pre_pub;                // [2L+1] pre_acc of the last avg-factor pass, read by finalize
  double* pre_acc;                // [2L+1] sum w(1-giou) per level, sum w*dfl per level, sum w
  int* keep_raw;                  // [N][sel_cap] NMS survivors as list positions, unordered (resolve pass)
  float* kd_loss;                 // [N][A] weighted KL of every ERS box candidate, at its anchor
  float* pos_rows;                // [N][pos_cap][68] box-logit gradient rows of positives that are also ERS box candidates
  unsigned long long* nms_nz;     // [N][sel_cap][nz_words] which words of a predecessor row are non-zero
  unsigned int* counters;         // [8] last-block tickets
  float* nms_score;               // [N][sel_cap] teacher confidence of each selected row, list order
  int* nms_cls;                   // [N][sel_cap] class ids in list order
  float4* nms_box;                // [N][sel_cap] class-offset teacher boxes, list order
  // Loss path only (NULL on the inference path): nms_box / nms_score / nms_cls are held GROUPED BY CLASS, so that the
  // pair kernel can skip every 64 x 64 tile whose row and column blocks share no class; nms_orig[p] is the list
  // position of the box at grouped position p, nms_tbox / nms_tscore / nms_tcls the list-order values before grouping.
  int* nms_orig;
  float4* nms_tbox;
  float* nms_tscore;
  int* nms_tcls;
  unsigned long long* nms_mask;   // [N][sel_cap][W] predecessor bit matrix, W = ceil(sel_cap/64)
  double* loss_acc;               // [3L + 2N]
  size_t bytes;
};

constexpr int kStashCtas = 160;        // the teacher pass runs on at most this many CTAs (one per SM)
constexpr int kStashPerCta = 64;      // stash rows per (image, CTA); an anchor that does not fit is simply not stashed
constexpr int kStashRows = kStashCtas * kStashPerCta;   // per image; < 65535 (t_slot is 16 bit)
inline __host__ __device__ int stash_pitch(int ori) { return (ori + kBoxCh + 3) & ~3; }   // floats per stash row (16 B multiple)

inline __host__ __device__ int nms_words(int sel_cap) { return (sel_cap + 63) / 64; }
inline __host__ __device__ int nms_nz_words(int sel_cap) { return (nms_words(sel_cap) + 63) / 64; }

// ---------------------------------------------------------------- the avg-factor exchange (exchange.cu)
// Every rank owns a symmetric buffer all peers have mapped: 2 x kMaxRanks slots (double buffered by the parity of
// the epoch), then the epoch counter and a status word.
constexpr int kMaxRanks = 64;
struct ExchangeSlot {
  float a0, a1;
  unsigned int epoch, pad;
};
struct ExchangePeers {
  unsigned char* buf[kMaxRanks];
};
constexpr size_t kExchangeSlotBytes = sizeof(ExchangeSlot) * 2 * kMaxRanks;   // then: epoch counter, status
struct ExchangeInfo {   // passed by value to the kernels that post / wait; world <= 1: no exchange
  ExchangePeers peers;
  int rank, world;
};

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ int level_of_tile(const Geo& g, int tile) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < kLevels; ++i) l += (tile >= g.tile_start[i]) ? 1 : 0;
  return l;
}

__device__ __forceinline__ int level_of_anchor(const Geo& g, int a) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < kLevels; ++i) l += (a >= g.start[i]) ? 1 : 0;
  return l;
}

// torch.sigmoid on fp32 is 1/(1+exp(-x)) evaluated in fp32 (ATen UnaryOpsKernel); the
// index-critical statistics (ERS, weights) use the same formula with IEEE ops.
__device__ __forceinline__ float sigmoid_ref(float x) {
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// 2^x, 2 ulp, denormal results flushed to zero: one MUFU, without the range fix-up (a compare
// and two multiplies) the non-ftz __expf / exp2f carry.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// float <-> unsigned int whose unsigned order is the floats' order
__device__ __forceinline__ unsigned int ordered_bits(float x) {
  const unsigned int u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Four anchors per thread.  VEC: four consecutive hw positions (one 16 B access per
// channel); otherwise positions tid, tid+128, tid+256, tid+384 of the tile (coalesced 4 B).
template <bool VEC>
struct Quad {
  int hw[4];
  bool ok[4];
  __device__ __forceinline__ Quad(int hw0, int limit) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      hw[k] = VEC ? hw0 + (int)threadIdx.x * 4 + k : hw0 + (int)threadIdx.x + kTileThreads * k;
      ok[k] = hw[k] < limit;
    }
  }
  // plane points at hw = 0 of one (image, channel) plane
  __device__ __forceinline__ void load(const float* __restrict__ plane, float (&v)[4], float fill) const {
    if (VEC) {
      if (ok[0]) {
        const float4 t = __ldcs(reinterpret_cast<const float4*>(plane + hw[0]));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
        v[0] = v[1] = v[2] = v[3] = fill;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = ok[k] ? __ldcs(plane + hw[k]) : fill;
    }
  }
  __device__ __forceinline__ void store(float* __restrict__ plane, const float (&v)[4]) const {
    if (VEC) {
      if (ok[0]) __stcs(reinterpret_cast<float4*>(plane + hw[0]), make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (ok[k]) __stcs(plane + hw[k], v[k]);
    }
  }
  __device__ __forceinline__ void store_zero(float* __restrict__ plane) const {
    const float z[4] = {0.f, 0.f, 0.f, 0.f};
    store(plane, z);
  }
};

// Post this rank's two local factors to every peer (thread t < world stores into peer t's buffer) under a new
// epoch.  Call with the whole CTA (one __syncthreads inside); `s_epoch` is a shared-memory word.
__device__ __forceinline__ void exchange_post(const ExchangeInfo& x, float a0, float a1, unsigned int* s_epoch) {
  unsigned char* mine = x.peers.buf[x.rank];
  unsigned int* ctr = reinterpret_cast<unsigned int*>(mine + kExchangeSlotBytes);
  if (threadIdx.x == 0) {
    *s_epoch = ctr[0] + 1u;
    ctr[0] = *s_epoch;
  }
  __syncthreads();
  const unsigned int e = *s_epoch;
  const int t = threadIdx.x;
  if (t < x.world) {
    ExchangeSlot* dst = reinterpret_cast<ExchangeSlot*>(x.peers.buf[t]) + (e & 1u) * kMaxRanks + x.rank;
    *reinterpret_cast<volatile float*>(&dst->a0) = a0;
    *reinterpret_cast<volatile float*>(&dst->a1) = a1;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&dst->epoch), "r"(e) : "memory");   // orders the two stores above
  }
}

// Wait until every rank's factors of epoch `e` have arrived in this rank's buffer and average them as reduce_mean
// does (t / W summed in rank order: identical bits on every rank).  Call with the whole CTA; `s_v` is shared memory
// [kMaxRanks][2].  A peer that never arrives (spin bound: minutes) is fatal, not silent: the factors become NaN.
__device__ __forceinline__ void exchange_wait(const ExchangeInfo& x, unsigned int e, float (*s_v)[2], float& avg0, float& avg1) {
  unsigned char* mine = x.peers.buf[x.rank];
  unsigned int* ctr = reinterpret_cast<unsigned int*>(mine + kExchangeSlotBytes);
  const int t = threadIdx.x;
  if (t < x.world) {
    const ExchangeSlot* src = reinterpret_cast<const ExchangeSlot*>(mine) + (e & 1u) * kMaxRanks + t;
    unsigned int seen = 0;
    long long spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(&src->epoch) : "memory");
    } while (seen != e && ++spins < (1ll << 28));
    const bool lost = seen != e;
    if (lost) ctr[1] = 1u;   // status word: a peer never arrived
    s_v[t][0] = lost ? __int_as_float(0x7fc00000) : *reinterpret_cast<const volatile float*>(&src->a0);
    s_v[t][1] = lost ? __int_as_float(0x7fc00000) : *reinterpret_cast<const volatile float*>(&src->a1);
  }
  __syncthreads();
  const float w = (float)x.world;
  avg0 = 0.f;
  avg1 = 0.f;
  for (int r = 0; r < x.world; ++r) {   // t.div_(world) then SUM, rank order (dist_utils.py:59-65)
    avg0 += s_v[r][0] / w;
    avg1 += s_v[r][1] / w;
  }
}

// ---------------------------------------------------------------- launch accounting (profile.cu)
enum KernelId { kKErsScan, kKErsFlags, kKErsSelect, kKAtssCand, kKAtssFin, kKAvg, kKNmsSort, kKNmsMask, kKNmsScan, kKNmsOrder,
                kKUpCheck, kKStudent, kKBoxFix, kKTeacherHead, kNumKernels };
void prof_begin(int id, cudaStream_t st);
void prof_end(int id, cudaStream_t st);
// Developer build only (-DERD_DEV_ABLATE): ERD_ABLATE=<mask> skips kernels by id to measure what
// each one costs inside the full schedule.  Results are garbage when set.
#ifdef ERD_DEV_ABLATE
bool ablated(int id);
#define ERD_ABLATED(id) ::erd::ablated(id)
#else
#define ERD_ABLATED(id) false
#endif
// ERD_LAUNCH(id, stream, kernel<<<...>>>(...)) counts the launch and, when profiling is on,
// brackets it with CUDA events on the launching stream.
#define ERD_LAUNCH(id, st, ...) \
  do {                          \
    if (ERD_ABLATED(id)) break; \
    ::erd::prof_begin(id, st);  \
    __VA_ARGS__;                \
    ::erd::prof_end(id, st);    \
  } while (0)


// ----------------------------------------------------------------------------- ATSS decode
struct LevelView {
  int W, vw, vh, stride, start;
  float half;
};

__device__ __forceinline__ LevelView level_view(const Geo& g, int l, int pad_h, int pad_w) {
  LevelView v;
  v.W = g.w[l];
  v.stride = g.stride[l];
  v.start = g.start[l];
  v.half = g.half[l];
  // valid_flags: x < min(ceil(pad_w / s), W), y < min(ceil(pad_h / s), H)  (anchor_generator.py:434-442)
  v.vw = min((pad_w + v.stride - 1) / v.stride, g.w[l]);
  v.vh = min((pad_h + v.stride - 1) / v.stride, g.h[l]);
  return v;
}

// Anchor a of image n: argmax table entry -> assigned_gt_inds (-1 invalid, 0 background, k > 0
// = GT k-1 of the image; atss_assigner.py:236-246, gfl_head.py:613-640).  Positives are appended
// to the image's list; returns the global GT row (>= 0) of a positive, -1 otherwise.
__device__ __forceinline__ int atss_decode_only(const Geo& g, const Workspace& ws, int pad_h, int pad_w, int first_gt,
                                                int32_t* __restrict__ gt_inds, int n, int a, unsigned long long key) {
  const int l = level_of_anchor(g, a);
  const LevelView v = level_view(g, l, pad_h, pad_w);
  const int r = a - v.start;
  const int x = r % v.W, y = r / v.W;
  int out = -1;
  if (x < v.vw && y < v.vh) out = key ? (int)(0xffffffffu - (unsigned int)(key & 0xffffffffull)) + 1 : 0;
  gt_inds[(size_t)n * g.A + a] = out;
  if (key) ws.atss_key[(size_t)n * g.A + a] = 0ull;   // leave the table clean for the next step
  return out <= 0 ? -1 : first_gt + out - 1;
}

__device__ __forceinline__ int atss_decode_key(const Geo& g, const Workspace& ws, int pad_h, int pad_w, int first_gt,
                                               int32_t* __restrict__ gt_inds, int n, int a, unsigned long long key) {
  const int gidx = atss_decode_only(g, ws, pad_h, pad_w, first_gt, gt_inds, n, a, key);
  if (gidx < 0) return -1;
  const int slot = atomicAdd(ws.pos_counter + n, 1);
  ws.pos_list[(size_t)n * g.A + slot] = make_int2(a, gidx);
  ws.pos_rec[(size_t)n * g.A + a].pslot = slot;
  return gidx;
}

__device__ __forceinline__ int atss_decode_anchor(const Geo& g, const Workspace& ws, const int32_t* __restrict__ pad_hw,
                                                  const int32_t* __restrict__ gt_offsets,
                                                  int32_t* __restrict__ gt_inds, int n, int a) {
  // candidates are valid anchors only, so the key of an invalid one is always 0
  const unsigned long long key = ws.atss_key[(size_t)n * g.A + a];
  return atss_decode_key(g, ws, pad_hw[n * 2], pad_hw[n * 2 + 1], gt_offsets[n], gt_inds, n, a, key);
}


cudaError_t launch_ers(const Geo& g, const Workspace& ws, const Ptr5& t_cls, const Ptr5& t_box,
                       int32_t* cls_inds, int32_t* cls_count, int32_t* box_inds, int32_t* box_count,
                       float* thr, uint8_t* sel_flags, cudaStream_t st);
// the three parts of launch_ers: the streaming pass (cache, thresholds), the flags + counts, the ordered lists
cudaError_t launch_teacher_pass(const Geo& g, const Workspace& ws, const Ptr5& t_cls, const Ptr5& t_box,
                                int32_t* cls_count, int32_t* box_count, int* tiles_per_img,
                                cudaStream_t st);   // also zeroes the two count vectors
// teacher_head.cu: the teacher head's last convolutions fused with the teacher pass (tcgen05 implicit GEMM)
int head_ncls_pad(int out_channels);
int head_partials_per_img(const Geo& g);
cudaError_t launch_head_pack(const float* w_oihw, int out_channels, float* packed, cudaStream_t st);
cudaError_t launch_teacher_head(const Geo& g, const Workspace& ws, const Ptr5& f_cls, const Ptr5& f_reg, const float* w_cls,
                                const float* w_reg, const float* b_cls, const float* b_reg, const float* scale,
                                const MPtr5* o_cls, const MPtr5* o_box, int32_t* cls_count, int32_t* box_count,
                                cudaStream_t st);
cudaError_t launch_ers_flags(const Geo& g, const Workspace& ws, int tiles_per_img, float* thr, uint8_t* sel_flags,
                             int32_t* cls_count, int32_t* box_count, cudaStream_t st);
cudaError_t launch_ers_lists(const Geo& g, const Workspace& ws, int32_t* cls_inds, int32_t* cls_count, int32_t* box_inds,
                             int32_t* box_count, const float* thr, uint8_t* sel_flags, cudaStream_t st);
// ---------------------------------------------------------------- host launchers (one per .cu)
cudaError_t launch_atss_candidates(const Geo& g, const Workspace& ws, const float* gt_boxes, const int32_t* gt_offsets,
                                   const int32_t* pad_hw, cudaStream_t st);
cudaError_t launch_atss(const Geo& g, const Workspace& ws, const float* gt_boxes, const int64_t* gt_labels,
                        const int32_t* gt_offsets, const int32_t* pad_hw, int32_t* gt_inds, int32_t* num_pos,
                        cudaStream_t st);
// atss_finalize + pos_prepass in one launch (the step's student-side chain is latency bound)
cudaError_t launch_assign_avg(const Geo& g, const Workspace& ws, const Ptr5& s_cls, const Ptr5& s_box,
                              const float* gt_boxes, const int64_t* gt_labels, const int32_t* gt_offsets,
                              const int32_t* pad_hw, int32_t* gt_inds, int32_t* num_pos, float* avg,
                              const ExchangeInfo* xchg, cudaStream_t st);   // xchg: post the factors to the peers
cudaError_t launch_avg(const Geo& g, const Workspace& ws, const Ptr5& s_cls, const Ptr5& s_box,
                       const float* gt_boxes, const int64_t* gt_labels, const int32_t* gt_offsets,
                       const int32_t* gt_inds, const int32_t* num_pos, float* avg, cudaStream_t st);
cudaError_t launch_nms(const Geo& g, const Workspace& ws, const int32_t* box_inds, const int32_t* box_count,
                       const int32_t* pad_hw, float iou_thr, int32_t* keep, int32_t* keep_count, uint8_t* sel_flags,
                       cudaStream_t st, cudaEvent_t prepped, cudaEvent_t resolved);

struct LossArgs {
  Ptr5 s_cls, s_box, t_cls, t_box;
  MPtr5 g_cls, g_box;
  const float* gt_boxes;
  const int64_t* gt_labels;
  const int32_t* gt_offsets;
  const int32_t* pad_hw;
  const int32_t* gt_inds;
  const int32_t* num_pos;
  const int32_t* cls_inds;
  const int32_t* cls_count;
  const uint8_t* sel_flags;
  const int32_t* box_inds;
  const int32_t* box_count;
  const int32_t* keep;
  const int32_t* keep_count;
  const float* avg;
  const float* upstream;
  const unsigned int* skip_flag;   // non-NULL: kernels return at once when *skip_flag == 0
  float* losses;
  float dlw;
  ExchangeInfo xchg;   // world > 1: the student pass waits for the peers' factors the assignment prepass posted
};
struct LossStreams {
  cudaEvent_t sel_ready;              // may be null: ERS selection already ordered before the caller's stream
  cudaEvent_t nms_done;               // may be null: NMS already ordered before the caller's stream
};
cudaError_t launch_loss(const Geo& g, const Workspace& ws, const LossArgs& a, cudaStream_t st, const LossStreams* ls);

}  // namespace erd
