// The dense student pass: ONE persistent kernel that reads every student logit once and writes
// every gradient element once -- QFL forward/backward on the new-class channels, the class-
// response L2 rows of the ERS set, the GIoU/DFL rows of the positives, the DFL-distribution KL
// rows of the ERS box candidates and the structural zeros of everything else.
// Reference: GFLHeadIncrementERD.loss_by_feat_single / distill_loss_by_image_single
// (dense_heads/gfl_head_increment_erd.py:142-322) and their autograd backward; closed forms in
// loss_math.cuh.
//
// Structure (one CTA per SM).  The CTA is four independent TEAMS, each its own
// pipeline over every fourth tile of the CTA's sequence (a tile = 32 anchors x all channels):
//   IO warp (1 per team)    double-buffers the team's two shared-memory slots.  Per tile: two 2-D TMA
//                           loads (cp.async.bulk.tensor) bring the [C x 32] class tile and the [68 x 32]
//                           box tile of the student; the per-anchor roles (assignment, ERS flags) go
//                           into the slot's header, the positives' records and the teacher columns of
//                           the tile's ERS anchors into its staging area (cp.async).  When the
//                           consumers are done with a slot, two TMA stores write it to the gradient
//                           tensors and the next load is issued.
//   4 consumer warps        transform the tile IN PLACE (logits -> gradients).  Dense part: lane =
//                           anchor column, warp = quarter of the channels, loads batched ahead of the
//                           arithmetic.  Sparse part: one warp per special column (ERS row / positive /
//                           box candidate) of the tile.
// Teams never synchronise with each other, and inside a team the only synchronisation is the pair of
// mbarriers per slot (full: IO -> consumers, done: consumers -> IO), always waited on in tile order,
// so no barrier phase can be skipped.
// Pyramid levels whose rows are not 16-byte aligned (H*W % 4 != 0: a few % of the anchors) cannot
// have a tensor map: the IO warp moves their tiles with 4-byte cp.async copies / plain stores;
// the consumers see no difference.
#include <cuda.h>   // CUtensorMap types; the encoder itself is looked up through the runtime (below)

#include <cstdio>
#include <cstdlib>

#include "loss_math.cuh"

namespace erd {

constexpr int kBT = 32;                                   // anchors per tile: one per lane
#ifndef ERD_STUDENT_TEAMS
#define ERD_STUDENT_TEAMS 4
#endif
#ifndef ERD_STUDENT_SPT
#define ERD_STUDENT_SPT 2
#endif
constexpr int kTeams = ERD_STUDENT_TEAMS;                 // independent pipelines per CTA
constexpr int kSPT = ERD_STUDENT_SPT;                     // slots per team
constexpr int kTeamWarps = 4;                             // consumer warps per team
constexpr int kConsumerWarps = kTeams * kTeamWarps;
constexpr int kConsumers = 32 * kConsumerWarps;
constexpr int kBThreads = kConsumers + 32 * kTeams;       // + one IO warp per team
constexpr int kSlots = kSPT * kTeams;
constexpr int kStageItems = 12;                           // special columns per tile whose records / teacher columns are staged in the slot
constexpr int kStudentFreeSms = 24;                       // SMs left to the kernels that run beside this pass
constexpr int kQflChunk = 5;                              // QFL elements a thread loads ahead of the arithmetic
static_assert(kBoxCh % kTeamWarps == 0, "box rows are split evenly over the team's warps");
constexpr int kBoxPerWarp = kBoxCh / kTeamWarps;

constexpr unsigned kRoleValid = 1u, kRolePos = 2u, kRoleCls = 4u, kRoleCand = 8u;
constexpr unsigned kRoleSpecial = kRolePos | kRoleCls | kRoleCand;

// Header of a slot, behind the tile's logits.  `rec` is filled by asynchronous copies (cp.async)
// that complete on the slot's full barrier, like the teacher columns behind the header.
struct __align__(32) TileHeader {
  PosRec rec[kStageItems];          // records of the tile's first positives (index: rec_of[item])
  float kd[kBT];                    // weighted KL of the items that are box candidates (consumers -> IO warp)
  int n, l, hw0, cnt;               // the tile: image, level, first anchor of the level, anchors
  int n_items, cls_k, tma;          // cls_k: K_cls of the tile's image (class-response normaliser)
  float inv_kc;                     // 1 / (K_cls * ori)
  unsigned char role[kBT];
  unsigned char item_col[kBT];      // special columns of the tile, ascending
  unsigned char col_item[kBT];      // column -> its item index
  unsigned char rec_of[kBT];        // item -> index into rec[], 255: not staged
};

static_assert(sizeof(TileHeader) % 16 == 0, "the teacher staging behind the header is the target of 16-byte bulk copies");

struct StudentArgs {
  Ptr5 t_cls, t_box;
  MPtr5 g_cls, g_box;        // used by the tiles of levels without a tensor map
  Ptr5 s_cls, s_box;
  const int32_t* gt_inds;
  const uint8_t* sel_flags;
  const int32_t* cls_count;
  const float* avg;
  ExchangeInfo xchg;                 // world > 1: wait for the peers' factors in the prologue (posted by the assignment prepass)
  const float* upstream;
  const unsigned int* skip_flag;
  float dlw;
  int tiles_per_img, total_tiles;
  int slot_bytes, hdr_off, tst_off;  // slot layout: [rows x 32 logits][TileHeader][tstage_items x stash_pitch(ori) teacher logits]
  int tstage_items;                  // items per tile whose teacher column is staged (more: read from global)
  int tcol_pitch;                    // floats per staged teacher column == per stash row (a multiple of 4)
  int lvl_tile_start[kLevels + 1];   // prefix of ceil(hw / kBT)
  int use_tma[kLevels];
};

struct __align__(64) StudentMaps {
  CUtensorMap s_cls[kLevels], s_box[kLevels], g_cls[kLevels], g_box[kLevels];   // [C x 32] / [68 x 32] tiles
  CUtensorMap s_new[kLevels];                                                   // [cn x 32]: the new-class rows only
};

struct BTile {
  int n, l, hw0, cnt;
};

__device__ __forceinline__ BTile b_tile(const Geo& g, const StudentArgs& A, int t) {
  BTile b;
  const int q = t / A.tiles_per_img;
  const int r = t - q * A.tiles_per_img;
  b.n = g.n_img - 1 - q;   // last image first: its teacher rows are the freshest in L2 (the teacher pass ran 0 .. N-1)
  b.l = 0;
#pragma unroll
  for (int i = 1; i < kLevels; ++i) b.l += (r >= A.lvl_tile_start[i]) ? 1 : 0;
  b.hw0 = (r - A.lvl_tile_start[b.l]) * kBT;
  b.cnt = min(kBT, g.hw[b.l] - b.hw0);
  return b;
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait_parity(unsigned long long* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// this thread's earlier cp.async copies arrive on `bar` when they have landed (the barrier's count includes it)
__device__ __forceinline__ void cp_async_arrive(unsigned long long* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar,
                                            unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(smem_addr(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_addr(bar)), "l"(policy) : "memory");
}
// 1-D bulk copy global -> shared, completing on `bar` (addresses and size multiples of 16 B)
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, const void* src,
                                             unsigned long long policy) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;"
               ::"l"(map), "r"(c0), "r"(c1), "r"(smem_addr(src)), "l"(policy) : "memory");
}

// Box rows of one special column (a positive and / or an ERS box candidate), by one whole warp, in
// place in the tile.  lane = side * 8 + b holds bins b, b + 8 (b == 0: also 16) of its side.
// `tb` / `tbs`: the teacher's box column of this anchor (staged in the slot: stride 1; global: stride H*W).
__device__ __forceinline__ void item_box_rows(const Geo& g, const Workspace& ws, const StudentArgs& A, const BTile& b,
                                              float* box, const TileHeader* hd, float* hd_kd, const float* tb, size_t tbs,
                                              int it, int icol, int hwI, bool is_pos, bool is_cand, float w_kd, const float (&k4)[4],
                                              float inv_T, size_t ga0, int lane) {
  constexpr int pitch = kBT;
  const int side = lane >> 3, bb = lane & 7;
  const float* brow_in = box + (size_t)(side * kBins) * pitch + icol;
  float* brow = box + (size_t)(side * kBins) * pitch + icol;
  int jb[3];
  bool ok[3];
  float zs[3], out[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    jb[i] = bb + 8 * i;
    ok[i] = jb[i] < kBins;
    zs[i] = ok[i] ? brow_in[(size_t)jb[i] * pitch] : -INFINITY;
    out[i] = 0.f;
  }
  if (is_cand) {   // DFL-distribution KL at temperature T (:204-221, kd_loss.py:12-37), written as if the NMS kept it
    float a[3], t[3];
    float ms = -INFINITY, mt = -INFINITY;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int ch = side * kBins + jb[i];
      const float zt = ok[i] ? tb[(size_t)ch * tbs] : -INFINITY;
      a[i] = zs[i] * inv_T;
      t[i] = zt * inv_T;
      ms = fmaxf(ms, a[i]);
      mt = fmaxf(mt, t[i]);
    }
    ms = oct_max(ms);
    mt = oct_max(mt);
    float es[3], et[3], ss = 0.f, st = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      a[i] -= ms;
      t[i] -= mt;
      es[i] = ok[i] ? __expf(a[i]) : 0.f;
      et[i] = ok[i] ? __expf(t[i]) : 0.f;
      ss += es[i];
      st += et[i];
    }
    ss = oct_sum(ss);
    st = oct_sum(st);
    const float lss = __logf(ss), lst = __logf(st);
    const float iss = 1.0f / ss, ist = 1.0f / st;
    float kl = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (ok[i]) {
        const float lps = a[i] - lss, lpt = t[i] - lst;          // log_softmax(z / T)
        const float ps = es[i] * iss, pt = et[i] * ist;          // softmax(z / T)
        if (pt > 0.f) kl += pt * (lpt - lps);
        out[i] = w_kd * (ps - pt);
      }
    }
    kl = warp_sum(kl);   // over the four sides
    const float kT = g.T;
    const float scale = upstream_of(A.upstream, acc_dbox(g, b.n)) * k4[0];   // dlw * w_ld / 4 * (T^2 / bins) / T
#pragma unroll
    for (int i = 0; i < 3; ++i) out[i] *= scale;
    if (lane == 0) hd_kd[it] = w_kd * (kl * k4[1]);   // .mean(1) * T*T; the IO warp files it
  }
  if (is_pos) {   // GIoU + DFL rows of a positive, through the softmax Jacobian (:285-310)
    const int ri = hd->rec_of[it];
    const PosRec rec = ri != 255 ? hd->rec[ri] : ws.pos_rec[ga0 + icol];
    float zm = fmaxf(fmaxf(zs[0], zs[1]), zs[2]);
    zm = oct_max(zm);
    float e[3], sum = 0.f, num = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      e[i] = ok[i] ? expf(zs[i] - zm) : 0.f;
      sum += e[i];
      num = fmaf((float)jb[i], e[i], num);
    }
    sum = oct_sum(sum);
    num = oct_sum(num);
    const float inv = 1.0f / sum;
    const float dmine = num * inv;                                   // Integral (:40-54,285)
    float d[4];
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) d[k2] = __shfl_sync(0xffffffffu, dmine, k2 * 8);
    const PosGeom pg = pos_geom(d, hwI % g.w[b.l], hwI / g.w[b.l], (float)g.stride[b.l], rec.gt);
    const DflTarget tg = dfl_target(pg, side);
    if (rec.label >= 0) {
      const float gd = pos_side_giou_grad(pg, side);
      const float cb = upstream_of(A.upstream, acc_bbox(b.l)) * k4[2] * rec.w * gd;   // w_bbox / (1 + eps) / avg2
      const float cd = upstream_of(A.upstream, acc_dfl(b.l)) * k4[3] * rec.w;           // w_dfl / 4 / avg2
      float* prow = is_cand ? ws.pos_rows + ((size_t)b.n * g.pos_cap + rec.pslot) * kBoxCh + side * kBins : nullptr;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (ok[i]) {
          const float pj = e[i] * inv;
          float gr = cb * pj * ((float)jb[i] - dmine);
          gr += cd * (tg.wl * (pj - (jb[i] == tg.yl ? 1.f : 0.f)) + tg.wr * (pj - (jb[i] == tg.yl + 1 ? 1.f : 0.f)));
          if (prow) prow[jb[i]] = gr;          // the take-back pass restores it if the NMS drops the candidate
          out[i] += gr;
        }
      }
    } else if (is_cand) {   // assigned to a GT outside the new-class range: no box loss
      float* prow = ws.pos_rows + ((size_t)b.n * g.pos_cap + rec.pslot) * kBoxCh + side * kBins;
#pragma unroll
      for (int i = 0; i < 3; ++i)
        if (ok[i]) prow[jb[i]] = 0.f;
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (ok[i]) brow[(size_t)jb[i] * pitch] = out[i];
}

// What a consumer warp needs to know beyond the tile itself.
struct ConsumerCtx {
  const Geo& g;
  const Workspace& ws;
  const StudentArgs& A;
  int lane, q, oq, cq;           // q: warp in its team = quarter of the channels
  float inv_avg1, avg2, inv_T;
  float kd_scale, kl_scale, cb0, cd0;   // per-kernel constants of the special columns (no division per column)
};

// One tile, one consumer warp: the dense part of this warp's quarter of the channels and the items
// of the tile that fall to this warp, in place in the slot.  `j`: index of the tile in the team's sequence.
__device__ __forceinline__ void consume_tile(const ConsumerCtx& cc, int j, float* data, TileHeader* hd, const float* tst,
                                             float& qfl_part, float& dcls_part) {
  const Geo& g = cc.g;
  const Workspace& ws = cc.ws;
  const StudentArgs& A = cc.A;
  const int lane = cc.lane, q = cc.q;
  const int C = g.C, ori = g.ori, cn = g.cn;
  BTile b;
  b.n = hd->n;
  b.l = hd->l;
  b.hw0 = hd->hw0;
  b.cnt = hd->cnt;
  const size_t ga0 = (size_t)b.n * g.A + g.start[b.l] + b.hw0;
  // Dense part and sparse items touch disjoint addresses of the tile (the dense part skips the
  // rows a special column's item owns), so nothing orders them inside a tile.
  // ------------------------------------------------------------ dense part: lane = column
  {
    const unsigned role = hd->role[lane];
    const float lw = (role & kRoleValid) ? 1.0f : 0.0f;                // label_weights, gfl_head.py:650-655,663
    const float gs = lw * (upstream_of(A.upstream, acc_cls(b.l)) * g.w_cls * cc.inv_avg1);
    int label = -1;
    float score = 0.f;
    if (role & kRolePos) {   // rare
      const int ri = hd->rec_of[hd->col_item[lane]];
      const PosRec* rec = ri != 255 ? &hd->rec[ri] : ws.pos_rec + ga0 + lane;
      label = rec->label;
      score = rec->score;
    }
    float* ncol = data + (size_t)ori * kBT + lane;       // the new-class rows of this column
    const int c0 = q * cc.cq, c1 = min(c0 + cc.cq, cn);
    const bool own_label = label >= c0 && label < c1;    // a positive's label channel lies in this thread's share
    const float x_label = own_label ? ncol[label * kBT] : 0.f;
    float loss = 0.f;
    for (int c = c0; c < c1; c += kQflChunk) {   // QFL, every element as a negative first: branch free (:260-261,317-320)
      float x[kQflChunk];
#pragma unroll
      for (int i = 0; i < kQflChunk; ++i) x[i] = c + i < c1 ? ncol[(c + i) * kBT] : -INFINITY;   // (-inf: loss 0, gradient 0)
#pragma unroll
      for (int i = 0; i < kQflChunk; ++i) {
        const QflTerm tn = qfl_neg(x[i]);
        loss += tn.loss;
        x[i] = gs * tn.grad;
      }
#pragma unroll
      for (int i = 0; i < kQflChunk; ++i)
        if (c + i < c1) ncol[(c + i) * kBT] = x[i];
    }
    if (own_label) {   // ... then the label channel of the (rare) positive is redone with its soft target
      const QflTerm tp = qfl_pos(x_label, score), tn = qfl_neg(x_label);
      loss += tp.loss - tn.loss;
      ncol[label * kBT] = gs * tp.grad;
    }
    qfl_part = fmaf(lw, loss, qfl_part);
    // structural zeros: the old-class rows of columns without a class-response row or a box
    // candidate's weight to read, the box rows of columns that are neither positive nor candidate
    if (!(role & (kRoleCls | kRoleCand))) {
      const int o0 = q * cc.oq, o1 = min(o0 + cc.oq, ori);
      float* ocol = data + lane;
#pragma unroll 5
      for (int c = o0; c < o1; ++c) ocol[c * kBT] = 0.f;
    }
    if (!(role & (kRolePos | kRoleCand))) {
      float* bcol = data + (size_t)(C + q * kBoxPerWarp) * kBT + lane;
#pragma unroll
      for (int r = 0; r < kBoxPerWarp; ++r) bcol[r * kBT] = 0.f;
    }
  }
  // ------------------------------------------------------------ sparse items: one WARP per special column
  // Item i of the team's tile j goes to its warp (i + j) % 4.  Lane layout for the box rows: lane = side * 8 + b,
  // holding bins b, b + 8 (and 16 for b == 0) of its side; reductions over a side are 8-lane shuffles.
  const int n_items = hd->n_items;
  const float k4[4] = {cc.kd_scale, cc.kl_scale, cc.cb0, cc.cd0};
  for (int it = (q - j) & (kTeamWarps - 1); it < n_items; it += kTeamWarps) {
    const int icol = hd->item_col[it];
    const unsigned irole = hd->role[icol];
    const bool is_cls = (irole & kRoleCls) != 0u, is_cand = (irole & kRoleCand) != 0u, is_pos = (irole & kRolePos) != 0u;
    const int hwI = b.hw0 + icol;
    // the teacher's column of this anchor: staged behind the header by the IO warp, or (more items in the tile
    // than the staging area holds) straight from global memory
    const bool staged = it < A.tstage_items;
    const int HW = g.hw[b.l];
    const float* tc = staged ? tst + (size_t)it * A.tcol_pitch : A.t_cls.p[b.l] + (size_t)b.n * ori * HW + hwI;
    const float* tb = staged ? tc + ori : A.t_box.p[b.l] + (size_t)b.n * kBoxCh * HW + hwI;
    const size_t ts = staged ? 1 : (size_t)HW;
    float* box = data + (size_t)C * kBT;
    if (is_cls || is_cand) {
      // old-class rows of the column: lane owns channels lane, lane + 32, ...
      const float scale_dc = upstream_of(A.upstream, acc_dcls(b.n)) * A.dlw * 2.0f * hd->inv_kc;   // 2 / (K ori)
      float mx_old = -INFINITY;
      for (int c = lane; c < ori; c += 32) {
        const float xs = data[c * kBT + icol];
        mx_old = fmaxf(mx_old, xs);
        float gr = 0.f;
        if (is_cls) {   // class-response L2 (:181-186,324-332): g = 2 (x_s - x_t) / (K ori)
          const float df = xs - tc[(size_t)c * ts];
          dcls_part = fmaf(df, df, dcls_part);
          gr = scale_dc * df;
        }
        data[c * kBT + icol] = gr;
      }
      if (is_pos || is_cand) {
        const float w_kd = sigmoid_ref(warp_max(mx_old));                                      // :217-218
        item_box_rows(g, ws, A, b, box, hd, hd->kd, tb, ts, it, icol, hwI, is_pos, is_cand, w_kd, k4, cc.inv_T, ga0, lane);
      }
    } else {
      item_box_rows(g, ws, A, b, box, hd, hd->kd, tb, ts, it, icol, hwI, is_pos, false, 0.f, k4, cc.inv_T, ga0, lane);
    }
  }
}

// ----------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(kBThreads, 1)
student_pass_kernel(Geo g, Workspace ws, StudentArgs A, const __grid_constant__ StudentMaps maps) {
  if (A.skip_flag && *A.skip_flag == 0u) return;
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ __align__(8) unsigned long long s_full[kSlots], s_done[kSlots];
  // loss sums of this CTA: [kLevels] QFL + [n_img] class-response squares, behind the slots (flushed once, at the end:
  // a global atomic in front of a tile's release-arrive would hold the slot for a full memory round trip)
  double* s_loss = reinterpret_cast<double*>(s_raw + (size_t)kSlots * A.slot_bytes);
  const int C = g.C, ori = g.ori, cn = g.cn;
  const int rows = C + kBoxCh;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(&s_full[s], 1 + 32);        // the IO warp's lane 0 (after arming the TMA bytes) + its 32 lanes' cp.async batches
      mbar_init(&s_done[s], kTeamWarps);    // one arrival per consumer warp of the team
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < kLevels + g.n_img; i += kBThreads) s_loss[i] = 0.0;
  __syncthreads();
  // The two normalisers.  Data-parallel runs: the mean over ranks of what the assignment prepass posted
  // (reduce_mean, dist_utils.py:59-65) -- every CTA waits for the peers' slots itself and computes the same bits;
  // CTA 0 also stores them for the finalize step and the caller.
  float avg0 = A.avg[0], avg1 = A.avg[1];
  if (A.xchg.world > 1) {
    __shared__ float s_xv[kMaxRanks][2];
    __shared__ unsigned int s_xe;
    if (threadIdx.x == 0)
      s_xe = *reinterpret_cast<volatile unsigned int*>(A.xchg.peers.buf[A.xchg.rank] + kExchangeSlotBytes);   // this step's epoch
    __syncthreads();
    exchange_wait(A.xchg, s_xe, s_xv, avg0, avg1);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      float* out = const_cast<float*>(A.avg);
      out[0] = avg0;
      out[1] = avg1;
    }
  }

  if (warp >= kConsumerWarps) {
    // ================================================================== IO warp of team `team`
    const int team = warp - kConsumerWarps;
    // L2 hints (measured, gpurun_out/r3 sweeps: 0.172 -> 0.167 ms per step against evict_first for everything): the
    // gradients are the lines somebody reads next (the conv backward), so they may stay; the logits are read once
    unsigned long long pol_load, pol_store;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_load));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_store));
    // Tile jj of the team is done: write the slot to the gradient tensors; returns when the slot may be overwritten.
    auto drain = [&](int jj) {
      const int s = team * kSPT + (jj % kSPT);
      mbar_wait_parity(&s_done[s], (uint32_t)(jj / kSPT) & 1u);
      const unsigned char* base = s_raw + (size_t)s * A.slot_bytes;
      const float* data = reinterpret_cast<const float*>(base);
      const TileHeader* hd = reinterpret_cast<const TileHeader*>(base + A.hdr_off);
      const int n = hd->n, l = hd->l, hw0 = hd->hw0, cnt = hd->cnt, n_items = hd->n_items;
      const size_t ga0 = (size_t)n * g.A + g.start[l] + hw0;
      // the candidates' weighted KL values of this tile -> ws.kd_loss (read by the take-back pass)
      for (int it = lane; it < n_items; it += 32) {
        const int icol = hd->item_col[it];
        if (hd->role[icol] & kRoleCand) ws.kd_loss[ga0 + icol] = hd->kd[it];
      }
      if (A.use_tma[l]) {
        if (lane == 0) {
          tma_store_2d(&maps.g_cls[l], hw0, n * C, data, pol_store);
          tma_store_2d(&maps.g_box[l], hw0, n * kBoxCh, data + (size_t)C * kBT, pol_store);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the engine has read the slot
        }
      } else if (lane < cnt) {   // rows of this level are not 16 B aligned: plain stores, 128 B per row and warp
        const int HW = g.hw[l];
        float* gc = A.g_cls.p[l] + (size_t)n * C * HW + hw0 + lane;
#pragma unroll 8
        for (int r = 0; r < C; ++r) __stcs(gc + (size_t)r * HW, data[r * kBT + lane]);
        float* gb = A.g_box.p[l] + (size_t)n * kBoxCh * HW + hw0 + lane;
#pragma unroll 4
        for (int r = 0; r < kBoxCh; ++r) __stcs(gb + (size_t)r * HW, data[(C + r) * kBT + lane]);
      }
      __syncwarp();
    };
    int j = 0;
    for (;; ++j) {
      const int t = blockIdx.x + (team + kTeams * j) * gridDim.x;
      if (t >= A.total_tiles) break;
      const BTile b = b_tile(g, A, t);
      const int HW = g.hw[b.l];
      // the tile's roles: requested before the slot is drained, so their latency hides behind the store
      const bool in = lane < b.cnt;
      const size_t ga = (size_t)b.n * g.A + g.start[b.l] + b.hw0 + lane;
      const int gi = in ? A.gt_inds[ga] : -1;
      const unsigned fl = in ? (unsigned)A.sel_flags[ga] : 0u;
      const unsigned srow = in ? (unsigned)ws.t_slot[ga] : 0u;   // the anchor's stash row + 1 (teacher pass), 0: none
      const int kcls = A.cls_count[b.n];
      if (j >= kSPT) drain(j - kSPT);
      const int s = team * kSPT + (j % kSPT);
      unsigned char* base = s_raw + (size_t)s * A.slot_bytes;
      float* data = reinterpret_cast<float*>(base);
      TileHeader* hd = reinterpret_cast<TileHeader*>(base + A.hdr_off);
      float* tst = reinterpret_cast<float*>(base + A.tst_off);
      unsigned long long* full = &s_full[s];
      const bool tma = A.use_tma[b.l] != 0;
      const unsigned role = (gi >= 0 ? kRoleValid : 0u) | (gi > 0 ? kRolePos : 0u) | ((fl & 1u) ? kRoleCls : 0u) |
                            ((fl & 2u) ? kRoleCand : 0u);
      // Only the new-class rows are read by every column (QFL).  The old-class rows are read by ERS columns
      // alone and the box rows by positives / box candidates alone, everything else in them is overwritten with
      // zeros unread: a tile without such a column does not fetch them (27 % resp. 46 % of the tile's bytes; on
      // clustered, trained-teacher responses most tiles have none).
      const bool need_old = __ballot_sync(0xffffffffu, (role & (kRoleCls | kRoleCand)) != 0u) != 0u;
      const bool need_box = __ballot_sync(0xffffffffu, (role & (kRolePos | kRoleCand)) != 0u) != 0u;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this warp's generic reads of the slot (drain) before the bulk writes
      if (tma) {
        if (lane == 0) {
          mbar_expect_tx(full, (uint32_t)(((need_old ? C : cn) + (need_box ? kBoxCh : 0)) * kBT * sizeof(float)));
          if (need_old) tma_load_2d(data, &maps.s_cls[b.l], b.hw0, b.n * C, full, pol_load);
          else tma_load_2d(data + (size_t)ori * kBT, &maps.s_new[b.l], b.hw0, b.n * C + ori, full, pol_load);
          if (need_box) tma_load_2d(data + (size_t)C * kBT, &maps.s_box[b.l], b.hw0, b.n * kBoxCh, full, pol_load);
        }
      } else {
        // rows not 16 B aligned: 4-byte asynchronous copies, a warp-wide 128 B request per row; lanes past
        // the level's end zero their column
        const float* sc = A.s_cls.p[b.l] + (size_t)b.n * C * HW + b.hw0 + lane;
        const float* sb = A.s_box.p[b.l] + (size_t)b.n * kBoxCh * HW + b.hw0 + lane;
#pragma unroll 8
        for (int r = need_old ? 0 : ori; r < C; ++r) {
          if (in) cp_async_4(data + r * kBT + lane, sc + (size_t)r * HW);
          else data[r * kBT + lane] = 0.f;
        }
        if (need_box) {
#pragma unroll 4
          for (int r = 0; r < kBoxCh; ++r) {
            if (in) cp_async_4(data + (C + r) * kBT + lane, sb + (size_t)r * HW);
            else data[(C + r) * kBT + lane] = 0.f;
          }
        }
      }
      // header: roles, the list of special columns, the positives' records
      const bool special = (role & kRoleSpecial) != 0u;
      const unsigned lt = (1u << lane) - 1u;
      const unsigned m = __ballot_sync(0xffffffffu, special);
      const int item = special ? __popc(m & lt) : 255;
      hd->role[lane] = (unsigned char)role;
      hd->col_item[lane] = (unsigned char)item;
      if (special) hd->item_col[item] = (unsigned char)lane;
      const bool pos = (role & kRolePos) != 0u;
      const unsigned mp = __ballot_sync(0xffffffffu, pos);
      if (pos) {
        const int ri = __popc(mp & lt);
        hd->rec_of[item] = (unsigned char)(ri < kStageItems ? ri : 255);
        if (ri < kStageItems) {
          const PosRec* src = ws.pos_rec + ga;
          cp_async_16(reinterpret_cast<char*>(&hd->rec[ri]), reinterpret_cast<const char*>(src));
          cp_async_16(reinterpret_cast<char*>(&hd->rec[ri]) + 16, reinterpret_cast<const char*>(src) + 16);
        }
      }
      if (lane == 0) {
        hd->n = b.n;
        hd->l = b.l;
        hd->hw0 = b.hw0;
        hd->cnt = b.cnt;
        hd->n_items = __popc(m);
        hd->cls_k = kcls;
        hd->inv_kc = 1.0f / ((float)kcls * (float)ori);   // K = 0: inf, and the (non-existent) rows' scale with it
        hd->tma = tma ? 1 : 0;
      }
      // The teacher's columns of the tile's first items: one bulk copy of the anchor's stash row (written by
      // the teacher pass while it had the tile in shared memory); an anchor the provisional thresholds missed
      // is gathered from the NCHW tensors, 4 bytes per row.
      unsigned mm = m;
      for (int it = 0; mm != 0u && it < A.tstage_items; ++it) {   // warp-uniform
        const int icol = __ffs(mm) - 1;
        mm &= mm - 1u;
        const unsigned irole = __shfl_sync(0xffffffffu, role, icol);
        const unsigned irow = __shfl_sync(0xffffffffu, srow, icol);
        float* dst = tst + (size_t)it * A.tcol_pitch;
        if (!(irole & (kRoleCls | kRoleCand))) continue;
        if (irow != 0u) {
          if (lane == 0) {
            const uint32_t bytes = (uint32_t)A.tcol_pitch * (uint32_t)sizeof(float);
            mbar_expect_tx(full, bytes);
            bulk_load(dst, ws.t_stash + ((size_t)b.n * kStashRows + (irow - 1u)) * A.tcol_pitch, bytes, full);
          }
          continue;
        }
        if (irole & kRoleCls) {
          const float* src = A.t_cls.p[b.l] + (size_t)b.n * ori * HW + b.hw0 + icol;
          for (int c = lane; c < ori; c += 32) cp_async_4(dst + c, src + (size_t)c * HW);
        }
        if (irole & kRoleCand) {
          const float* src = A.t_box.p[b.l] + (size_t)b.n * kBoxCh * HW + b.hw0 + icol;
          for (int r = lane; r < kBoxCh; r += 32) cp_async_4(dst + ori + r, src + (size_t)r * HW);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(full);   // releases the header written by all lanes
#ifndef ERD_RACECHECK
      cp_async_arrive(full);              // every lane: its copies (possibly none) count towards the slot
#else
      // compute-sanitizer racecheck tracks cp.async only through wait_group: this build (scripts/sanitize.sh)
      // completes the copies synchronously so that the tool can check everything else about the slot protocol
      asm volatile("cp.async.wait_all;" ::: "memory");
      mbar_arrive(full);
#endif
    }
#pragma unroll
    for (int d = kSPT; d >= 1; --d)
      if (j >= d) drain(j - d);
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    return;
  }

  // ==================================================================== consumers
  // The consumer warps of a team never synchronise with each other: inside a tile the dense part and
  // the items own disjoint addresses, and a warp that is done with its share of a tile moves on to the
  // team's next slot.  Loss sums travel in registers and are flushed (warp shuffle + one shared-memory
  // fp64 atomic) when the level / image of the team's tile sequence changes.
  const int team = warp / kTeamWarps, q = warp % kTeamWarps;
  const float inv_avg1 = 1.0f / (float)((double)avg0 + (double)kEps32);   // losses/utils.py:60-61
  const float avg2 = fmaxf(avg1, 1.0f);                                   // :407 clamp_(min=1)
  const ConsumerCtx cc{g, ws, A, lane, q, (ori + kTeamWarps - 1) / kTeamWarps, (cn + kTeamWarps - 1) / kTeamWarps,
                       inv_avg1, avg2, 1.0f / g.T,
                       A.dlw * g.w_ld / 4.0f * (g.T * g.T / (float)kBins) / g.T, g.T * g.T / (float)kBins,
                       g.w_bbox / (1.0f + kEps32) / avg2, g.w_dfl / 4.0f / avg2};
  int cur_img = -1, cur_lvl = -1;
  float dcls_part = 0.f;   // sum (x_s - x_t)^2 of the current image, this thread
  float qfl_part = 0.f;    // QFL loss sum of the current (image, level), this thread
  auto flush = [&]() {
    const float vq = warp_sum(qfl_part), vd = warp_sum(dcls_part);
    if (lane == 0 && cur_img >= 0) {
      if (vq != 0.f) atomicAdd(s_loss + cur_lvl, (double)vq);
      if (vd != 0.f) atomicAdd(s_loss + kLevels + cur_img, (double)vd);
    }
    qfl_part = 0.f;
    dcls_part = 0.f;
  };
  for (int j = 0;; ++j) {
    const int t = blockIdx.x + (team + kTeams * j) * gridDim.x;
    if (t >= A.total_tiles) break;
    const int s = team * kSPT + (j % kSPT);
    unsigned char* base = s_raw + (size_t)s * A.slot_bytes;
    float* data = reinterpret_cast<float*>(base);
    TileHeader* hd = reinterpret_cast<TileHeader*>(base + A.hdr_off);
    const float* tst = reinterpret_cast<const float*>(base + A.tst_off);
    mbar_wait_parity(&s_full[s], (uint32_t)(j / kSPT) & 1u);
    const int n = hd->n, l = hd->l;
    if (n != cur_img || l != cur_lvl) {   // warp-uniform
      flush();
      cur_img = n;
      cur_lvl = l;
    }
    consume_tile(cc, j, data, hd, tst, qfl_part, dcls_part);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> the TMA store
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_done[s]);
  }
  flush();
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");   // consumers only: every warp has flushed
  for (int i = threadIdx.x; i < kLevels + g.n_img; i += kConsumers)
    if (s_loss[i] != 0.0) atomicAdd(ws.loss_acc + (i < kLevels ? acc_cls(i) : acc_dcls(i - kLevels)), s_loss[i]);
}

}  // namespace erd

namespace erd {

// ----------------------------------------------------------------------------- host side
// cuTensorMapEncodeTiled comes from the driver through the runtime's entry-point lookup, so the
// library links against nothing but cudart.
typedef CUresult (*TmaEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static TmaEncodeFn tma_encoder() {
  static TmaEncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (TmaEncodeFn)p;
  }
  return fn;
}

// [rows_total x hw] fp32 row-major, tiles of [box_rows x box_cols]; out-of-range columns read as zero / are not written
bool tma_encode_rows(void* map, const void* base, int hw, long long rows_total, int box_rows, int box_cols) {
  TmaEncodeFn enc = tma_encoder();
  if (!enc || box_rows > 256 || box_cols > 256 || (hw & 3) || ((uintptr_t)base & 15)) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)hw, (cuuint64_t)rows_total};
  const cuuint64_t strides[1] = {(cuuint64_t)hw * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  return enc((CUtensorMap*)map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t launch_student(const Geo& g, const Workspace& ws, const LossArgs& a, cudaStream_t st) {
  StudentArgs A;
  A.t_cls = a.t_cls;
  A.t_box = a.t_box;
  A.g_cls = a.g_cls;
  A.g_box = a.g_box;
  A.s_cls = a.s_cls;
  A.s_box = a.s_box;
  A.gt_inds = a.gt_inds;
  A.sel_flags = a.sel_flags;
  A.cls_count = a.cls_count;
  A.avg = a.avg;
  A.xchg = a.xchg;
  A.upstream = a.upstream;
  A.skip_flag = a.skip_flag;
  A.dlw = a.dlw;
  int tiles = 0;
  for (int l = 0; l < kLevels; ++l) {
    A.lvl_tile_start[l] = tiles;
    tiles += (g.hw[l] + kBT - 1) / kBT;
  }
  A.lvl_tile_start[kLevels] = tiles;
  A.tiles_per_img = tiles;
  A.total_tiles = tiles * g.n_img;
  // slot layout: [rows x 32 logits][TileHeader][tstage_items x stash_pitch(ori) teacher logits]; eight slots and the
  // loss sums must fit the 227 KB a CTA may have -- the teacher staging takes what is left, up to kStageItems
  const int rows = g.C + kBoxCh;
  const size_t data_bytes = (size_t)rows * kBT * sizeof(float);   // a multiple of 128
  const size_t tcol_bytes = (size_t)stash_pitch(g.ori) * sizeof(float);   // a multiple of 16
  const size_t tail_bytes = ((size_t)(kLevels + g.n_img) * sizeof(double) + 127) & ~(size_t)127;
  const size_t budget = (size_t)227 * 1024 - 1024 - tail_bytes;
  const size_t fixed = data_bytes + sizeof(TileHeader);
  if (fixed * kSlots > budget) return cudaErrorInvalidValue;   // num_classes too large for this tiling
  size_t stage = ((budget / kSlots) & ~(size_t)127) - fixed;
  int tstage = (int)(stage / tcol_bytes);
  if (tstage > kStageItems) tstage = kStageItems;
  A.tstage_items = tstage;
  A.tcol_pitch = stash_pitch(g.ori);
  A.hdr_off = (int)data_bytes;
  A.tst_off = (int)fixed;
  A.slot_bytes = (int)((fixed + (size_t)tstage * tcol_bytes + 127) & ~(size_t)127);
  // tensor maps, cached on the pointers they were built for (the training loop reuses its buffers)
  struct MapCache {
    const void* key[4 * kLevels];
    int hw[kLevels], n_img, C;
    int use_tma[kLevels];
    StudentMaps maps;
    bool valid = false;
  };
  static thread_local MapCache cache;
  const void* key[4 * kLevels];
  for (int l = 0; l < kLevels; ++l) {
    key[l] = a.s_cls.p[l];
    key[kLevels + l] = a.s_box.p[l];
    key[2 * kLevels + l] = a.g_cls.p[l];
    key[3 * kLevels + l] = a.g_box.p[l];
  }
  bool hit = cache.valid && cache.n_img == g.n_img && cache.C == g.C;
  for (int i = 0; hit && i < 4 * kLevels; ++i) hit = cache.key[i] == key[i];
  for (int l = 0; hit && l < kLevels; ++l) hit = cache.hw[l] == g.hw[l];
  if (!hit) {
    for (int l = 0; l < kLevels; ++l) {
      const long long rc = (long long)g.n_img * g.C, rb = (long long)g.n_img * kBoxCh;
      cache.use_tma[l] = g.vec[l] && tma_encode_rows(&cache.maps.s_cls[l], a.s_cls.p[l], g.hw[l], rc, g.C, kBT) &&
                         tma_encode_rows(&cache.maps.s_new[l], a.s_cls.p[l], g.hw[l], rc, g.cn, kBT) &&
                         tma_encode_rows(&cache.maps.s_box[l], a.s_box.p[l], g.hw[l], rb, kBoxCh, kBT) &&
                         tma_encode_rows(&cache.maps.g_cls[l], a.g_cls.p[l], g.hw[l], rc, g.C, kBT) &&
                         tma_encode_rows(&cache.maps.g_box[l], a.g_box.p[l], g.hw[l], rb, kBoxCh, kBT);
      cache.hw[l] = g.hw[l];
    }
    for (int i = 0; i < 4 * kLevels; ++i) cache.key[i] = key[i];
    cache.n_img = g.n_img;
    cache.C = g.C;
    cache.valid = true;
  }
  for (int l = 0; l < kLevels; ++l) A.use_tma[l] = cache.use_tma[l];
  static int sms = 0;
  static size_t smem_set = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const size_t smem = (size_t)kSlots * A.slot_bytes + tail_bytes;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(student_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    smem_set = smem;
  }
  // One CTA of this kernel fills an SM (registers, shared memory), so nothing else can share it.  The teacher-
  // side NMS chain runs BESIDE this pass and is all latency: it gets a few SMs of its own, or it would only
  // start when the first CTAs of this pass exit (measured: the whole chain then lands behind the pass).
  static const int free_sms = [] {
    const char* e = getenv("ERD_STUDENT_FREE_SMS");
    const int v = e ? atoi(e) : kStudentFreeSms;
    return v < 0 ? 0 : v;
  }();
  int grid = sms - free_sms;
  if (grid < 1) grid = 1;
  if (grid > A.total_tiles) grid = A.total_tiles;
  ERD_LAUNCH(kKStudent, st, (student_pass_kernel<<<grid, kBThreads, smem, st>>>(g, ws, A, cache.maps)));
  return cudaGetLastError();
}

}  // namespace erd
