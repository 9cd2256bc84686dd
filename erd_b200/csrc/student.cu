// The dense student pass: ONE persistent kernel that reads every student logit once and writes
// every gradient element once -- QFL forward/backward on the new-class channels, the class-
// response L2 rows of the ERS set, the GIoU/DFL rows of the positives, the DFL-distribution KL
// rows of the ERS box candidates and the structural zeros of everything else.
// Reference: GFLHeadIncrementERD.loss_by_feat_single / distill_loss_by_image_single
// (dense_heads/gfl_head_increment_erd.py:142-322) and their autograd backward; closed forms in
// loss_math.cuh.
//
// Structure (one CTA per SM, warp specialised):
//   load warp   : per tile of 64 anchors, two 2-D TMA loads (cp.async.bulk.tensor) bring the
//                 [C x 64] class tile and the [68 x 64] box tile of the student into a ring slot;
//                 meanwhile the warp gathers the tile's per-anchor roles (assignment, ERS flags,
//                 the positives' targets) into the slot's header.
//   8 consumer  : transform the tile IN PLACE in shared memory (logits -> gradients).  Dense part:
//   warps         thread = (anchor column, quarter of the channels), conflict free.  Sparse part:
//                 one quad of lanes per ERS row / positive / box candidate of the tile.
//   store warp  : two 2-D TMA stores write the slot to the gradient tensors; the slot returns to
//                 the load warp once the TMA engine has read it.
// Pyramid levels whose rows are not 16-byte aligned (H*W % 4 != 0: a few % of the anchors) cannot
// have a tensor map; their tiles take the same path with plain loads / stores by the consumers.
#include <cuda.h>   // CUtensorMap types; the encoder itself is looked up through the runtime (below)

#include <cstdio>
#include <cstdlib>

#include "loss_math.cuh"

namespace erd {

constexpr int kBT = 64;                    // anchors per tile
constexpr int kBConsumers = 512;           // 16 consumer warps ...
constexpr int kBTeams = 4;                 // ... in 4 teams; tile k of the CTA is consumed by team k % 4
constexpr int kBTeamWarps = kBConsumers / 32 / kBTeams;
constexpr int kBTeamThreads = 32 * kBTeamWarps;
constexpr int kBGroups = kBTeamThreads / kBT;   // channel groups of the dense part: thread = (column, group)
constexpr int kBLoaders = 3;               // loader warps (tile k of the CTA is loaded by warp k % 3)
constexpr int kBThreads = kBConsumers + 32 * (kBLoaders + 1);   // + the store warp
constexpr int kStageItems = 12;            // positives per tile whose records are staged in the slot (more: read from global)

constexpr unsigned kRoleValid = 1u, kRolePos = 2u, kRoleCls = 4u, kRoleCand = 8u;
constexpr unsigned kRoleSpecial = kRolePos | kRoleCls | kRoleCand;

// Header of a ring slot, behind the tile's logits.  `rec` is filled by asynchronous copies
// (cp.async) that complete on the slot's full barrier.
struct __align__(32) TileHeader {
  PosRec rec[kStageItems];          // records of the tile's first positives (index: rec_of[item])
  unsigned char role[kBT];
  unsigned char item_col[kBT];      // special columns of the tile, ascending
  unsigned char col_item[kBT];      // column -> its item index
  unsigned char rec_of[kBT];        // item -> index into rec[], 255: not staged
  float kd[kBT];                    // weighted KL of the items that are box candidates (consumers -> store warp)
  int n_items, cls_k, pad0, pad1;   // cls_k: K_cls of the tile's image (class-response normaliser)
};

struct StudentArgs {
  Ptr5 t_cls, t_box;
  MPtr5 g_cls, g_box;        // used by the non-TMA tiles
  Ptr5 s_cls, s_box;
  const int32_t* gt_inds;
  const uint8_t* sel_flags;
  const int32_t* cls_count;
  const float* avg;
  const float* upstream;
  const unsigned int* skip_flag;
  float dlw;
  int tiles_per_img, total_tiles, stages, stage_bytes;
  int dev;                           // TEMPORARY dev switches: 1 skip sparse, 2 skip dense, 4 skip stores
  unsigned long long* trace;         // TEMPORARY: per-tile timestamps of CTA 0 [tile][8]
  int lvl_tile_start[kLevels + 1];   // prefix of ceil(hw / kBT)
  int use_tma[kLevels];
};

struct __align__(64) StudentMaps {
  CUtensorMap s_cls[kLevels], s_box[kLevels], g_cls[kLevels], g_box[kLevels];   // [C x 64] / [68 x 64] tiles
  CUtensorMap t_cls[kLevels], t_box[kLevels];                                   // [ori x 4] / [68 x 4]: one anchor's teacher column
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

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define TRACE(kk, slotid) do { if (A.trace && blockIdx.x == 0 && lane == 0) A.trace[(size_t)(kk) * 16 + (slotid)] = gtime(); } while (0)

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
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, const void* src,
                                             unsigned long long policy) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;"
               ::"l"(map), "r"(c0), "r"(c1), "r"(smem_addr(src)), "l"(policy) : "memory");
}

// Box rows of one special column (a positive and / or an ERS box candidate), by one whole warp, in
// place in the tile.  lane = side * 8 + b holds bins b, b + 8 (b == 0: also 16) of its side.
__device__ __forceinline__ void item_box_rows(const Geo& g, const Workspace& ws, const StudentArgs& A, const BTile& b,
                                              const float* box_in, float* box_out, int pitch, const TileHeader* hd, float* hd_kd, const float* tbox_staged,
                                              const float* tbox_global, int HW, int it, int icol, int hwI, bool is_pos,
                                              bool is_cand, float w_kd, float avg2, float inv_T, size_t ga0, int lane) {
  const int side = lane >> 3, bb = lane & 7;
  const float* brow_in = box_in + (size_t)(side * kBins) * pitch + icol;
  float* brow = box_out + (size_t)(side * kBins) * pitch + icol;
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
      const float zt = !ok[i] ? -INFINITY : tbox_staged ? tbox_staged[ch * 4] : __ldg(tbox_global + (size_t)ch * HW);
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
    const float scale = upstream_of(A.upstream, acc_dbox(g, b.n)) * A.dlw * g.w_ld / 4.0f * (kT * kT / (float)kBins) / kT;
#pragma unroll
    for (int i = 0; i < 3; ++i) out[i] *= scale;
    if (lane == 0) hd_kd[it] = w_kd * (kl / (float)kBins * (kT * kT));   // .mean(1) * T*T; the store warp files it
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
      const float cb = upstream_of(A.upstream, acc_bbox(b.l)) * g.w_bbox / (1.0f + kEps32) / avg2 * rec.w * gd;
      const float cd = upstream_of(A.upstream, acc_dfl(b.l)) * g.w_dfl / 4.0f / avg2 * rec.w;
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
  const StudentMaps& maps;
  float* tcol;                   // this warp's teacher-column staging buffer
  unsigned long long* tbar;      // ... and the barrier its fetches complete on
  unsigned long long pol_keep;
  int tbox_off, lane, cwarp, twarp, col, q, oq, cq;   // cwarp: warp in the CTA, twarp: warp in its team
  float inv_avg1, avg2, inv_T;
};

// One tile, one consumer warp: the dense part of this warp's threads and the items of the tile that
// fall to this warp.  RING: the tile is in the ring slot `data` (logits in, gradients out, in place);
// else (level without a tensor map) it is read from / written to global memory directly.
template <bool RING>
__device__ __forceinline__ void consume_tile(const ConsumerCtx& cc, const BTile& b, int k, float* data, TileHeader* hd,
                                             float& qfl_part, float& dcls_part, uint32_t& tphase) {
  const Geo& g = cc.g;
  const Workspace& ws = cc.ws;
  const StudentArgs& A = cc.A;
  const StudentMaps& maps = cc.maps;
  float* tcol = cc.tcol;
  const int tbox_off = cc.tbox_off, lane = cc.lane, cwarp = cc.cwarp, twarp = cc.twarp, col = cc.col, q = cc.q, oq = cc.oq, cq = cc.cq;
  (void)cwarp;
  const float inv_avg1 = cc.inv_avg1, avg2 = cc.avg2, inv_T = cc.inv_T;
  const int C = g.C, ori = g.ori, cn = g.cn;
  constexpr int bq = (kBoxCh + kBGroups - 1) / kBGroups;
  const int HW = g.hw[b.l];
    // Start the asynchronous fetch of item `it`'s teacher column into this warp's staging buffer: two
    // TMA loads of [rows x 4] boxes -- the aligned group of four columns holding the anchor (the lines
    // were left in L2 by the teacher pass).  Returns true when a fetch is in flight or nothing needs fetching.
    auto fetch_teacher = [&](int item) {
      const int icol = hd->item_col[item];
      const unsigned irole = hd->role[icol];
      if (!RING || !(irole & (kRoleCls | kRoleCand))) return true;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of the buffer before the bulk write
      __syncwarp();
      if (lane == 0) {
        const int hwI = b.hw0 + icol;
        const uint32_t bytes = ((irole & kRoleCls) ? (uint32_t)ori * 16u : 0u) + ((irole & kRoleCand) ? (uint32_t)kBoxCh * 16u : 0u);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(cc.tbar)), "r"(bytes) : "memory");
        // (a box must start 16 B aligned in global memory: fetch the aligned group of four columns around the anchor)
        if (irole & kRoleCls) tma_load_2d(tcol, &maps.t_cls[b.l], hwI & ~3, b.n * ori, cc.tbar, cc.pol_keep);
        if (irole & kRoleCand) tma_load_2d(tcol + tbox_off, &maps.t_box[b.l], hwI & ~3, b.n * kBoxCh, cc.tbar, cc.pol_keep);
      }
      return true;
    };
    const size_t ga0 = (size_t)b.n * g.A + g.start[b.l] + b.hw0;
    // this warp's first item of the tile: request its teacher column now, use it after the dense part
    const int n_items = (A.dev & 1) ? 0 : hd->n_items;
    int it = (twarp - k / kBTeams) & (kBTeamWarps - 1);   // item i of the team's j-th tile goes to its warp (i + j) % 4
    bool fetched = false;
    if (it < n_items) fetched = fetch_teacher(it);
    // Where the tile lives: in the ring slot (logits in, gradients out, in place), or -- levels without a
    // tensor map -- straight in global memory.  Everything below addresses rows through (pointer, pitch).
    constexpr bool in_ring = RING;
    const size_t goff_c = (size_t)b.n * C * HW + b.hw0, goff_b = (size_t)b.n * kBoxCh * HW + b.hw0;
    const float* cls_in = in_ring ? data : A.s_cls.p[b.l] + goff_c;
    float* cls_out = in_ring ? data : A.g_cls.p[b.l] + goff_c;
    const float* box_in = in_ring ? data + (size_t)C * kBT : A.s_box.p[b.l] + goff_b;
    float* box_out = in_ring ? data + (size_t)C * kBT : A.g_box.p[b.l] + goff_b;
    const int pitch = RING ? kBT : HW;   // a compile-time stride in the ring
    // Dense part and sparse items touch disjoint addresses of the tile (the dense part skips the
    // rows a special column's item owns), so nothing orders them inside a tile.
    // ------------------------------------------------------------ dense part
    {
      const unsigned role = hd->role[col];
      const float lw = (role & kRoleValid) ? 1.0f : 0.0f;                // label_weights, gfl_head.py:650-655,663
      const float gs = lw * (upstream_of(A.upstream, acc_cls(b.l)) * g.w_cls * inv_avg1);
      int label = -1;
      float score = 0.f;
      if (role & kRolePos) {   // rare
        const int ri = hd->rec_of[hd->col_item[col]];
        const PosRec* rec = ri != 255 ? &hd->rec[ri] : ws.pos_rec + ga0 + col;
        label = rec->label;
        score = rec->score;
      }
      float loss = 0.f;
      const int c0 = q * cq, c1 = min(c0 + cq, cn);
      if (in_ring || col < b.cnt) {   // (columns past a level's end: zero-filled in the ring, absent in global memory)
        const float* nin = cls_in + (size_t)ori * pitch + col;
        float* nout = cls_out + (size_t)ori * pitch + col;
        const bool own_label = label >= c0 && label < c1;    // a positive's label channel lies in this thread's share
        const float x_label = own_label ? nin[(size_t)label * pitch] : 0.f;
#pragma unroll 4
        for (int c = c0; c < ((A.dev & 2) ? c0 : c1); ++c) {   // QFL, every element as a negative first: branch free (:260-261,317-320)
          const QflTerm tn = qfl_neg(nin[(size_t)c * pitch]);
          loss = fmaf(lw, tn.loss, loss);
          nout[(size_t)c * pitch] = gs * tn.grad;
        }
        if (own_label) {   // ... then the label channel of the (rare) positive is redone with its soft target
          const QflTerm tp = qfl_pos(x_label, score), tn = qfl_neg(x_label);
          loss += lw * (tp.loss - tn.loss);
          nout[(size_t)label * pitch] = gs * tp.grad;
        }
        qfl_part += loss;
        // structural zeros: the old-class rows of columns without a class-response row or a box
        // candidate's weight to read, the box rows of columns that are neither positive nor candidate
        if (!(role & (kRoleCls | kRoleCand))) {
          const int o0 = q * oq, o1 = min(o0 + oq, ori);
          for (int c = o0; c < o1; ++c) cls_out[(size_t)c * pitch + col] = 0.f;
        }
        if (!(role & (kRolePos | kRoleCand))) {
          float* brow = box_out + (size_t)(q * bq) * pitch + col;
#pragma unroll
          for (int j = 0; j < bq; ++j)
            if (q * bq + j < kBoxCh) brow[(size_t)j * pitch] = 0.f;
        }
      }
    }
    // ------------------------------------------------------------ sparse items: one WARP per special column
    // Item i of tile k goes to warp (i + k) % 16.  Lane layout for the box rows: lane = side * 8 + b,
    // holding bins b, b + 8 (and 16 for b == 0) of its side; reductions over a side are 8-lane shuffles.
    for (; it < n_items; it += kBTeamWarps) {
      if (!fetched) fetched = fetch_teacher(it);   // a second item of this warp in the same tile (rare)
      const int icol = hd->item_col[it];
      const unsigned irole = hd->role[icol];
      const bool is_cls = (irole & kRoleCls) != 0u, is_cand = (irole & kRoleCand) != 0u, is_pos = (irole & kRolePos) != 0u;
      const bool staged = in_ring;   // teacher column in this warp's staging buffer, else straight from global
      const int hwI = b.hw0 + icol;
      const float* tc_g = A.t_cls.p[b.l] + (size_t)b.n * ori * HW + hwI;
      const float* tb_g = A.t_box.p[b.l] + (size_t)b.n * kBoxCh * HW + hwI;
      if (staged && (is_cls || is_cand)) {
        mbar_wait_parity(cc.tbar, tphase);
        tphase ^= 1u;
      }
      fetched = false;
      // old-class rows of the column: lane owns channels lane, lane + 32, ...
      if (is_cls || is_cand) {
        const float kc = (float)hd->cls_k * (float)ori;
        const float scale_dc = upstream_of(A.upstream, acc_dcls(b.n)) * A.dlw * 2.0f / kc;
        float mx_old = -INFINITY;
        for (int c = lane; c < ori; c += 32) {
          const float xs = cls_in[(size_t)c * pitch + icol];
          mx_old = fmaxf(mx_old, xs);
          float gr = 0.f;
          if (is_cls) {   // class-response L2 (:181-186,324-332): g = 2 (x_s - x_t) / (K ori)
            const float xt = staged ? tcol[c * 4 + (hwI & 3)] : __ldg(tc_g + (size_t)c * HW);
            const float df = xs - xt;
            dcls_part = fmaf(df, df, dcls_part);
            gr = scale_dc * df;
          }
          cls_out[(size_t)c * pitch + icol] = gr;
        }
        if (is_pos || is_cand) {
          const float w_kd = sigmoid_ref(warp_max(mx_old));                                      // :217-218
          item_box_rows(g, ws, A, b, box_in, box_out, pitch, hd, hd->kd, staged ? tcol + tbox_off + (hwI & 3) : nullptr, tb_g, HW, it, icol, hwI, is_pos, is_cand,
                        w_kd, avg2, inv_T, ga0, lane);
        }
      } else {
        item_box_rows(g, ws, A, b, box_in, box_out, pitch, hd, hd->kd, nullptr, tb_g, HW, it, icol, hwI, is_pos, false, 0.f, avg2, inv_T,
                      ga0, lane);
      }
    }
}

// ----------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(kBThreads, 1)
student_pass_kernel(Geo g, Workspace ws, StudentArgs A, const __grid_constant__ StudentMaps maps) {
  if (A.skip_flag && *A.skip_flag == 0u) return;
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ __align__(8) unsigned long long s_full[8], s_done[8], s_empty[8];
  // loss sums of this CTA: [kLevels] QFL + [n_img] class-response squares, behind the ring (flushed once, at the end:
  // a global atomic in front of a tile's release-arrive would hold the slot for a full memory round trip)
  double* s_loss = reinterpret_cast<double*>(s_raw + (size_t)A.stages * A.stage_bytes);
  const int S = A.stages;
  const int C = g.C, ori = g.ori, cn = g.cn;
  const int rows = C + kBoxCh;
  // per consumer warp: a staging buffer for one anchor's teacher column ([ori + 68 rows][4 floats], filled by
  // two TMA loads of [rows x 4] boxes whose first column is the anchor) and the barrier those loads complete on
  const int tbox_off = ((ori * 16 + 127) & ~127) / 4;                  // floats: the box part starts 128 B aligned (TMA destination)
  const int tcol_stride = tbox_off + ((kBoxCh * 16 + 127) & ~127) / 4;   // floats per warp
  float* s_tcol = reinterpret_cast<float*>(s_raw + (size_t)A.stages * A.stage_bytes + (((size_t)(kLevels + g.n_img) * sizeof(double) + 127) & ~(size_t)127));
  __shared__ __align__(8) unsigned long long s_tbar[kBConsumers / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&s_full[s], 1 + 32);   // the loader's lane 0 (after arming the TMA bytes) + its 32 lanes' cp.async batches
      mbar_init(&s_done[s], kBTeamWarps);        // one arrival per warp of the consuming team
      mbar_init(&s_empty[s], 1);
    }
    for (int w = 0; w < kBConsumers / 32; ++w) mbar_init(&s_tbar[w], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < kLevels + g.n_img; i += kBThreads) s_loss[i] = 0.0;
  __syncthreads();
  auto slot_data = [&](int s) { return reinterpret_cast<float*>(s_raw + (size_t)s * A.stage_bytes); };
  auto slot_head = [&](int s) {
    return reinterpret_cast<TileHeader*>(s_raw + (size_t)s * A.stage_bytes + (size_t)rows * kBT * sizeof(float));
  };
  if (warp >= kBConsumers / 32 && warp < kBConsumers / 32 + kBLoaders) {
    // ================================================================== loader warps
    // Everything a loader does per tile is asynchronous (TMA, cp.async) except the tile's roles,
    // which it needs in registers to know what to fetch: those are requested one of ITS tiles ahead
    // (three of the CTA's tiles), so their latency is off the path.
    const int lw = warp - kBConsumers / 32;
    unsigned long long pol_stream;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
    int gi[2] = {-1, -1};
    unsigned fl[2] = {0u, 0u};
    int kcls = 0;
    auto fetch_roles = [&](int t) {
      if (t >= A.total_tiles) return;
      const BTile b = b_tile(g, A, t);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col = lane + 32 * h;
        const size_t ga = (size_t)b.n * g.A + g.start[b.l] + b.hw0 + col;
        gi[h] = col < b.cnt ? A.gt_inds[ga] : -1;
        fl[h] = col < b.cnt ? (unsigned)A.sel_flags[ga] : 0u;
      }
      kcls = A.cls_count[b.n];
    };
    fetch_roles(blockIdx.x + lw * gridDim.x);
    for (int k = lw; ; k += kBLoaders) {
      const int t = blockIdx.x + k * gridDim.x;
      if (t >= A.total_tiles) break;
      const int slot = k % S;
      const uint32_t ph = (uint32_t)(k / S) & 1u;
      const BTile b = b_tile(g, A, t);
      const int HW = g.hw[b.l];
      mbar_wait_parity(&s_empty[slot], ph ^ 1u);   // first pass over the ring: passes at once
      TRACE(k, 0);
      float* data = slot_data(slot);
      TileHeader* hd = slot_head(slot);
      if (A.use_tma[b.l]) {
        if (lane == 0) {
          mbar_expect_tx(&s_full[slot], (uint32_t)(rows * kBT * sizeof(float)));
          tma_load_2d(data, &maps.s_cls[b.l], b.hw0, b.n * C, &s_full[slot], pol_stream);
          tma_load_2d(data + (size_t)C * kBT, &maps.s_box[b.l], b.hw0, b.n * kBoxCh, &s_full[slot], pol_stream);
        }
      }
      // (tiles of levels without a tensor map -- rows not 16 B aligned -- never enter the ring: the
      // consumers read and write global memory directly; the slot only carries the header)
      TRACE(k, 8);
      // the roles requested one iteration ago
      const int cgi[2] = {gi[0], gi[1]};
      const unsigned cfl[2] = {fl[0], fl[1]};
      const int ckcls = kcls;
      unsigned role[2];
      int item[2];
      int n_pos = 0;
      int n_items = 0;
      if (cgi[0] == 12345678 && cfl[1] == 99u) n_items = 1;   // (touch the registers: wait for the loads)
      TRACE(k, 9);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col = lane + 32 * h;
        role[h] = (cgi[h] >= 0 ? kRoleValid : 0u) | (cgi[h] > 0 ? kRolePos : 0u) | ((cfl[h] & 1u) ? kRoleCls : 0u) |
                  ((cfl[h] & 2u) ? kRoleCand : 0u);
        const bool special = (role[h] & kRoleSpecial) != 0u;
        const unsigned m = __ballot_sync(0xffffffffu, special);
        item[h] = special ? n_items + __popc(m & ((1u << lane) - 1u)) : 255;
        n_items += __popc(m);
        hd->role[col] = (unsigned char)role[h];
        hd->col_item[col] = (unsigned char)item[h];
        if (special) hd->item_col[item[h]] = (unsigned char)col;
      }
      // (second sweep, all lanes converged: the positives' staging indices)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col = lane + 32 * h;
        const bool pos = (role[h] & kRolePos) != 0u;
        const unsigned mp = __ballot_sync(0xffffffffu, pos);
        const int ri = n_pos + __popc(mp & ((1u << lane) - 1u));
        n_pos += __popc(mp);
        if (pos) {
          hd->rec_of[item[h]] = (unsigned char)(ri < kStageItems ? ri : 255);
          if (ri < kStageItems) {
            const PosRec* src = ws.pos_rec + (size_t)b.n * g.A + g.start[b.l] + b.hw0 + col;
            cp_async_16(reinterpret_cast<char*>(&hd->rec[ri]), reinterpret_cast<const char*>(src));
            cp_async_16(reinterpret_cast<char*>(&hd->rec[ri]) + 16, reinterpret_cast<const char*>(src) + 16);
          }
        }
      }
      if (lane == 0) {
        hd->n_items = n_items;
        hd->cls_k = ckcls;
      }
      TRACE(k, 11);
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_full[slot]);   // releases the header written by all lanes
      cp_async_arrive(&s_full[slot]);   // every lane: its copies (possibly none) count towards the slot
      TRACE(k, 1);
      // Request the next tile's roles only now: a release-arrive waits for every load the thread has
      // in flight, so loads issued earlier would hold back this tile's full barrier by their latency.
      fetch_roles(blockIdx.x + (k + kBLoaders) * gridDim.x);
    }
    return;
  }

  if (warp == kBConsumers / 32 + kBLoaders) {
    // ================================================================== store warp
    unsigned long long pol_stream;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
    int k = 0;
    for (int t = blockIdx.x; t < A.total_tiles; t += gridDim.x, ++k) {
      const int slot = k % S;
      const uint32_t ph = (uint32_t)(k / S) & 1u;
      const BTile b = b_tile(g, A, t);
      mbar_wait_parity(&s_done[slot], ph);
      TRACE(k, 4);
      const float* data = slot_data(slot);
      {   // the candidates' weighted KL values of this tile -> ws.kd_loss (read by the take-back pass)
        const TileHeader* hd = slot_head(slot);
        const int n_items = hd->n_items;
        const size_t ga0 = (size_t)b.n * g.A + g.start[b.l] + b.hw0;
        for (int it = lane; it < n_items; it += 32) {
          const int icol = hd->item_col[it];
          if (hd->role[icol] & kRoleCand) ws.kd_loss[ga0 + icol] = hd->kd[it];
        }
      }
      if (A.use_tma[b.l]) {
        if (lane == 0 && !(A.dev & 4)) {
          tma_store_2d(&maps.g_cls[b.l], b.hw0, b.n * C, data, pol_stream);
          tma_store_2d(&maps.g_box[b.l], b.hw0, b.n * kBoxCh, data + (size_t)C * kBT, pol_stream);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the engine has read the slot
        }
      }
      __syncwarp();
      TRACE(k, 5);
      // (relaxed: the slot's reads are complete -- consumed by the stores / the bulk group; nothing to publish)
      if (lane == 0) asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&s_empty[slot])) : "memory");
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    return;
  }

  // ==================================================================== consumers
  // The consumer warps never synchronise with each other: inside a tile the dense part and the
  // items own disjoint addresses, and a warp that is done with its share of a tile moves on to the
  // next slot.  Loss sums travel in registers and are flushed (warp shuffle + one fp64 atomic) when
  // the level / image of the CTA's tile sequence changes.
  const int ctid = threadIdx.x;            // 0 .. kBConsumers - 1
  const int cwarp = ctid >> 5;
  const int team = cwarp / kBTeamWarps, twarp = cwarp % kBTeamWarps;
  const int ttid = ctid - team * kBTeamThreads;   // thread in the team
  const int col = ttid & (kBT - 1), q = ttid / kBT;
  const int oq = (ori + kBGroups - 1) / kBGroups, cq = (cn + kBGroups - 1) / kBGroups;
  constexpr int bq = (kBoxCh + kBGroups - 1) / kBGroups;
  const float inv_avg1 = 1.0f / (float)((double)A.avg[0] + (double)kEps32);   // losses/utils.py:60-61
  const float avg2 = fmaxf(A.avg[1], 1.0f);                                   // :407 clamp_(min=1)
  const float inv_T = 1.0f / g.T;
  unsigned long long pol_keep;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
  float* tcol = s_tcol + (size_t)cwarp * tcol_stride;   // this warp's teacher-column staging: [ori][4] | [68][4]
  uint32_t tphase = 0;
  const ConsumerCtx cc{g, ws, A, maps, tcol, &s_tbar[cwarp], pol_keep, tbox_off, lane, cwarp, twarp, col, q, oq, cq, inv_avg1, avg2, inv_T};
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
  for (int k = team;; k += kBTeams) {
    const int t = blockIdx.x + k * gridDim.x;
    if (t >= A.total_tiles) break;
    const int slot = k % S;
    const uint32_t ph = (uint32_t)(k / S) & 1u;
    const BTile b = b_tile(g, A, t);
    float* data = slot_data(slot);
    TileHeader* hd = slot_head(slot);
    if (b.n != cur_img || b.l != cur_lvl) {   // warp-uniform
      flush();
      cur_img = b.n;
      cur_lvl = b.l;
    }
    mbar_wait_parity(&s_full[slot], ph);
    if (twarp == 0) TRACE(k, 2);
    if (twarp == 3) TRACE(k, 6);
    if (A.use_tma[b.l]) consume_tile<true>(cc, b, k, data, hd, qfl_part, dcls_part, tphase);
    else consume_tile<false>(cc, b, k, data, hd, qfl_part, dcls_part, tphase);
    if (twarp == 0) TRACE(k, 13);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> the TMA store
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_done[slot]);
    if (twarp == 0) TRACE(k, 3);
    if (twarp == 3) TRACE(k, 7);
  }
  flush();
  asm volatile("bar.sync 1, %0;" ::"n"(kBConsumers) : "memory");   // consumers only: every warp has flushed
  for (int i = ctid; i < kLevels + g.n_img; i += kBConsumers)
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

static int g_dev_mask = -1, g_dev_stages = -1;
static unsigned long long* g_trace = nullptr;   // TEMPORARY tuning hooks (erd_student_dev)

static int env_int(const char* name, int dflt, int lo, int hi) {
  const char* e = getenv(name);
  if (!e) return dflt;
  const int v = atoi(e);
  return v < lo || v > hi ? dflt : v;
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
  const int rows = g.C + kBoxCh;
  A.stage_bytes = (int)(((size_t)rows * kBT * sizeof(float) + sizeof(TileHeader) + 127) & ~(size_t)127);
  // (the ring must be at least as deep as there are loader warps and consumer teams: nobody may get
  // more than one phase ahead of a slot's barriers)
  int want_stages = g_dev_stages >= kBTeams ? g_dev_stages : env_int("ERD_STUDENT_STAGES", 5, kBTeams, 8);
  const size_t tail_bytes = (((size_t)(kLevels + g.n_img) * sizeof(double) + 127) & ~(size_t)127) +
                            (size_t)(kBConsumers / 32) * (((g.ori * 16 + 127) & ~127) + ((kBoxCh * 16 + 127) & ~127));   // loss sums + teacher-column staging
  const int max_smem = 227 * 1024 - 1024 - (int)tail_bytes;
  int S = max_smem / A.stage_bytes;
  if (S > want_stages) S = want_stages;
  if (S > 8) S = 8;
  // A team only waits on the full barriers of its own tiles, so a slot must always be consumed by the
  // same team (S a multiple of the team count): a team that waited on a slot whose previous tile belongs
  // to another team could run a whole phase ahead of it, and a parity wait cannot tell phase j+1 from j-1.
  S -= S % kBTeams;
  if (S < kBTeams) return cudaErrorInvalidValue;   // num_classes too large for the smallest ring
  A.stages = S;
  A.trace = g_trace;
  A.dev = g_dev_mask >= 0 ? g_dev_mask : env_int("ERD_STUDENT_DEV", 0, 0, 255);
  // tensor maps, cached on the pointers they were built for (the training loop reuses its buffers)
  struct MapCache {
    const void* key[6 * kLevels];
    int hw[kLevels], n_img, C, ori;
    int use_tma[kLevels];
    StudentMaps maps;
    bool valid = false;
  };
  static thread_local MapCache cache;
  const void* key[6 * kLevels];
  for (int l = 0; l < kLevels; ++l) {
    key[l] = a.s_cls.p[l];
    key[kLevels + l] = a.s_box.p[l];
    key[2 * kLevels + l] = a.g_cls.p[l];
    key[3 * kLevels + l] = a.g_box.p[l];
    key[4 * kLevels + l] = a.t_cls.p[l];
    key[5 * kLevels + l] = a.t_box.p[l];
  }
  bool hit = cache.valid && cache.n_img == g.n_img && cache.C == g.C && cache.ori == g.ori;
  for (int i = 0; hit && i < 6 * kLevels; ++i) hit = cache.key[i] == key[i];
  for (int l = 0; hit && l < kLevels; ++l) hit = cache.hw[l] == g.hw[l];
  if (!hit) {
    static int allow_tma = env_int("ERD_STUDENT_TMA", 1, 0, 1);
    for (int l = 0; l < kLevels; ++l) {
      const long long rc = (long long)g.n_img * g.C, rb = (long long)g.n_img * kBoxCh;
      cache.use_tma[l] = allow_tma && g.vec[l] &&
                         tma_encode_rows(&cache.maps.s_cls[l], a.s_cls.p[l], g.hw[l], rc, g.C, kBT) &&
                         tma_encode_rows(&cache.maps.s_box[l], a.s_box.p[l], g.hw[l], rb, kBoxCh, kBT) &&
                         tma_encode_rows(&cache.maps.g_cls[l], a.g_cls.p[l], g.hw[l], rc, g.C, kBT) &&
                         tma_encode_rows(&cache.maps.g_box[l], a.g_box.p[l], g.hw[l], rb, kBoxCh, kBT) &&
                         tma_encode_rows(&cache.maps.t_cls[l], a.t_cls.p[l], g.hw[l], (long long)g.n_img * g.ori, g.ori, 4) &&
                         tma_encode_rows(&cache.maps.t_box[l], a.t_box.p[l], g.hw[l], rb, kBoxCh, 4);
      cache.hw[l] = g.hw[l];
    }
    for (int i = 0; i < 6 * kLevels; ++i) cache.key[i] = key[i];
    cache.n_img = g.n_img;
    cache.C = g.C;
    cache.ori = g.ori;
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
  const size_t smem = (size_t)S * A.stage_bytes + tail_bytes;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(student_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    smem_set = smem;
  }
  const int grid = A.total_tiles < sms ? A.total_tiles : sms;
  ERD_LAUNCH(kKStudent, st, (student_pass_kernel<<<grid, kBThreads, smem, st>>>(g, ws, A, cache.maps)));
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess)
    fprintf(stderr, "student_pass launch failed: %s grid=%d threads=%d smem=%zu S=%d stage=%d tail=%zu params=%zu\n", cudaGetErrorString(le), grid,
            kBThreads, smem, S, A.stage_bytes, tail_bytes, sizeof(Geo) + sizeof(Workspace) + sizeof(StudentArgs) + sizeof(StudentMaps));
  return le;
}

}  // namespace erd

extern "C" void erd_student_trace(unsigned long long* p) { erd::g_trace = p; }

extern "C" void erd_student_dev(int mask, int stages) {
  erd::g_dev_mask = mask;
  erd::g_dev_stages = stages;
}
