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

#include <cstdlib>

#include "loss_math.cuh"

namespace erd {

constexpr int kBT = 64;                    // anchors per tile
constexpr int kBConsumers = 512;           // 16 consumer warps
constexpr int kBGroups = kBConsumers / kBT;   // channel groups of the dense part: thread = (column, group)
constexpr int kBLoaders = 3;               // loader warps (tile k of the CTA is loaded by warp k % 3)
constexpr int kBThreads = kBConsumers + 32 * (kBLoaders + 1);   // + the store warp
constexpr int kStageItems = 12;            // special columns per tile whose teacher rows / records are staged in the slot

constexpr unsigned kRoleValid = 1u, kRolePos = 2u, kRoleCls = 4u, kRoleCand = 8u;
constexpr unsigned kRoleSpecial = kRolePos | kRoleCls | kRoleCand;

// Header of a ring slot, behind the tile's logits.  `rec` and the teacher rows behind the header
// are filled by asynchronous copies (cp.async) that complete on the slot's full barrier.
struct __align__(32) TileHeader {
  PosRec rec[kStageItems];          // records of the staged items that are positives
  unsigned char role[kBT];
  unsigned char item_col[kBT];      // special columns of the tile, ascending
  unsigned char col_item[kBT];      // column -> its item index
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
  int sub_start[kLevels];            // prefix of ceil(hw / 32): the teacher pass's tiles (two per student tile)
  int subs_per_img;
  int use_tma[kLevels];
};

struct __align__(64) StudentMaps {
  CUtensorMap s_cls[kLevels], s_box[kLevels], g_cls[kLevels], g_box[kLevels];
};

struct BTile {
  int n, l, hw0, cnt;
};

__device__ __forceinline__ BTile b_tile(const Geo& g, const StudentArgs& A, int t) {
  BTile b;
  b.n = t / A.tiles_per_img;
  const int r = t - b.n * A.tiles_per_img;
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
                                              float* data, const TileHeader* hd, float* hd_kd, const float* tbox_staged,
                                              const float* tbox_global, int HW, int it, int icol, int hwI, bool is_pos,
                                              bool is_cand, bool staged, float w_kd, float avg2, float inv_T, size_t ga0,
                                              int lane) {
  const int side = lane >> 3, bb = lane & 7;
  const int C = g.C;
  float* brow = data + (size_t)(C + side * kBins) * kBT + icol;
  int jb[3];
  bool ok[3];
  float zs[3], out[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    jb[i] = bb + 8 * i;
    ok[i] = jb[i] < kBins;
    zs[i] = ok[i] ? brow[jb[i] * kBT] : -INFINITY;
    out[i] = 0.f;
  }
  if (is_cand) {   // DFL-distribution KL at temperature T (:204-221, kd_loss.py:12-37), written as if the NMS kept it
    float a[3], t[3];
    float ms = -INFINITY, mt = -INFINITY;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int ch = side * kBins + jb[i];
      const float zt = !ok[i] ? -INFINITY : tbox_staged ? tbox_staged[ch] : __ldg(tbox_global + (size_t)ch * HW);
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
    const PosRec rec = staged ? hd->rec[it] : ws.pos_rec[ga0 + icol];
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
    if (ok[i]) brow[jb[i] * kBT] = out[i];
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
  const int ori_pad = (ori + 3) & ~3;
  const int trow_len = ori_pad + kBoxCh;   // staged teacher row of an item: [class logits, padded to 16 B | 68 box logits]
  // teacher rows come from the teacher pass's compact stash when it describes the current selection,
  // else (caller-provided index lists) straight from the teacher tensors
  const bool use_stash = *reinterpret_cast<volatile unsigned int*>(ws.stash_valid) != 0u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&s_full[s], 1 + 32);   // the loader's lane 0 (after arming the TMA bytes) + its 32 lanes' cp.async batches
      mbar_init(&s_done[s], kBConsumers / 32);   // one arrival per consumer warp
      mbar_init(&s_empty[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < kLevels + g.n_img; i += kBThreads) s_loss[i] = 0.0;
  __syncthreads();
  auto slot_data = [&](int s) { return reinterpret_cast<float*>(s_raw + (size_t)s * A.stage_bytes); };
  auto slot_head = [&](int s) {
    return reinterpret_cast<TileHeader*>(s_raw + (size_t)s * A.stage_bytes + (size_t)rows * kBT * sizeof(float));
  };
  auto slot_trow = [&](int s) {
    return reinterpret_cast<float*>(s_raw + (size_t)s * A.stage_bytes + (size_t)rows * kBT * sizeof(float) +
                                    sizeof(TileHeader));
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
    int2 sbase[2] = {make_int2(0, 0), make_int2(0, 0)};
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
      kcls = use_stash ? ws.stash_cnt[b.n * 2] : A.cls_count[b.n];
      if (use_stash) {
        const int sub = A.sub_start[b.l] + b.hw0 / 32;
        const int2* sp = ws.stash_base + (size_t)b.n * A.subs_per_img + sub;
        sbase[0] = sp[0];
        sbase[1] = b.cnt > 32 ? sp[1] : make_int2(0, 0);
      }
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
      float* trow = slot_trow(slot);
      if (A.use_tma[b.l]) {
        if (lane == 0) {
          mbar_expect_tx(&s_full[slot], (uint32_t)(rows * kBT * sizeof(float)));
          tma_load_2d(data, &maps.s_cls[b.l], b.hw0, b.n * C, &s_full[slot], pol_stream);
          tma_load_2d(data + (size_t)C * kBT, &maps.s_box[b.l], b.hw0, b.n * kBoxCh, &s_full[slot], pol_stream);
        }
      } else {
        // rows of this level are not 16 B aligned (H*W % 4 != 0): no tensor map; the tile comes in as
        // 4-byte asynchronous copies, a warp-wide 128 B request per half row, columns past the level's end zeroed
        const float* sc = A.s_cls.p[b.l] + (size_t)b.n * C * HW + b.hw0;
        const float* sb = A.s_box.p[b.l] + (size_t)b.n * kBoxCh * HW + b.hw0;
        const int c0 = lane, c1 = lane + 32;
        const bool in0 = c0 < b.cnt, in1 = c1 < b.cnt;
#pragma unroll 8
        for (int r = 0; r < rows; ++r) {
          const float* src = r < C ? sc + (size_t)r * HW : sb + (size_t)(r - C) * HW;
          float* dst = data + r * kBT;
          if (in0) cp_async_4(dst + c0, src + c0);
          else dst[c0] = 0.f;
          if (in1) cp_async_4(dst + c1, src + c1);
          else dst[c1] = 0.f;
        }
      }
      TRACE(k, 8);
      // the roles requested one iteration ago
      const int cgi[2] = {gi[0], gi[1]};
      const unsigned cfl[2] = {fl[0], fl[1]};
      const int ckcls = kcls;
      const int2 csb[2] = {sbase[0], sbase[1]};
      unsigned role[2];
      int item[2];
      int cslot[2], bslot[2];   // stash rows of this lane's columns (the stash orders a 32-anchor tile's rows by anchor)
      int n_items = 0;
      if (cgi[0] == 12345678 && cfl[1] == 99u) n_items = 1;   // (touch the registers: wait for the loads)
      TRACE(k, 9);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col = lane + 32 * h;
        role[h] = (cgi[h] >= 0 ? kRoleValid : 0u) | (cgi[h] > 0 ? kRolePos : 0u) | ((cfl[h] & 1u) ? kRoleCls : 0u) |
                  ((cfl[h] & 2u) ? kRoleCand : 0u);
        const bool special = (role[h] & kRoleSpecial) != 0u;
        {
          const unsigned below = (1u << lane) - 1u;
          cslot[h] = csb[h].x + __popc(__ballot_sync(0xffffffffu, (role[h] & kRoleCls) != 0u) & below);
          bslot[h] = csb[h].y + __popc(__ballot_sync(0xffffffffu, (role[h] & kRoleCand) != 0u) & below);
        }
        const unsigned m = __ballot_sync(0xffffffffu, special);
        item[h] = special ? n_items + __popc(m & ((1u << lane) - 1u)) : 255;
        n_items += __popc(m);
        hd->role[col] = (unsigned char)role[h];
        hd->col_item[col] = (unsigned char)item[h];
        if (special) hd->item_col[item[h]] = (unsigned char)col;
      }
      if (lane == 0) {
        hd->n_items = n_items;
        hd->cls_k = ckcls;
      }
      TRACE(k, 10);
      // stage the teacher rows and the positives' records of the first kStageItems items: the lanes
      // walk the items, each copying a strided share of the item's words
      const int staged = (A.dev & 8) ? 0 : min(n_items, kStageItems);   // dev 8: TIMING ONLY, teacher rows read as garbage
      for (int it = 0; it < staged; ++it) {
        // the item's column and role live in the lane that owns the column
        const int owner_h0 = __ffs(__ballot_sync(0xffffffffu, item[0] == it));
        const int owner_h1 = __ffs(__ballot_sync(0xffffffffu, item[1] == it));
        const int col = owner_h0 ? owner_h0 - 1 : owner_h1 - 1 + 32;
        const unsigned r = __shfl_sync(0xffffffffu, owner_h0 ? role[0] : role[1], (col & 31));
        const size_t hw = (size_t)b.hw0 + col;
        float* dst = trow + (size_t)it * trow_len;
        if (use_stash) {
          // contiguous rows: 16-byte copies, the class chunks first, then the 17 box chunks
          const int cs = __shfl_sync(0xffffffffu, owner_h0 ? cslot[0] : cslot[1], (col & 31));
          const int bs = __shfl_sync(0xffffffffu, owner_h0 ? bslot[0] : bslot[1], (col & 31));
          const int nc = (r & kRoleCls) ? ori_pad / 4 : 0, nb = (r & kRoleCand) ? kBoxCh / 4 : 0;
          const float* srcc = ws.stash_cls + ((size_t)b.n * g.sel_cap + cs) * ori_pad;
          const float* srcb = ws.stash_box + ((size_t)b.n * g.sel_cap + bs) * kBoxCh;
          for (int q = lane; q < nc + nb; q += 32) {
            if (q < nc) cp_async_16(dst + 4 * q, srcc + 4 * q);
            else cp_async_16(dst + ori_pad + 4 * (q - nc), srcb + 4 * (q - nc));
          }
        } else {
          if (r & kRoleCls) {
            const float* src = A.t_cls.p[b.l] + (size_t)b.n * ori * HW + hw;
            for (int c = lane; c < ori; c += 32) cp_async_4(dst + c, src + (size_t)c * HW);
          }
          if (r & kRoleCand) {
            const float* src = A.t_box.p[b.l] + (size_t)b.n * kBoxCh * HW + hw;
            for (int c = lane; c < kBoxCh; c += 32) cp_async_4(dst + ori_pad + c, src + (size_t)c * HW);
          }
        }
        if ((r & kRolePos) && lane < 2) {
          const PosRec* src = ws.pos_rec + (size_t)b.n * g.A + g.start[b.l] + hw;
          cp_async_16(reinterpret_cast<char*>(&hd->rec[it]) + 16 * lane, reinterpret_cast<const char*>(src) + 16 * lane);
        }
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
      } else {
        const int HW = g.hw[b.l];
        float* gc = A.g_cls.p[b.l] + (size_t)b.n * C * HW + b.hw0;
        float* gb2 = A.g_box.p[b.l] + (size_t)b.n * kBoxCh * HW + b.hw0;
        const int c0 = lane, c1 = lane + 32;
        const bool in0 = c0 < b.cnt, in1 = c1 < b.cnt;
        for (int r0 = 0; r0 < rows; r0 += 8) {   // eight rows of shared-memory reads in flight per lane
          float v0[8], v1[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = min(r0 + i, rows - 1);
            v0[i] = data[r * kBT + c0];
            v1[i] = data[r * kBT + c1];
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = r0 + i;
            if (r < rows) {
              float* dst = r < C ? gc + (size_t)r * HW : gb2 + (size_t)(r - C) * HW;
              if (in0) __stcs(dst + c0, v0[i]);
              if (in1) __stcs(dst + c1, v1[i]);
            }
          }
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
  const int col = ctid & (kBT - 1), q = ctid / kBT;
  const int cwarp = ctid >> 5;
  const int oq = (ori + kBGroups - 1) / kBGroups, cq = (cn + kBGroups - 1) / kBGroups;
  constexpr int bq = (kBoxCh + kBGroups - 1) / kBGroups;
  const float inv_avg1 = 1.0f / (float)((double)A.avg[0] + (double)kEps32);   // losses/utils.py:60-61
  const float avg2 = fmaxf(A.avg[1], 1.0f);                                   // :407 clamp_(min=1)
  const float inv_T = 1.0f / g.T;
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
  int k = 0;
  for (int t = blockIdx.x; t < A.total_tiles; t += gridDim.x, ++k) {
    const int slot = k % S;
    const uint32_t ph = (uint32_t)(k / S) & 1u;
    const BTile b = b_tile(g, A, t);
    const int HW = g.hw[b.l];
    float* data = slot_data(slot);
    TileHeader* hd = slot_head(slot);
    const float* trow = slot_trow(slot);
    if (b.n != cur_img || b.l != cur_lvl) {   // warp-uniform
      flush();
      cur_img = b.n;
      cur_lvl = b.l;
    }
    mbar_wait_parity(&s_full[slot], ph);
    if (cwarp == 0) TRACE(k, 2);
    if (cwarp == 15) TRACE(k, 6);
    const size_t ga0 = (size_t)b.n * g.A + g.start[b.l] + b.hw0;
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
        const int it = hd->col_item[col];
        const PosRec* rec = it < kStageItems ? &hd->rec[it] : ws.pos_rec + ga0 + col;
        label = rec->label;
        score = rec->score;
      }
      float loss = 0.f;
      const int c0 = q * cq, c1 = min(c0 + cq, cn);
      float* nrow = data + (size_t)ori * kBT + col;
      const bool own_label = label >= c0 && label < c1;    // a positive's label channel lies in this thread's share
      const float x_label = own_label ? nrow[label * kBT] : 0.f;
#pragma unroll 4
      for (int c = c0; c < ((A.dev & 2) ? c0 : c1); ++c) {   // QFL, every element as a negative first: branch free (:260-261,317-320)
        const QflTerm tn = qfl_neg(nrow[c * kBT]);
        loss = fmaf(lw, tn.loss, loss);
        nrow[c * kBT] = gs * tn.grad;
      }
      if (own_label) {   // ... then the label channel of the (rare) positive is redone with its soft target
        const QflTerm tp = qfl_pos(x_label, score), tn = qfl_neg(x_label);
        loss += lw * (tp.loss - tn.loss);
        nrow[label * kBT] = gs * tp.grad;
      }
      qfl_part += loss;
      // structural zeros: the old-class rows of columns without a class-response row or a box
      // candidate's weight to read, the box rows of columns that are neither positive nor candidate
      if (!(role & (kRoleCls | kRoleCand))) {
        const int o0 = q * oq, o1 = min(o0 + oq, ori);
        for (int c = o0; c < o1; ++c) data[c * kBT + col] = 0.f;
      }
      if (!(role & (kRolePos | kRoleCand))) {
        float* brow = data + (size_t)(C + q * bq) * kBT + col;
#pragma unroll
        for (int j = 0; j < bq; ++j)
          if (q * bq + j < kBoxCh) brow[j * kBT] = 0.f;
      }
    }
    if (cwarp == 0) TRACE(k, 12);
    // ------------------------------------------------------------ sparse items: one WARP per special column
    // Item i of tile k goes to warp (i + k) % 16.  Lane layout for the box rows: lane = side * 8 + b,
    // holding bins b, b + 8 (and 16 for b == 0) of its side; reductions over a side are 8-lane shuffles.
    const int n_items = (A.dev & 1) ? 0 : hd->n_items;
    for (int it = (cwarp - k) & (kBConsumers / 32 - 1); it < n_items; it += kBConsumers / 32) {
      const int icol = hd->item_col[it];
      const unsigned irole = hd->role[icol];
      const bool is_cls = (irole & kRoleCls) != 0u, is_cand = (irole & kRoleCand) != 0u, is_pos = (irole & kRolePos) != 0u && !(A.dev & 8);
      const bool staged = it < kStageItems;
      const int hwI = b.hw0 + icol;
      const float* tc_s = trow + (size_t)it * trow_len;                                        // staged teacher row
      const float* tc_g = A.t_cls.p[b.l] + (size_t)b.n * ori * HW + hwI;                        // or straight from global
      const float* tb_g = A.t_box.p[b.l] + (size_t)b.n * kBoxCh * HW + hwI;
      // old-class rows of the column: lane owns channels lane, lane + 32, ...
      if (is_cls || is_cand) {
        const float kc = (float)hd->cls_k * (float)ori;
        const float scale_dc = upstream_of(A.upstream, acc_dcls(b.n)) * A.dlw * 2.0f / kc;
        float mx_old = -INFINITY;
        for (int c = lane; c < ori; c += 32) {
          const float xs = data[c * kBT + icol];
          mx_old = fmaxf(mx_old, xs);
          float gr = 0.f;
          if (is_cls) {   // class-response L2 (:181-186,324-332): g = 2 (x_s - x_t) / (K ori)
            const float xt = staged ? tc_s[c] : __ldg(tc_g + (size_t)c * HW);
            const float df = xs - xt;
            dcls_part = fmaf(df, df, dcls_part);
            gr = scale_dc * df;
          }
          data[c * kBT + icol] = gr;
        }
        if (is_pos || is_cand) {
          const float w_kd = sigmoid_ref(warp_max(mx_old));                                      // :217-218
          item_box_rows(g, ws, A, b, data, hd, hd->kd, staged ? tc_s + ori_pad : nullptr, tb_g, HW, it, icol, hwI, is_pos, is_cand,
                        staged, w_kd, avg2, inv_T, ga0, lane);
        }
      } else {
        item_box_rows(g, ws, A, b, data, hd, hd->kd, nullptr, tb_g, HW, it, icol, hwI, is_pos, false, staged, 0.f, avg2, inv_T,
                      ga0, lane);
      }
    }
    if (cwarp == 0) TRACE(k, 13);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> the TMA store
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_done[slot]);
    if (cwarp == 0) TRACE(k, 3);
    if (cwarp == 15) TRACE(k, 7);
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
  {
    int subs = 0;
    for (int l = 0; l < kLevels; ++l) {
      A.sub_start[l] = subs;
      subs += (g.hw[l] + 31) / 32;
    }
    A.subs_per_img = subs;
  }
  A.total_tiles = tiles * g.n_img;
  const int rows = g.C + kBoxCh;
  A.stage_bytes = (int)(((size_t)rows * kBT * sizeof(float) + sizeof(TileHeader) +
                         (size_t)kStageItems * (((g.ori + 3) & ~3) + kBoxCh) * sizeof(float) + 127) & ~(size_t)127);
  // (the ring must be at least as deep as there are loader warps: a loader may only ever be one
  // phase ahead of a slot's empty barrier)
  int want_stages = g_dev_stages >= kBLoaders ? g_dev_stages : env_int("ERD_STUDENT_STAGES", 5, kBLoaders, 8);
  const int max_smem = 227 * 1024 - 2048 - (int)((kLevels + g.n_img) * sizeof(double));
  int S = max_smem / A.stage_bytes;
  if (S > want_stages) S = want_stages;
  if (S > 8) S = 8;
  if (S < kBLoaders) return cudaErrorInvalidValue;   // num_classes too large for the smallest ring
  A.stages = S;
  A.trace = g_trace;
  A.dev = g_dev_mask >= 0 ? g_dev_mask : env_int("ERD_STUDENT_DEV", 0, 0, 255);
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
    static int allow_tma = env_int("ERD_STUDENT_TMA", 1, 0, 1);
    for (int l = 0; l < kLevels; ++l) {
      const long long rc = (long long)g.n_img * g.C, rb = (long long)g.n_img * kBoxCh;
      cache.use_tma[l] = allow_tma && g.vec[l] &&
                         tma_encode_rows(&cache.maps.s_cls[l], a.s_cls.p[l], g.hw[l], rc, g.C, kBT) &&
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
  const size_t smem = (size_t)S * A.stage_bytes + (size_t)(kLevels + g.n_img) * sizeof(double);
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(student_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    smem_set = smem;
  }
  const int grid = A.total_tiles < sms ? A.total_tiles : sms;
  ERD_LAUNCH(kKStudent, st, (student_pass_kernel<<<grid, kBThreads, smem, st>>>(g, ws, A, cache.maps)));
  return cudaGetLastError();
}

}  // namespace erd

extern "C" void erd_student_trace(unsigned long long* p) { erd::g_trace = p; }

extern "C" void erd_student_dev(int mask, int stages) {
  erd::g_dev_mask = mask;
  erd::g_dev_stages = stages;
}
