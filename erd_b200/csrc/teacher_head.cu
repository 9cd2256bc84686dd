// Teacher head-output producer fusion: the teacher head's LAST 3x3 convolutions
// (gfl_cls: 256 -> ori classes, gfl_reg: 256 -> 68 box-distribution channels, + bias, box branch x Scale;
// mmdet/models/dense_heads/gfl_head.py:228-230, called under no_grad from
// mmdet/models/detectors/gfl_increment_erd.py:205) as ONE tcgen05 implicit GEMM whose epilogue is the teacher
// pass: the accumulators never become a tensor that is re-read -- the per-anchor cache (max sigmoid, first
// argmax, max box logit, softmax integrals), the fp64 threshold sums and the stash rows come straight out of
// tensor memory.  The 432 B/anchor read of teacher.cu disappears from the loss step (the logits can still be
// emitted, write-only, for callers that want them / for stash misses).
//
// GEMM view.  M = anchors, N = output channels (ori padded to 16, and 68 padded to 80), K = 9 taps x 256
// channels, TF32 operands (what cuDNN runs the fp32 reference convolution in by default), fp32 accumulation
// in TMEM.  A CTA works on a 16 x 16 pixel patch of one level of one image = two M=128 MMAs (left and right
// 8-pixel halves) per tower that share every weight tile.
//
// A operand without im2col, without re-loading per tap.  The patch WITH its halo (18 x 18 pixels, zero filled
// outside the map = the convolution's padding) is staged ONCE per 32-channel slice, in the no-swizzle K-major
// core-matrix layout [16-byte channel chunk][halo pixel][16 B]: a pixel's row is 16 B, 8 consecutive pixels
// of a halo row are one 8 x 16 B core matrix, the next 8-row group of the MMA's M is the next halo row
// (stride-byte-offset 18 * 16 B), the next 16 B of K is the next plane (leading-byte-offset 324 * 16 B).
// Each of the 9 taps is then the SAME buffer seen through a descriptor whose start address is shifted by
// (dy * 18 + dx) * 16 B -- 16-byte alignment is all the no-swizzle layout asks of a start address.  So every
// input byte crosses L2 -> shared memory once per patch (1.27x with the halo), not nine times.
// Input layout: NHWC fp32 (torch channels_last), so a pixel's 32 channels are one 128-byte line.
//
// Warp roles (320 threads, one CTA per SM, persistent over the patches):
//   warps 0-3  A producers: 16-byte loads -> 16-byte shared stores (zeros = padding) into the stage, two stages;
//   warp  4    B loader: one elected lane, 1-D bulk copies (cp.async.bulk) of the pre-packed weight tile of
//              (32-channel slice, tap) -- the packed image IS the shared-memory image -- up to four stages;
//   warp  5    TMEM allocation + the single MMA-issuing thread (tcgen05.mma kind::tf32, commit -> mbarriers);
//   warps 6-9  epilogue: tcgen05.ld of the own lane quadrant, one anchor per lane, exactly the arithmetic of
//              teacher.cu (identical bits for identical logits).  Accumulators live in a ring of per-half
//              buffers in TMEM (four for ori <= 48, three up to 80), handed back half by half, so the epilogue of
//              patch i runs under the MMAs of patch i + 1.
#include <cuda.h>

#include <cstdlib>

#include "erd_common.cuh"

namespace erd {

constexpr int kHC = 256;                          // tower channels (feat_channels of every gfl_increment config)
constexpr int kKC = 32;                           // channels per A stage
constexpr int kNKC = kHC / kKC;                   // 8 slices
constexpr int kPlanes = kKC / 4;                  // 16-byte chunks per pixel and slice
constexpr int kPatch = 16;
constexpr int kHalo = kPatch + 2;                 // 18
constexpr int kHaloPx = kHalo * kHalo;            // 324
constexpr int kAPlane = kHaloPx * 16;             // 5 184 B: one 16-byte chunk of every halo pixel (LBO of A)
constexpr int kATower = kAPlane * kPlanes;        // 41 472 B
constexpr int kAStageBytes = 2 * kATower;         // cls tower + reg tower
constexpr int kAStages = 2;
constexpr int kBStages = 4;                        // at most (as many as fit: 4 for ori <= 48, else 3)
constexpr int kRegPad = 80;                       // 68 box channels padded to a legal UMMA N
constexpr int kHeadThreads = 320;
constexpr int kCopiesPerTower = kHaloPx * kPlanes;                 // 2 592 16-byte copies
constexpr int kCopyIters = (kCopiesPerTower + 127) / 128;         // 21 per producer thread
constexpr int kCopyBatch = 7;                                      // loads in flight per thread and tower

struct HeadArgs {
  Ptr5 f_cls, f_reg;            // NHWC (N, H, W, 256)
  MPtr5 o_cls, o_box;           // optional NCHW logits (emit)
  const float* w_cls;           // packed by head_pack_kernel
  const float* w_reg;
  const float* b_cls;
  const float* b_reg;
  float scale[kLevels];
  int32_t* cls_count;
  int32_t* box_count;
  int ncls_pad;
  int tiles_x[kLevels];
  int lvl_tile_start[kLevels + 1];
  int tiles_per_img, total_tiles;
  int stash_pitch;
  int emit;
  int n_hbuf, h_stride;         // accumulator buffers in TMEM, one per patch HALF (a ring of 2..4), and their column stride
  int b_stage_bytes, b_stages;
  int exp;                      // developer experiments (-DERD_HEAD_EXP, env ERD_HEAD_EXP): results are garbage when set
};

struct HTile {
  int n, l, y0, x0, sub;
};

__device__ __forceinline__ HTile h_tile(const HeadArgs& A, int t) {
  HTile b;
  b.n = t / A.tiles_per_img;
  b.sub = t - b.n * A.tiles_per_img;
  b.l = 0;
#pragma unroll
  for (int i = 1; i < kLevels; ++i) b.l += (b.sub >= A.lvl_tile_start[i]) ? 1 : 0;
  const int r = b.sub - A.lvl_tile_start[b.l];
  const int ty = r / A.tiles_x[b.l];
  b.y0 = ty * kPatch;
  b.x0 = (r - ty * A.tiles_x[b.l]) * kPatch;
  return b;
}

__device__ __forceinline__ uint32_t h_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void h_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void h_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void h_commit(uint32_t bar) {   // arrives when every MMA issued so far has completed
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle, version 1 (Blackwell):
// start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | 1 << 46 (built as two 32-bit words by the issuer)
__device__ __forceinline__ void h_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
      ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void h_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float4 h_ldg16(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void h_sts16(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// one lane of a converged warp (the same one every time): what issues MMAs / commits.  The surrounding control flow
// stays warp-uniform, so descriptors live in uniform registers (a lane == 0 branch made the compiler waterfall
// five R2UR broadcasts in a loop around every MMA: 60-100 cycles per issue, 4x the MMA's own time).
__device__ __forceinline__ bool h_elect() {
  uint32_t pred;
  asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t h_desc2(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
// instruction descriptor: D fp32, A / B tf32, both K-major, M = 128, N
__host__ __device__ constexpr uint32_t h_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

__global__ void __launch_bounds__(kHeadThreads, 1) teacher_head_kernel(Geo g, Workspace ws, HeadArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long s_bar[2 * kAStages + 2 * kBStages + 6];
  __shared__ uint32_t s_tmem;
  unsigned char* sA = smem;
  unsigned char* sB = smem + kAStages * kAStageBytes;
  float* s_bias = reinterpret_cast<float*>(sB + A.b_stages * A.b_stage_bytes);   // [ncls_pad] class, [80] box
  int* s_stash_cnt = reinterpret_cast<int*>(s_bias + A.ncls_pad + kRegPad);      // [n_img]
  const uint32_t bar0 = h_smem(s_bar);
  auto full_a = [&](int s) { return bar0 + 8u * s; };
  auto empty_a = [&](int s) { return bar0 + 8u * (kAStages + s); };
  auto full_b = [&](int s) { return bar0 + 8u * (2 * kAStages + s); };
  auto empty_b = [&](int s) { return bar0 + 8u * (2 * kAStages + kBStages + s); };
  auto t_full = [&](int s) { return bar0 + 8u * (2 * kAStages + 2 * kBStages + s); };
  auto t_empty = [&](int s) { return bar0 + 8u * (2 * kAStages + 2 * kBStages + 2 + s); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ori = g.ori;
  const int ncp = A.ncls_pad;

  if (threadIdx.x == 0) {
    auto init = [](uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); };
    for (int s = 0; s < kAStages; ++s) { init(full_a(s), 128); init(empty_a(s), 1); }
    for (int s = 0; s < kBStages; ++s) { init(full_b(s), 1); init(empty_b(s), 1); }
    for (int s = 0; s < 2; ++s) init(t_full(s), 1);
    for (int s = 0; s < 4; ++s) init(t_empty(s), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < ncp + kRegPad; i += blockDim.x)
    s_bias[i] = i < ncp ? (i < ori ? A.b_cls[i] : 0.f) : (i - ncp < kBoxCh ? A.b_reg[i - ncp] : 0.f);
  for (int i = threadIdx.x; i < g.n_img; i += blockDim.x) s_stash_cnt[i] = 0;
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < g.n_img; i += blockDim.x) {
      A.cls_count[i] = 0;
      A.box_count[i] = 0;
    }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(h_smem(&s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;

  if (warp < 4) {
    // ------------------------------------------------------------------ A producers
    // Lane mapping: a warp instruction copies 4 pixels x 8 chunks (four full 128-byte lines of global memory); the 8
    // lanes of each quarter warp (one shared-memory phase of a 16-byte access) take chunks {2a, 2a+1} of 4 consecutive
    // pixels: plane stride 5 184 = 64 mod 128, so they land in 8 distinct 16-byte bank groups (j-major lanes were a 4-way
    // conflict: 26 wavefronts per LDGSTS).
    const int p4 = lane & 3;
    const int j = 2 * (lane >> 3) + ((lane >> 2) & 1);
    const int px0 = 4 * warp + p4;              // + 16 * i
    uint32_t it = 0;
    for (int t = blockIdx.x; t < A.total_tiles; t += gridDim.x) {
      const HTile b = h_tile(A, t);
      const int H = g.h[b.l], W = g.w[b.l];
      uint32_t off[kCopyIters];   // source of copy i in 16-byte units, ~0: outside the map (zero fill)
#pragma unroll
      for (int i = 0; i < kCopyIters; ++i) {
        const int px = px0 + 16 * i;
        const int hy = px / kHalo, hx = px - hy * kHalo;
        const int gy = b.y0 - 1 + hy, gx = b.x0 - 1 + hx;
        const bool ok = px < kHaloPx && gy >= 0 && gy < H && gx >= 0 && gx < W;
        off[i] = ok ? (uint32_t)(((b.n * H + gy) * W + gx) * (kHC / 4) + j) : 0xFFFFFFFFu;
      }
      const float4* fc = reinterpret_cast<const float4*>(A.f_cls.p[b.l]);
      const float4* fr = reinterpret_cast<const float4*>(A.f_reg.p[b.l]);
      for (int kc = 0; kc < kNKC; ++kc, ++it) {
        const int s = it % kAStages;
        h_wait(empty_a(s), ((it / kAStages) & 1u) ^ 1u);
        const uint32_t dst = h_smem(sA + (size_t)s * kAStageBytes) + j * kAPlane + px0 * 16;
        // global -> registers -> 16-byte shared stores, in three batches of 14 loads per thread (28 KB in flight per
        // SM).  (cp.async writes shared memory sector by sector as the data return: 21 wavefronts per warp
        // instruction instead of 4, a quarter of the shared-memory bandwidth the MMA operand fetch needs.)
#ifdef ERD_HEAD_EXP
        if (!(A.exp & 1))
#endif
#pragma unroll
        for (int b0 = 0; b0 < kCopyIters; b0 += kCopyBatch) {
          float4 vc[kCopyBatch], vr[kCopyBatch];
#pragma unroll
          for (int i = 0; i < kCopyBatch; ++i) {
            const bool ok = off[b0 + i] != 0xFFFFFFFFu;
            vc[i] = ok ? h_ldg16(fc + off[b0 + i] + kc * kPlanes) : make_float4(0.f, 0.f, 0.f, 0.f);
            vr[i] = ok ? h_ldg16(fr + off[b0 + i] + kc * kPlanes) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int i = 0; i < kCopyBatch; ++i) {
            if (px0 + 16 * (b0 + i) < kHaloPx) {
              h_sts16(dst + (b0 + i) * 256, vc[i]);
              h_sts16(dst + kATower + (b0 + i) * 256, vr[i]);
            }
          }
        }
        // make the stage visible to the tensor core's (async) proxy and hand it over
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        h_arrive(full_a(s));
      }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ B loader (warp-uniform loops, one lane issues)
    uint32_t ib = 0;
    const uint32_t cls_bytes = (uint32_t)ncp * 128u, reg_bytes = (uint32_t)kRegPad * 128u;
    for (int t = blockIdx.x; t < A.total_tiles; t += gridDim.x) {
      for (int kt = 0; kt < kNKC * 9; ++kt, ++ib) {   // (slice, tap) in the order the packed weights are stored
        const int s = ib % A.b_stages;
        h_wait(empty_b(s), ((ib / A.b_stages) & 1u) ^ 1u);
#ifdef ERD_HEAD_EXP
        if (A.exp & 2) {
          if (h_elect()) h_arrive(full_b(s));
          __syncwarp();
          continue;
        }
#endif
        if (h_elect()) {
          asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(full_b(s)), "r"(cls_bytes + reg_bytes) : "memory");
          const uint32_t dst = h_smem(sB + (size_t)s * A.b_stage_bytes);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst), "l"(A.w_cls + (size_t)kt * ncp * 32), "r"(cls_bytes), "r"(full_b(s)) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst + cls_bytes), "l"(A.w_reg + (size_t)kt * kRegPad * 32), "r"(reg_bytes), "r"(full_b(s)) : "memory");
        }
        __syncwarp();
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loops, one lane issues)
    const uint32_t idesc_cls = h_idesc(ncp), idesc_reg = h_idesc(kRegPad);
    const uint32_t lbo_cls = (uint32_t)ncp * 16u, lbo_reg = (uint32_t)kRegPad * 16u;
    // descriptor words: lo = start >> 4 | LBO >> 4 << 16, hi = SBO >> 4 | version 1 << 14
    uint32_t hi_a = (uint32_t)(kHalo * 16 >> 4) | (1u << 14);
    const uint32_t hi_b = (128u >> 4) | (1u << 14);
#ifdef ERD_HEAD_EXP
    if (A.exp & 128) hi_a = (256u >> 4) | (1u << 14);   // 8-pixel groups 256 B apart: every core matrix 128-byte aligned (with no tap shift)
#endif
    const uint32_t lo_a_lbo = (uint32_t)(kAPlane >> 4) << 16, lo_c_lbo = (lbo_cls >> 4) << 16, lo_r_lbo = (lbo_reg >> 4) << 16;
    uint32_t ia = 0, ib = 0, itile = 0;
    for (int t = blockIdx.x; t < A.total_tiles; t += gridDim.x, ++itile) {
      // the patch's two halves accumulate in two consecutive buffers of the half-buffer ring; the epilogue hands each
      // back as soon as it has read it, so with three buffers (49..80 old classes) the next patch still starts before
      // the epilogue of this one is through, and with four (up to 48) a whole patch ahead
      const uint32_t hb0 = (2 * itile) % A.n_hbuf, hb1 = (2 * itile + 1) % A.n_hbuf;
      const HTile b = h_tile(A, t);
      const bool right_half = b.x0 + 8 < g.w[b.l];   // a patch on the map's right edge may hold no pixel in its right half
      h_wait(t_empty(hb0), (((2 * itile) / A.n_hbuf) & 1u) ^ 1u);       // the epilogue has drained these buffers
      h_wait(t_empty(hb1), (((2 * itile + 1) / A.n_hbuf) & 1u) ^ 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t dbuf[2] = {tmem + hb0 * A.h_stride, tmem + hb1 * A.h_stride};
      const int fb = itile & 1;
      for (int kc = 0; kc < kNKC; ++kc, ++ia) {
        const int sa = ia % kAStages;
        h_wait(full_a(sa), (ia / kAStages) & 1u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint32_t a_stage = h_smem(sA + (size_t)sa * kAStageBytes);
        for (int tap = 0; tap < 9; ++tap, ++ib) {
          const int sb = ib % A.b_stages;
          h_wait(full_b(sb), (ib / A.b_stages) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const int dy = tap / 3, dx = tap - dy * 3;
          // the tap: the same buffer, shifted start (all in 16-byte units from here on)
#ifdef ERD_HEAD_EXP
          const uint32_t a0 = ((a_stage + ((A.exp & 128) ? 0 : (dy * kHalo + dx) * 16)) >> 4) | lo_a_lbo;
#else
          const uint32_t a0 = ((a_stage + (dy * kHalo + dx) * 16) >> 4) | lo_a_lbo;
#endif
          const uint32_t bc = h_smem(sB + (size_t)sb * A.b_stage_bytes);
          const uint32_t bc0 = (bc >> 4) | lo_c_lbo, br0 = ((bc + ncp * 128) >> 4) | lo_r_lbo;
#ifdef ERD_HEAD_EXP
          if ((A.exp & 64) && h_elect()) {   // no MMAs at all: the barrier protocol alone
          }
          if ((A.exp & 64) && h_elect()) {
            h_commit(empty_b(sb));
            if (tap == 8) h_commit(empty_a(sa));
            if (tap == 8 && kc == kNKC - 1) h_commit(t_full(fb));
          }
          if (!(A.exp & 64))
#endif
          if (h_elect()) {
#pragma unroll
            for (int s = 0; s < kKC / 8; ++s) {     // K = 8 per MMA: two 16-byte planes
#pragma unroll
              for (int h = 0; h < 2; ++h) {         // left / right 8-pixel half of the patch: M = 16 rows x 8 pixels
                if (h == 1 && !right_half) continue;
                const uint32_t accum = (kc | tap | s) ? 1u : 0u;
                const uint32_t dh = dbuf[h];
                const uint32_t a_lo = a0 + (2 * s * kAPlane + h * 128) / 16;
#ifdef ERD_HEAD_EXP
                if (!(A.exp & 8))
#endif
                h_mma(dh, h_desc2(a_lo, hi_a), h_desc2(bc0 + 2 * s * (lbo_cls >> 4), hi_b), idesc_cls, accum);
#ifdef ERD_HEAD_EXP
                if (!(A.exp & 4))
#endif
                h_mma(dh + ncp, h_desc2(a_lo + kATower / 16, hi_a), h_desc2(br0 + 2 * s * (lbo_reg >> 4), hi_b), idesc_reg, accum);
              }
            }
            h_commit(empty_b(sb));
            if (tap == 8) h_commit(empty_a(sa));
            if (tap == 8 && kc == kNKC - 1) h_commit(t_full(fb));
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue = the teacher pass
    const int q = warp & 3;   // TMEM lane quadrant this warp may read
    const unsigned int pc_bits = __ldg(ws.pthr_state), pb_bits = __ldg(ws.pthr_state + 1);
    const float pthr_c = pc_bits ? from_ordered_bits(~pc_bits) : INFINITY;
    const float pthr_b = pb_bits ? from_ordered_bits(~pb_bits) : INFINITY;
    uint32_t itile = 0;
    for (int t = blockIdx.x; t < A.total_tiles; t += gridDim.x, ++itile) {
      const HTile b = h_tile(A, t);
      const int H = g.h[b.l], W = g.w[b.l], HW = H * W;
      const float scale = A.scale[b.l];
      h_wait(t_full(itile & 1), (itile >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      double sums[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const uint32_t hb = (2 * itile + h) % A.n_hbuf;
        if (h == 1 && b.x0 + 8 >= W) {   // no pixel there: no MMA was issued for this half (warp-uniform)
          if (lane == 0) h_arrive(t_empty(hb));
          break;
        }
        const int mrow = 32 * q + lane;
        const int y = b.y0 + (mrow >> 3), x = b.x0 + 8 * h + (mrow & 7);
        const bool in = y < H && x < W;
        const int hw = y * W + x;
        const uint32_t taddr = tmem + hb * A.h_stride + ((uint32_t)(32 * q) << 16);
        float* oc = A.emit && in ? A.o_cls.p[b.l] + (size_t)b.n * ori * HW + hw : nullptr;
        float* ob = A.emit && in ? A.o_box.p[b.l] + (size_t)b.n * kBoxCh * HW + hw : nullptr;
        // class logits: first maximum (argmax semantics of torch.max, gfl_head_increment_erd.py:194-195)
        float best = 0.f;
        int arg = 0;
        for (int c0 = 0; c0 < ori; c0 += 8) {
          float v[8];
          h_ld8(taddr + c0, v);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = c0 + e;
            if (c < ori) {
              const float val = v[e] + s_bias[c];
              if (oc) oc[(size_t)c * HW] = val;
              if (c == 0) {
                best = val;
              } else if (val > best) {
                best = val;
                arg = c;
              }
            }
          }
        }
        // box logits: (conv + bias) * Scale (gfl_head.py:229), then the Integral (gfl_head_increment_erd.py:40-54)
        float z[72];
#pragma unroll
        for (int ch = 0; ch < 9; ++ch) h_ld8(taddr + ncp + 8 * ch, *reinterpret_cast<float(*)[8]>(&z[8 * ch]));
#pragma unroll
        for (int c = 0; c < kBoxCh; ++c) {
          z[c] = (z[c] + s_bias[ncp + c]) * scale;
          if (ob) ob[(size_t)c * HW] = z[c];
        }
        float u = -INFINITY;
        float dist[4];
        {   // the four integrals advanced together, divisions last: teacher.cu's scan, operation for operation
          const float kL2e = 1.4426950408889634f;
          float bias[4], sum[4], num[4];
#pragma unroll
          for (int sd = 0; sd < 4; ++sd) {
            const float* zs = z + sd * kBins;
            float mx = zs[0];
#pragma unroll
            for (int k = 1; k < kBins; ++k) mx = fmaxf(mx, zs[k]);
            bias[sd] = -mx * kL2e;
            sum[sd] = 0.f;
            num[sd] = 0.f;
            u = fmaxf(u, mx);
          }
#pragma unroll
          for (int k = 0; k < kBins; ++k) {
#pragma unroll
            for (int sd = 0; sd < 4; ++sd) {
              const float e = ex2_approx(fmaf(z[sd * kBins + k], kL2e, bias[sd]));
              sum[sd] += e;
              num[sd] = fmaf((float)k, e, num[sd]);
            }
          }
#pragma unroll
          for (int sd = 0; sd < 4; ++sd) dist[sd] = __fdiv_rn(num[sd], sum[sd]);
        }
        const float m = sigmoid_ref(best);
        const size_t ga = (size_t)b.n * g.A + g.start[b.l] + hw;
        // stash: the column of every anchor that clears the provisional thresholds (teacher.cu)
        {
          const bool want = in && (m > pthr_c || u > pthr_b);
          const unsigned wm = __ballot_sync(0xffffffffu, want);
          unsigned short myslot = 0;
          if (wm) {   // warp-uniform
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_stash_cnt[b.n], __popc(wm));
            base = __shfl_sync(0xffffffffu, base, 0);
            const int idx = base + __popc(wm & ((1u << lane) - 1u));
            const bool ok = want && idx < kStashPerCta;
            const int row = (int)blockIdx.x * kStashPerCta + idx;
            float* dst = ws.t_stash + ((size_t)b.n * kStashRows + (ok ? row : 0)) * A.stash_pitch;
            if (ok) myslot = (unsigned short)(row + 1);
            for (int c0 = 0; c0 < ori; c0 += 8) {   // the class part again from TMEM (the load is warp-collective)
              float v[8];
              h_ld8(taddr + c0, v);
              if (ok) {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  if (c0 + e < ori) dst[c0 + e] = v[e] + s_bias[c0 + e];
              }
            }
            if (ok) {
#pragma unroll
              for (int c = 0; c < kBoxCh; ++c) dst[ori + c] = z[c];
            }
          }
          if (in) ws.t_slot[ga] = myslot;
        }
        // every lane's TMEM reads of this half are complete (h_ld8 waits): hand its buffer back to the MMA thread now,
        // the rest of the half (cache, sums) works from registers
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) h_arrive(t_empty(hb));
        if (in) {
          sums[0] += (double)m;
          sums[1] += (double)m * (double)m;
          sums[2] += (double)u;
          sums[3] += (double)u * (double)u;
          ws.t_m[ga] = m;
          ws.t_arg[ga] = arg;
          ws.t_u[ga] = u;
          ws.t_dist[ga] = make_float4(dist[0], dist[1], dist[2], dist[3]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) sums[i] = warp_sum(sums[i]);
      if (lane < 4) {
        const double v = lane == 0 ? sums[0] : lane == 1 ? sums[1] : lane == 2 ? sums[2] : sums[3];
        __stcg(ws.ers_part + ((size_t)b.n * A.tiles_per_img * 4 + (size_t)b.sub * 4 + q) * 4 + lane, v);
      }
    }
  }
  // ---------------------------------------------------------------------- teardown
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// OIHW (O, 256, 3, 3) -> the shared-memory image of every (slice, tap) weight tile:
// [slice kc][tap][16-byte chunk j][output row n < Opad][4 floats], channel = kc * 32 + j * 4 + e; rows >= O are zero.
__global__ void head_pack_kernel(const float* __restrict__ w, int O, int Opad, float* __restrict__ out) {
  const int total = Opad * kHC * 9;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int e = idx & 3;
    int r = idx >> 2;
    const int n = r % Opad;
    r /= Opad;
    const int j = r & 7;
    r >>= 3;
    const int tap = r % 9, kc = r / 9;
    const int c = kc * kKC + j * 4 + e;
    out[idx] = n < O ? w[((size_t)n * kHC + c) * 9 + tap] : 0.f;
  }
}

// ----------------------------------------------------------------------------- host side
int head_ncls_pad(int ori) { return (ori + 15) & ~15; }

// 4 threshold partials per patch (one per epilogue warp); what launch_ers_flags is told to reduce
int head_partials_per_img(const Geo& g) {
  int tiles = 0;
  for (int l = 0; l < kLevels; ++l) tiles += ((g.h[l] + kPatch - 1) / kPatch) * ((g.w[l] + kPatch - 1) / kPatch);
  return tiles * 4;
}

cudaError_t launch_head_pack(const float* w, int O, float* out, cudaStream_t st) {
  const int Opad = head_ncls_pad(O);
  head_pack_kernel<<<256, 256, 0, st>>>(w, O, Opad, out);
  return cudaGetLastError();
}

cudaError_t launch_teacher_head(const Geo& g, const Workspace& ws, const Ptr5& f_cls, const Ptr5& f_reg, const float* w_cls,
                                const float* w_reg, const float* b_cls, const float* b_reg, const float* scale,
                                const MPtr5* o_cls, const MPtr5* o_box, int32_t* cls_count, int32_t* box_count,
                                cudaStream_t st) {
  HeadArgs A;
  A.f_cls = f_cls;
  A.f_reg = f_reg;
  A.emit = (o_cls && o_box) ? 1 : 0;
  A.exp = 0;
#ifdef ERD_HEAD_EXP
  if (const char* e = getenv("ERD_HEAD_EXP")) A.exp = atoi(e);
#endif
  for (int l = 0; l < kLevels; ++l) {
    A.o_cls.p[l] = A.emit ? o_cls->p[l] : nullptr;
    A.o_box.p[l] = A.emit ? o_box->p[l] : nullptr;
    A.scale[l] = scale[l];
  }
  A.w_cls = w_cls;
  A.w_reg = w_reg;
  A.b_cls = b_cls;
  A.b_reg = b_reg;
  A.cls_count = cls_count;
  A.box_count = box_count;
  A.ncls_pad = head_ncls_pad(g.ori);
  if (A.ncls_pad > 256) return cudaErrorInvalidValue;
  int tiles = 0;
  for (int l = 0; l < kLevels; ++l) {
    A.lvl_tile_start[l] = tiles;
    A.tiles_x[l] = (g.w[l] + kPatch - 1) / kPatch;
    tiles += ((g.h[l] + kPatch - 1) / kPatch) * A.tiles_x[l];
  }
  A.lvl_tile_start[kLevels] = tiles;
  A.tiles_per_img = tiles;
  A.total_tiles = tiles * g.n_img;
  A.stash_pitch = stash_pitch(g.ori);
  A.h_stride = A.ncls_pad + kRegPad;
  if (2 * A.h_stride > 512) return cudaErrorInvalidValue;
  A.n_hbuf = 512 / A.h_stride;   // 4: a whole patch ahead (ori <= 48), 3: half a patch ahead (ori <= 80), 2: none
  if (A.n_hbuf > 4) A.n_hbuf = 4;
  A.b_stage_bytes = (A.ncls_pad + kRegPad) * 128;
  const size_t fixed = (size_t)kAStages * kAStageBytes + (size_t)(A.ncls_pad + kRegPad) * 4 + (((size_t)g.n_img * 4 + 15) & ~(size_t)15);
  A.b_stages = kBStages;
  while (A.b_stages > 2 && fixed + (size_t)A.b_stages * A.b_stage_bytes > 227 * 1024 - 256) --A.b_stages;
  const size_t smem = fixed + (size_t)A.b_stages * A.b_stage_bytes;
  if (smem > 227 * 1024 - 256) return cudaErrorInvalidValue;
  static size_t smem_set = 0;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(teacher_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    smem_set = smem;
  }
  int grid = sms;
  if (grid > A.total_tiles) grid = A.total_tiles;
  if (grid > kStashCtas) grid = kStashCtas;   // the stash is laid out per CTA
  ERD_LAUNCH(kKTeacherHead, st, (teacher_head_kernel<<<grid, kHeadThreads, smem, st>>>(g, ws, A)));
  return cudaGetLastError();
}

}  // namespace erd
