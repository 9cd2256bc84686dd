// The teacher pass: ONE persistent streaming kernel over the teacher head outputs.
//   scan        per anchor m = max_c sigmoid(t_cls) (sigmoid is monotone: sigmoid(max logit)), the
//               first argmax class, u = max_j raw box logit and the four softmax-integral distances
//               -- the per-anchor cache every later stage reads (ordered lists, NMS, the student
//               pass's roles) -- plus fp64 partial sums per tile;
//   (the thresholds thr = mean + 2 std come out of those sums in the short flags kernel that follows,
//   ers.cu: a reduction inside this kernel would put a fence and an atomic behind every tile).
// Reference: GFLIncrementERD.sel_pos / sel_pos_single (mmdet/models/detectors/
// gfl_increment_erd.py:143-200); the Integral decode fused into the scan is
// gfl_head_increment_erd.py:40-54,189-195.
//
// Structure: one CTA per SM, up to 16 warps, each its own pipeline with its own shared-memory slot
// for one tile of 32 anchors -- [ori x 32] class logits and [68 x 32] box logits, two 2-D TMA
// loads (cp.async.bulk.tensor), one anchor per lane: request, wait, scan, publish, request the
// next.  Nothing synchronises the warps.  The loads carry an L2 evict_first hint: the tensors are
// streamed once (the student pass reads the stash, not these lines).
// Levels whose rows are not 16-byte aligned come in as 4-byte cp.async copies.
//
// The stash.  The student pass needs the teacher's whole logit column of every ERS anchor (class-
// response L2, box-distribution KL), which in NCHW costs a 64-byte DRAM access per 4-byte element.
// The scan has that column in shared memory, but the thresholds that decide the selection only exist
// once the whole image has been scanned.  So the stash works with PROVISIONAL thresholds: the lowest
// mean + 1.7 std (the selection is mean + 2 std) any image of the PREVIOUS call had, which the flags kernel
// leaves behind -- the teacher is frozen and consecutive batches are drawn from the same data, so the
// estimate is good, and it costs nothing (a sampling pass over this call's images in front of the scan was
// measured: +12..15 us on the critical path for the same hit rate).  The scan copies the column of every
// anchor that clears them to a compact stash row.  The estimate only
// decides what is stashed, never what is selected: an ERS anchor that was not stashed is read from
// the tensors by the student pass as before, so results do not depend on the estimate.
#include <cuda.h>

#include <cstdlib>

#include "erd_common.cuh"

namespace erd {

constexpr int kAT = 32;                      // anchors per tile = one warp
constexpr int kAMaxWarps = 16;               // consumer warps == ring slots (each warp owns one slot)
constexpr int kAMaxThreads = 32 * kAMaxWarps;

struct TeacherArgs {
  Ptr5 t_cls, t_box;
  int32_t* cls_count;    // (N,) zeroed here for the flags kernel that follows
  int32_t* box_count;
  int tiles_per_img, total_tiles, stages, stage_bytes;
  int lvl_tile_start[kLevels + 1];   // prefix of ceil(hw / kAT)
  int use_tma[kLevels];
  int stash_pitch;                   // floats per stash row
};

struct __align__(64) TeacherMaps {
  CUtensorMap t_cls[kLevels], t_box[kLevels];
};

struct ATile {
  int n, l, hw0, cnt, sub;   // sub: index of the tile inside its image
};

__device__ __forceinline__ ATile a_tile(const Geo& g, const TeacherArgs& A, int t) {
  ATile b;
  b.n = t / A.tiles_per_img;
  b.sub = t - b.n * A.tiles_per_img;
  b.l = 0;
#pragma unroll
  for (int i = 1; i < kLevels; ++i) b.l += (b.sub >= A.lvl_tile_start[i]) ? 1 : 0;
  b.hw0 = (b.sub - A.lvl_tile_start[b.l]) * kAT;
  b.cnt = min(kAT, g.hw[b.l] - b.hw0);
  return b;
}

__device__ __forceinline__ uint32_t t_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(t_smem(bar)), "r"(parity) : "memory");
}

constexpr int kTeacherFreeSms = 24;     // SMs left to the kernels that run beside this pass

__global__ void __launch_bounds__(kAMaxThreads, 1)
teacher_pass_kernel(Geo g, Workspace ws, TeacherArgs A, const __grid_constant__ TeacherMaps maps) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ __align__(8) unsigned long long s_full[kAMaxWarps];
  int* s_stash_cnt = reinterpret_cast<int*>(s_raw + (size_t)A.stages * A.stage_bytes);   // [n_img] stash rows this CTA has used
  // Every warp is its own pipeline over the tiles k = warp, warp + W, ... of the CTA's sequence, with
  // its own shared-memory slot: request the tile, scan it, publish its sums, wait for the image's
  // thresholds, extract the selected rows from the slot.  Warps only meet at the per-image flag.
  // A warp never waits for a flag while it holds an unscanned tile, and its tiles come in image
  // order, so by induction over the images every flag is eventually published.
  const int W = A.stages;   // warps == slots
  const int ori = g.ori;
  const int rows = ori + kBoxCh;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < W; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t_smem(&s_full[s])), "r"(1 + 32));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < g.n_img; i += blockDim.x) s_stash_cnt[i] = 0;
  __syncthreads();
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < g.n_img; i += blockDim.x) {
      A.cls_count[i] = 0;
      A.box_count[i] = 0;
    }
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));   // streamed once: the student pass reads the stash, not these lines
  float* data = reinterpret_cast<float*>(s_raw + (size_t)warp * A.stage_bytes);
  unsigned long long* full = &s_full[warp];

  auto request_tile = [&](const ATile& b) {   // start loading a tile into this warp's slot
    const int HW = g.hw[b.l];
    if (A.use_tma[b.l]) {
      if (lane == 0) {
        asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(t_smem(full)),
                     "r"((uint32_t)(rows * kAT * sizeof(float))) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
            ::"r"(t_smem(data)), "l"(&maps.t_cls[b.l]), "r"(b.hw0), "r"(b.n * ori), "r"(t_smem(full)), "l"(pol) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
            ::"r"(t_smem(data + (size_t)ori * kAT)), "l"(&maps.t_box[b.l]), "r"(b.hw0), "r"(b.n * kBoxCh), "r"(t_smem(full)), "l"(pol) : "memory");
      }
    } else {
      // rows of this level are not 16 B aligned (H*W % 4 != 0): 4-byte asynchronous copies, a warp-wide
      // 128 B request per row; lanes past the level's end zero their column
      const float* sc = A.t_cls.p[b.l] + (size_t)b.n * ori * HW + b.hw0 + lane;
      const float* sb = A.t_box.p[b.l] + (size_t)b.n * kBoxCh * HW + b.hw0 + lane;
      const bool in = lane < b.cnt;
#pragma unroll 8
      for (int r = 0; r < rows; ++r) {
        const float* src = r < ori ? sc + (size_t)r * HW : sb + (size_t)(r - ori) * HW;
        if (in) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(t_smem(data + r * kAT + lane)), "l"(src) : "memory");
        else data[r * kAT + lane] = 0.f;
      }
    }
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(t_smem(full)) : "memory");
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(t_smem(full)) : "memory");
  };
  auto request = [&](int k) {   // tile k of the CTA's sequence
    const int t = blockIdx.x + k * gridDim.x;
    if (t < A.total_tiles) request_tile(a_tile(g, A, t));
  };

  uint32_t ph = 0;
  request(warp);
  // provisional thresholds left behind by the previous call's flags kernel (0: none yet -> nothing is stashed)
  const unsigned int pc_bits = __ldg(ws.pthr_state), pb_bits = __ldg(ws.pthr_state + 1);
  const float pthr_c = pc_bits ? from_ordered_bits(~pc_bits) : INFINITY;
  const float pthr_b = pb_bits ? from_ordered_bits(~pb_bits) : INFINITY;
  for (int k = warp;; k += W) {
    const int t = blockIdx.x + k * gridDim.x;
    if (t >= A.total_tiles) break;
    const ATile b = a_tile(g, A, t);
    const float* col = data + lane;
    t_wait(full, ph);
    ph ^= 1u;
    // (the copies of an unaligned level are complete here -- they arrived on the barrier; the wait is free and
    // tells tools that track cp.async only through wait_group, compute-sanitizer racecheck, as much)
    asm volatile("cp.async.wait_all;" ::: "memory");
    // ---- scan: one anchor per lane.  Lanes past the level's end hold zeros: computing on them
    // unconditionally keeps the loops free of predicates (their results are discarded).
    // The lane's instruction stream is what bounds this pass (4 warps per scheduler, every value a dependent chain),
    // so the chains are laid out for instruction-level parallelism -- without changing a single result bit:
    // the class argmax as four interleaved first-maximum chains (classes c = r mod 4) merged first-index-stable, the
    // four softmax integrals advanced together one bin at a time (each side's own operation order is unchanged), the
    // four IEEE divisions at the end (their slow-path branch used to fence one side's code off from the next).
    float cb[4];
    int ca[4];
    cb[0] = col[0];              // chain 0 starts from class 0 (a NaN there propagates as in torch.max's scan order)
    ca[0] = 0;
#pragma unroll
    for (int r = 1; r < 4; ++r) { cb[r] = -INFINITY; ca[r] = r; }
    {
      int c = 0;
      for (; c + 8 <= ori; c += 8) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = col[(c + e) * kAT];
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (v[e] > cb[e & 3]) { cb[e & 3] = v[e]; ca[e & 3] = c + e; }
      }
      for (; c < ori; c += 4) {   // ori % 8 classes left: chains stay (c mod 4)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (c + e < ori) {
            const float v = col[(c + e) * kAT];
            if (v > cb[e]) { cb[e] = v; ca[e] = c + e; }
          }
      }
    }
    // first maximum overall = the smallest index among the chains' maxima (argmax semantics of torch.max,
    // gfl_head_increment_erd.py:194-195)
    auto take = [](float& b0, int& a0, float b1, int a1) {
      if (b1 > b0 || (b1 == b0 && a1 < a0)) { b0 = b1; a0 = a1; }
    };
    take(cb[0], ca[0], cb[1], ca[1]);
    take(cb[2], ca[2], cb[3], ca[3]);
    take(cb[0], ca[0], cb[2], ca[2]);
    const float best = cb[0];
    const int arg = ca[0];
    float u = -INFINITY;
    float dist[4];
    {
      float z[4][kBins];
#pragma unroll
      for (int sd = 0; sd < 4; ++sd)
#pragma unroll
        for (int j = 0; j < kBins; ++j) z[sd][j] = col[(size_t)(ori + sd * kBins + j) * kAT];
      const float kL2e = 1.4426950408889634f;
      float bias[4], sum[4], num[4];
#pragma unroll
      for (int sd = 0; sd < 4; ++sd) {
        float mx = z[sd][0];
#pragma unroll
        for (int j = 1; j < kBins; ++j) mx = fmaxf(mx, z[sd][j]);
        bias[sd] = -mx * kL2e;
        sum[sd] = 0.f;
        num[sd] = 0.f;
        u = fmaxf(u, mx);
      }
#pragma unroll
      for (int j = 0; j < kBins; ++j) {
#pragma unroll
        for (int sd = 0; sd < 4; ++sd) {
          const float e = ex2_approx(fmaf(z[sd][j], kL2e, bias[sd]));   // exp(z - mx), 2 ulp
          sum[sd] += e;
          num[sd] = fmaf((float)j, e, num[sd]);
        }
      }
#pragma unroll
      for (int sd = 0; sd < 4; ++sd) dist[sd] = __fdiv_rn(num[sd], sum[sd]);   // Integral (:40-54)
    }
    const bool in = lane < b.cnt;
    const float m = sigmoid_ref(best);
    const size_t ga = (size_t)b.n * g.A + g.start[b.l] + b.hw0 + lane;
    // ---- stash: the columns that clear the provisional thresholds, while the tile is still in the slot
    {
      const bool want = in && (m > pthr_c || u > pthr_b);
      unsigned wm = __ballot_sync(0xffffffffu, want);
      unsigned short myslot = 0;
      if (wm) {   // warp-uniform
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_stash_cnt[b.n], __popc(wm));   // shared memory: this CTA's region of the image
        base = __shfl_sync(0xffffffffu, base, 0);
        const int idx = base + __popc(wm & ((1u << lane) - 1u));
        const bool ok = want && idx < kStashPerCta;
        const int row = (int)blockIdx.x * kStashPerCta + idx;
        if (ok) myslot = (unsigned short)(row + 1);
        unsigned om = __ballot_sync(0xffffffffu, ok);
        while (om) {
          const int c = __ffs(om) - 1;
          om &= om - 1u;
          const int r0 = __shfl_sync(0xffffffffu, row, c);
          float* dst = ws.t_stash + ((size_t)b.n * kStashRows + r0) * A.stash_pitch;
          for (int r = lane; r < rows; r += 32) dst[r] = data[r * kAT + c];
        }
      }
      if (in) ws.t_slot[ga] = myslot;
    }
    // every lane has read the slot: refill it now, so the next tile's latency overlaps the publishing below
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of the slot before the next bulk write
    __syncwarp();
    request(k + W);
    // the tile's fp64 sums (the flags kernel that follows reduces them to the image's thresholds)
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (in) {
      acc[0] = (double)m;
      acc[1] = (double)m * (double)m;
      acc[2] = (double)u;
      acc[3] = (double)u * (double)u;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = warp_sum(acc[i]);
    if (lane < 4) {
      const double v = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
      __stcg(ws.ers_part + ((size_t)b.n * A.tiles_per_img + b.sub) * 4 + lane, v);
    }
    if (in) {   // the per-anchor cache
      ws.t_m[ga] = m;
      ws.t_arg[ga] = arg;
      ws.t_u[ga] = u;
      ws.t_dist[ga] = make_float4(dist[0], dist[1], dist[2], dist[3]);
    }
  }
}

// ----------------------------------------------------------------------------- host side
bool tma_encode_rows(void* map, const void* base, int hw, long long rows_total, int box_rows, int box_cols);   // student.cu

cudaError_t launch_teacher_pass(const Geo& g, const Workspace& ws, const Ptr5& t_cls, const Ptr5& t_box,
                                int32_t* cls_count, int32_t* box_count, int* tiles_per_img, cudaStream_t st) {
  TeacherArgs A;
  A.t_cls = t_cls;
  A.t_box = t_box;
  A.cls_count = cls_count;
  A.box_count = box_count;
  int tiles = 0;
  for (int l = 0; l < kLevels; ++l) {
    A.lvl_tile_start[l] = tiles;
    tiles += (g.hw[l] + kAT - 1) / kAT;
  }
  A.lvl_tile_start[kLevels] = tiles;
  A.tiles_per_img = tiles;
  A.total_tiles = tiles * g.n_img;
  if (tiles_per_img) *tiles_per_img = tiles;
  const int rows = g.ori + kBoxCh;
  A.stage_bytes = rows * kAT * (int)sizeof(float);   // a multiple of 128
  const int tail_bytes = (g.n_img * (int)sizeof(int) + 7) & ~7;   // stash counters
  int S = (224 * 1024 - tail_bytes) / A.stage_bytes;   // consumer warps == ring slots
  if (S > kAMaxWarps) S = kAMaxWarps;
  if (S < 4) return cudaErrorInvalidValue;   // ori_classes too large for this tiling
  A.stages = S;
  A.stash_pitch = stash_pitch(g.ori);
  struct MapCache {
    const void* key[2 * kLevels];
    int hw[kLevels], n_img, ori;
    int use_tma[kLevels];
    TeacherMaps maps;
    bool valid = false;
  };
  static thread_local MapCache cache;
  bool hit = cache.valid && cache.n_img == g.n_img && cache.ori == g.ori;
  for (int l = 0; hit && l < kLevels; ++l)
    hit = cache.key[l] == t_cls.p[l] && cache.key[kLevels + l] == t_box.p[l] && cache.hw[l] == g.hw[l];
  if (!hit) {
    for (int l = 0; l < kLevels; ++l) {
      cache.use_tma[l] = g.vec[l] &&
                         tma_encode_rows(&cache.maps.t_cls[l], t_cls.p[l], g.hw[l], (long long)g.n_img * g.ori, g.ori, kAT) &&
                         tma_encode_rows(&cache.maps.t_box[l], t_box.p[l], g.hw[l], (long long)g.n_img * kBoxCh, kBoxCh, kAT);
      cache.key[l] = t_cls.p[l];
      cache.key[kLevels + l] = t_box.p[l];
      cache.hw[l] = g.hw[l];
    }
    cache.n_img = g.n_img;
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
  const size_t smem = (size_t)S * A.stage_bytes + (size_t)tail_bytes;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(teacher_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    smem_set = smem;
  }
  // (a CTA of this kernel fills its SM; the assignment chain that runs beside it gets a few SMs of its own)
  static const int free_sms = [] {
    const char* e = getenv("ERD_TEACHER_FREE_SMS");
    const int v = e ? atoi(e) : kTeacherFreeSms;
    return v < 0 ? 0 : v;
  }();
  int grid = sms - free_sms;
  if (grid < 1) grid = 1;
  if (grid > A.total_tiles) grid = A.total_tiles;
  if (grid > kStashCtas) grid = kStashCtas;   // the stash is laid out per CTA
  ERD_LAUNCH(kKErsScan, st, (teacher_pass_kernel<<<grid, 32 * S, smem, st>>>(g, ws, A, cache.maps)));
  return cudaGetLastError();
}

}  // namespace erd
