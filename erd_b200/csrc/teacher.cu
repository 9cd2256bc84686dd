// The teacher pass: ONE persistent kernel over the teacher head outputs that does everything
// Elastic Response Selection needs except the ordered index lists:
//   scan     per anchor m = max_c sigmoid(t_cls) (sigmoid is monotone: sigmoid(max logit)), the first
//            argmax class, u = max_j raw box logit and the four softmax-integral distances -- the
//            per-anchor cache the NMS and the ordered lists read -- plus fp64 partial sums per tile;
//   thresholds  the warp that completes an image's last tile turns the partial sums into
//            thr = mean + 2 std (unbiased) with a fixed reduction tree and publishes them
//            (epoch-stamped flag, release / acquire);
//   extract  once an image's thresholds are known, every warp revisits its tiles of that image --
//            whose logits are still in L2 -- writes the per-anchor selection flags and copies the
//            selected anchors' teacher rows into a compact, row-major STASH, so the student pass
//            reads 160 + 272 contiguous bytes per selected anchor instead of 108 scattered sectors.
// Reference: GFLIncrementERD.sel_pos / sel_pos_single (mmdet/models/detectors/
// gfl_increment_erd.py:143-200); the Integral decode fused into the scan is
// gfl_head_increment_erd.py:40-54,189-195.
//
// Structure: one CTA per SM.  A loader warp streams tiles of 32 anchors -- [ori x 32] class logits
// and [68 x 32] box logits, two 2-D TMA loads (cp.async.bulk.tensor) -- into a ring of shared-memory
// slots; 16 consumer warps take the tiles round robin, one anchor per lane, and never synchronise
// with each other.  Levels whose rows are not 16-byte aligned come in as 4-byte cp.async copies.
#include <cuda.h>

#include <cstdlib>

#include "erd_common.cuh"

namespace erd {

constexpr int kAT = 32;                      // anchors per tile = one warp
constexpr int kAWarps = 16;                  // consumer warps
constexpr int kAThreads = 32 * (kAWarps + 1);   // + the loader warp
constexpr int kAMaxStages = 16;

struct TeacherArgs {
  Ptr5 t_cls, t_box;
  float* thr_out;        // (N, 2)
  uint8_t* sel_flags;    // (N, A)
  int tiles_per_img, total_tiles, stages, stage_bytes;
  int lvl_tile_start[kLevels + 1];   // prefix of ceil(hw / kAT)
  int use_tma[kLevels];
  int l2_keep;
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

// mean + 2 std (unbiased) of one image from its per-tile fp64 partial sums, by one warp: fixed
// lane -> tile mapping and a fixed shuffle tree, so the thresholds are bit-reproducible run to run.
__device__ __forceinline__ void image_thresholds(const Geo& g, const Workspace& ws, const TeacherArgs& A, int n, int lane) {
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const double* p = ws.ers_part + (size_t)n * A.tiles_per_img * 4;
  for (int t = lane; t < A.tiles_per_img; t += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] += __ldcg(p + (size_t)t * 4 + i);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = warp_sum(acc[i]);
  if (lane < 2) {
    const double s1 = acc[lane * 2], s2 = acc[lane * 2 + 1];
    const double An = (double)g.A;
    const double mean = s1 / An;
    double var = (s2 - s1 * s1 / An) / (An - 1.0);   // A == 1 -> NaN, as torch.std
    if (var < 0.0) var = 0.0;
    A.thr_out[n * 2 + lane] = __fadd_rn((float)mean, __fmul_rn(2.0f, (float)sqrt(var)));   // gfl_increment_erd.py:149,157
  }
}

__global__ void __launch_bounds__(kAThreads, 1)
teacher_pass_kernel(Geo g, Workspace ws, TeacherArgs A, const __grid_constant__ TeacherMaps maps) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ __align__(8) unsigned long long s_full[kAMaxStages], s_empty[kAMaxStages];
  const int S = A.stages;
  const int ori = g.ori;
  const int rows = ori + kBoxCh;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t_smem(&s_full[s])), "r"(1 + 32));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(t_smem(&s_empty[s])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(ws.teacher_epoch) + 1u;   // this launch's stamp

  if (warp == kAWarps) {
    // ================================================================== loader
    unsigned long long pol;
    if (A.l2_keep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    int k = 0;
    for (int t = blockIdx.x; t < A.total_tiles; t += gridDim.x, ++k) {
      const int slot = k % S;
      const uint32_t ph = (uint32_t)(k / S) & 1u;
      const ATile b = a_tile(g, A, t);
      const int HW = g.hw[b.l];
      t_wait(&s_empty[slot], ph ^ 1u);
      float* data = reinterpret_cast<float*>(s_raw + (size_t)slot * A.stage_bytes);
      if (A.use_tma[b.l]) {
        if (lane == 0) {
          asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(t_smem(&s_full[slot])),
                       "r"((uint32_t)(rows * kAT * sizeof(float))) : "memory");
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
              ::"r"(t_smem(data)), "l"(&maps.t_cls[b.l]), "r"(b.hw0), "r"(b.n * ori), "r"(t_smem(&s_full[slot])), "l"(pol) : "memory");
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
              ::"r"(t_smem(data + (size_t)ori * kAT)), "l"(&maps.t_box[b.l]), "r"(b.hw0), "r"(b.n * kBoxCh), "r"(t_smem(&s_full[slot])), "l"(pol) : "memory");
        }
      } else {
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
      if (lane == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(t_smem(&s_full[slot])) : "memory");
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(t_smem(&s_full[slot])) : "memory");
    }
    return;
  }

  // ==================================================================== consumer warps
  // Tile k of the CTA belongs to warp k % 16.  A warp scans its tiles in order and, between scans,
  // extracts those of its earlier tiles whose image thresholds have been published; it only ever
  // BLOCKS on a threshold after its last scan, so every scan -- and with it every threshold --
  // completes no matter how the warps interleave.
  const float4 kZero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  (void)kZero4;
  const int cap = g.sel_cap;
  const int ori_pad = (ori + 3) & ~3;
  int k_extract = warp;   // next tile of this warp to extract
  auto flag_ready = [&](int n) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ws.img_flag + n) : "memory");
    return v == epoch;
  };
  auto extract = [&](int kk) {
    const int t = blockIdx.x + kk * gridDim.x;
    const ATile b = a_tile(g, A, t);
    const int HW = g.hw[b.l];
    const size_t ga = (size_t)b.n * g.A + g.start[b.l] + b.hw0 + lane;
    const bool in = lane < b.cnt;
    const float thr_c = __ldcg(A.thr_out + b.n * 2), thr_b = __ldcg(A.thr_out + b.n * 2 + 1);
    const bool c = in && ws.t_m[ga] > thr_c, bx = in && ws.t_u[ga] > thr_b;   // strict (gfl_increment_erd.py:150,158)
    if (in) A.sel_flags[ga] = (uint8_t)((c ? 1 : 0) | (bx ? 2 : 0));
    const unsigned mc = __ballot_sync(0xffffffffu, c), mb = __ballot_sync(0xffffffffu, bx);
    int base_c = 0, base_b = 0;
    if (lane == 0) {
      if (mc) base_c = atomicAdd(ws.stash_cnt + b.n * 2, __popc(mc));
      if (mb) base_b = atomicAdd(ws.stash_cnt + b.n * 2 + 1, __popc(mb));
      ws.stash_base[(size_t)b.n * A.tiles_per_img + b.sub] = make_int2(base_c, base_b);
    }
    base_c = __shfl_sync(0xffffffffu, base_c, 0);
    base_b = __shfl_sync(0xffffffffu, base_b, 0);
    // the selected anchors' rows: the whole warp copies one row at a time, all gathers of the tile in flight
    const float* tc = A.t_cls.p[b.l] + (size_t)b.n * ori * HW + b.hw0;
    const float* tb = A.t_box.p[b.l] + (size_t)b.n * kBoxCh * HW + b.hw0;
    unsigned m = mc;
    int rank = 0;
    while (m) {
      const int j = __ffs(m) - 1;
      m &= m - 1u;
      float* dst = ws.stash_cls + ((size_t)b.n * cap + base_c + rank) * ori_pad;
      for (int ch = lane; ch < ori; ch += 32) dst[ch] = __ldcg(tc + (size_t)ch * HW + j);
      ++rank;
    }
    m = mb;
    rank = 0;
    while (m) {
      const int j = __ffs(m) - 1;
      m &= m - 1u;
      float* dst = ws.stash_box + ((size_t)b.n * cap + base_b + rank) * kBoxCh;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int ch = lane + 32 * i;
        if (ch < kBoxCh) dst[ch] = __ldcg(tb + (size_t)ch * HW + j);
      }
      ++rank;
    }
  };

  int k = warp;
  for (;; k += kAWarps) {
    const int t = blockIdx.x + k * gridDim.x;
    if (t >= A.total_tiles) break;
    const int slot = k % S;
    const uint32_t ph = (uint32_t)(k / S) & 1u;
    const ATile b = a_tile(g, A, t);
    const float* col = reinterpret_cast<const float*>(s_raw + (size_t)slot * A.stage_bytes) + lane;
    t_wait(&s_full[slot], ph);
    // ---- scan: one anchor per lane.  Lanes past the level's end hold zeros: computing on them
    // unconditionally keeps the loops free of predicates (their results are discarded).
    float best = col[0];
    int arg = 0;
    for (int c = 1; c < ori; ++c) {   // first maximum: argmax semantics of torch.max (gfl_head_increment_erd.py:194-195)
      const float v = col[c * kAT];
      if (v > best) { best = v; arg = c; }
    }
    float u = -INFINITY;
    float dist[4];
#pragma unroll
    for (int sd = 0; sd < 4; ++sd) {
      const float* scol = col + (size_t)(ori + sd * kBins) * kAT;
      float z[kBins];
#pragma unroll
      for (int j = 0; j < kBins; ++j) z[j] = scol[j * kAT];
      float mx = z[0];
#pragma unroll
      for (int j = 1; j < kBins; ++j) mx = fmaxf(mx, z[j]);
      const float kL2e = 1.4426950408889634f;
      const float bias = -mx * kL2e;
      float sum = 0.f, num = 0.f;
#pragma unroll
      for (int j = 0; j < kBins; ++j) {
        const float e = ex2_approx(fmaf(z[j], kL2e, bias));   // exp(z - mx), 2 ulp
        sum += e;
        num = fmaf((float)j, e, num);
      }
      dist[sd] = __fdiv_rn(num, sum);                           // Integral (:40-54)
      u = fmaxf(u, mx);
    }
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(t_smem(&s_empty[slot])) : "memory");
    const bool in = lane < b.cnt;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (in) {
      const float m = sigmoid_ref(best);
      const size_t ga = (size_t)b.n * g.A + g.start[b.l] + b.hw0 + lane;
      ws.t_m[ga] = m;
      ws.t_arg[ga] = arg;
      ws.t_u[ga] = u;
      ws.t_dist[ga] = make_float4(dist[0], dist[1], dist[2], dist[3]);
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
    // ---- the image's last tile computes its thresholds and opens the image for extraction
    __threadfence();
    __syncwarp();
    int last = 0;
    if (lane == 0) last = atomicAdd(ws.img_cnt + b.n, 1) == A.tiles_per_img - 1;
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
      __threadfence();
      image_thresholds(g, ws, A, b.n, lane);
      if (lane == 0) {
        ws.img_cnt[b.n] = 0;                 // clean for the next launch
        ws.stash_cnt[b.n * 2] = 0;
        ws.stash_cnt[b.n * 2 + 1] = 0;
      }
      __threadfence();
      __syncwarp();
      if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ws.img_flag + b.n), "r"(epoch) : "memory");
    }
    // ---- extract what is ready, without blocking
    while (k_extract <= k) {
      const int te = blockIdx.x + k_extract * gridDim.x;
      if (!flag_ready(te / A.tiles_per_img)) break;
      extract(k_extract);
      k_extract += kAWarps;
    }
  }
  // ---- drain: the remaining tiles of this warp, now waiting for their thresholds
  for (; k_extract < k; k_extract += kAWarps) {
    const int te = blockIdx.x + k_extract * gridDim.x;
    while (!flag_ready(te / A.tiles_per_img)) __nanosleep(200);
    extract(k_extract);
  }
  // ---- the last warp of the grid to finish stamps the epoch (and leaves the ticket clean)
  __threadfence();
  __syncwarp();
  if (lane == 0) {
    const unsigned int total = gridDim.x * kAWarps;
    if (atomicAdd(ws.teacher_done, 1u) == total - 1u) {
      *ws.teacher_done = 0u;
      *ws.teacher_epoch = epoch;
      ws.stash_valid[0] = 1u;
    }
  }
}

// ----------------------------------------------------------------------------- host side
bool tma_encode_rows(void* map, const void* base, int hw, long long rows_total, int box_rows, int box_cols);   // student.cu

static int t_env_int(const char* name, int dflt, int lo, int hi) {
  const char* e = getenv(name);
  if (!e) return dflt;
  const int v = atoi(e);
  return v < lo || v > hi ? dflt : v;
}

// returns the number of tiles per image (the ordered-list kernel does not need it any more; kept for symmetry)
cudaError_t launch_teacher_pass(const Geo& g, const Workspace& ws, const Ptr5& t_cls, const Ptr5& t_box, float* thr,
                                uint8_t* sel_flags, cudaStream_t st) {
  TeacherArgs A;
  A.t_cls = t_cls;
  A.t_box = t_box;
  A.thr_out = thr;
  A.sel_flags = sel_flags;
  int tiles = 0;
  for (int l = 0; l < kLevels; ++l) {
    A.lvl_tile_start[l] = tiles;
    tiles += (g.hw[l] + kAT - 1) / kAT;
  }
  A.lvl_tile_start[kLevels] = tiles;
  A.tiles_per_img = tiles;
  A.total_tiles = tiles * g.n_img;
  const int rows = g.ori + kBoxCh;
  A.stage_bytes = rows * kAT * (int)sizeof(float);   // a multiple of 128
  int S = (220 * 1024) / A.stage_bytes;
  static int want = t_env_int("ERD_TEACHER_STAGES", 12, 2, kAMaxStages);
  if (S > want) S = want;
  if (S > kAMaxStages) S = kAMaxStages;
  if (S < 2) return cudaErrorInvalidValue;
  A.stages = S;
  static int l2_keep = t_env_int("ERD_TEACHER_L2", 1, 0, 1);
  A.l2_keep = l2_keep;
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
  const size_t smem = (size_t)S * A.stage_bytes;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(teacher_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    smem_set = smem;
  }
  const int grid = A.total_tiles < sms ? A.total_tiles : sms;
  ERD_LAUNCH(kKErsScan, st, (teacher_pass_kernel<<<grid, kAThreads, smem, st>>>(g, ws, A, cache.maps)));
  return cudaGetLastError();
}

}  // namespace erd
