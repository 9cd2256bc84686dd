// Elastic Response Selection: one streaming pass over the teacher head outputs, then
// per-image thresholds and ordered stream compaction.
// Reference: GFLIncrementERD.sel_pos / sel_pos_single
// (mmdet/models/detectors/gfl_increment_erd.py:143-200); the Integral decode fused into the
// pass is gfl_head_increment_erd.py:40-54,189-195.
#include <cstdlib>
#include "erd_common.cuh"

namespace erd {

// ----------------------------------------------------------------------------- pass 1
// Per anchor: m = max_c sigmoid(t_cls) (sigmoid is monotone, so sigmoid(max logit)), the
// first argmax class, u = max_j raw box logit, and the four softmax-integral distances.
// Per CTA: sums of m, m^2, u, u^2 in fp64 (deterministic two-level reduction).
//
// Loads are decoupled from the arithmetic: a CTA owns T consecutive anchors of one (image,
// level) and requests its whole [ori + 68] x T logit tile up front with bulk asynchronous
// copies (cp.async.bulk, completion on mbarriers, one barrier per channel group); every
// thread then reduces its own anchor's column out of shared memory while the co-resident
// CTAs' copies are in flight.  This keeps > 200 KB of HBM requests outstanding per SM
// independent of register pressure (a register-staged version of this pass sat at 2.7 TB/s,
// latency-bound).  Bulk copies need 16 B aligned rows (hw % 4 == 0); tiles of the other levels
// (a few % of the anchors) are filled with ordinary loads.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}

template <int T>
__global__ void __launch_bounds__(T) ers_scan_kernel(Geo g, Workspace ws, Ptr5 t_cls, Ptr5 t_box) {
  extern __shared__ __align__(128) float s_tile[];   // [ori + 68][T]
  __shared__ __align__(8) unsigned long long s_bar[5];
  __shared__ double s_red[T / 32][4];
  const int n = blockIdx.y;
  int tile = blockIdx.x, l = 0;
#pragma unroll
  for (int i = 0; i < kLevels - 1; ++i) {
    const int tl = (g.hw[i] + T - 1) / T;
    if (l == i && tile >= tl) { tile -= tl; ++l; }
  }
  const int HW = g.hw[l];
  const int hw0 = tile * T;
  const int cnt = min(T, HW - hw0);
  const int ori = g.ori;
  const bool bulk = g.vec[l] != 0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int b = 0; b < 5; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[b])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (!bulk) {
    // unaligned level: every thread fetches its own column, 8 loads in flight
    if ((int)threadIdx.x < cnt) {
      const float* cb = t_cls.p[l] + (size_t)n * ori * HW + hw0 + threadIdx.x;
#pragma unroll 8
      for (int c = 0; c < ori; ++c) s_tile[(size_t)c * T + threadIdx.x] = __ldg(cb + (size_t)c * HW);
      const float* bb = t_box.p[l] + (size_t)n * kBoxCh * HW + hw0 + threadIdx.x;
#pragma unroll 8
      for (int c = 0; c < kBoxCh; ++c) s_tile[(size_t)(ori + c) * T + threadIdx.x] = __ldg(bb + (size_t)c * HW);
    }
  } else if (threadIdx.x < 32) {
    const uint32_t row_bytes = (uint32_t)cnt * 4u;
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_bar[0])), "r"(row_bytes * (uint32_t)ori) : "memory");
#pragma unroll
      for (int b = 1; b < 5; ++b)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_bar[b])), "r"(row_bytes * (uint32_t)kBins) : "memory");
    }
    __syncwarp();
    const float* cbase = t_cls.p[l] + (size_t)n * ori * HW + hw0;
    for (int c = threadIdx.x; c < ori; c += 32)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(s_tile + (size_t)c * T)), "l"(cbase + (size_t)c * HW), "r"(row_bytes), "r"(smem_u32(&s_bar[0])) : "memory");
    const float* bbase = t_box.p[l] + (size_t)n * kBoxCh * HW + hw0;
    for (int c = threadIdx.x; c < kBoxCh; c += 32)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(s_tile + (size_t)(ori + c) * T)), "l"(bbase + (size_t)c * HW), "r"(row_bytes),
                     "r"(smem_u32(&s_bar[1 + c / kBins])) : "memory");
  }
  const bool on = (int)threadIdx.x < cnt;
  const float* col = s_tile + threadIdx.x;
  // class logits: first maximum (argmax semantics of torch.max) and its sigmoid
  if (bulk) mbar_wait(&s_bar[0], 0);
  float best = -INFINITY;
  int arg = 0;
  if (on) {
#pragma unroll 8
    for (int c = 0; c < ori; ++c) {
      const float v = col[(size_t)c * T];
      if (v > best) { best = v; arg = c; }
    }
  }
  float u = -INFINITY;
  float dist[4];
#pragma unroll
  for (int sd = 0; sd < 4; ++sd) {
    if (bulk) mbar_wait(&s_bar[1 + sd], 0);
    const float* scol = col + (size_t)(ori + sd * kBins) * T;
    float z[kBins];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kBins; ++j) {
      z[j] = on ? scol[(size_t)j * T] : 0.f;
      mx = fmaxf(mx, z[j]);
    }
    float sum = 0.f, num = 0.f;
#pragma unroll
    for (int j = 0; j < kBins; ++j) {
      const float e = __expf(z[j] - mx);
      sum += e;
      num = fmaf((float)j, e, num);
    }
    dist[sd] = __fdiv_rn(num, sum);
    u = fmaxf(u, mx);
  }
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  if (on) {
    const float m = sigmoid_ref(best);
    const size_t a = (size_t)n * g.A + g.start[l] + hw0 + threadIdx.x;
    ws.t_m[a] = m;
    ws.t_arg[a] = arg;
    ws.t_u[a] = u;
    ws.t_dist[a] = make_float4(dist[0], dist[1], dist[2], dist[3]);
    acc[0] = (double)m;
    acc[1] = (double)m * (double)m;
    acc[2] = (double)u;
    acc[3] = (double)u * (double)u;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = warp_sum(acc[i]);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) s_red[threadIdx.x >> 5][i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double sm = 0.0;
    for (int w = 0; w < T / 32; ++w) sm += s_red[w][threadIdx.x];
    ws.ers_part[((size_t)n * gridDim.x + blockIdx.x) * 4 + threadIdx.x] = sm;
  }
}

// ----------------------------------------------------------------------------- pass 1, pipelined
// Persistent, warp-specialised form of the same pass (the default): two CTAs per SM; warp 0 is
// the producer (bulk copies of 17-row x 256-anchor stages into a ring of shared-memory slots,
// full/empty mbarriers per slot, L2 evict_last hint), warps 1..8 are consumers (one anchor per
// thread).  The producer runs ahead by the whole ring no matter what the consumers are doing.
// Ring depth is a trade: a deep ring (6 stages = 30 MB of requests in flight over the chip)
// floods the memory system's queues and makes every kernel running beside the scan 2-3x slower
// without making the scan faster; 3 stages measured best inside the full step (0.189 ms against
// 0.199 ms with 4 and 0.200 ms with 2, 256-anchor tiles), and 224-anchor tiles -- both CTAs' rings
// inside a 100 KB carve-out -- shave another 2 us.
constexpr int kPipeT = 224;                 // anchors per tile = consumer threads (7 warps)
constexpr int kPipeRows = kBins;            // rows per stage
constexpr int kPipeThreads = kPipeT + 32;
constexpr int kPipeStageFloats = kPipeRows * kPipeT;

struct PipeTile {
  int n, l, hw0, cnt, part;
};

__device__ __forceinline__ PipeTile pipe_tile(const Geo& g, int t, int tiles_per_img) {
  PipeTile p;
  p.n = t / tiles_per_img;
  int tile = t - p.n * tiles_per_img;
  p.part = tile;
  p.l = 0;
#pragma unroll
  for (int i = 0; i < kLevels - 1; ++i) {
    const int tl = (g.hw[i] + kPipeT - 1) / kPipeT;
    if (p.l == i && tile >= tl) { tile -= tl; ++p.l; }
  }
  p.hw0 = tile * kPipeT;
  p.cnt = min(kPipeT, g.hw[p.l] - p.hw0);
  return p;
}

__global__ void __launch_bounds__(kPipeThreads, 2) ers_scan_pipe_kernel(Geo g, Workspace ws, Ptr5 t_cls, Ptr5 t_box,
                                                                        int tiles_per_img, int total_tiles, int stages, int l2_keep) {
  extern __shared__ __align__(128) float s_ring[];   // [stages][kPipeRows][kPipeT]
  __shared__ __align__(8) unsigned long long s_full[16], s_empty[16];
  __shared__ double s_red[kPipeT / 32][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ori = g.ori;
  const int cls_chunks = (ori + kPipeRows - 1) / kPipeRows;
  if (threadIdx.x == 0) {
    for (int b = 0; b < stages; ++b) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_full[b])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_empty[b])), "r"(kPipeT / 32));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int slot = 0;
  uint32_t phase = 0;   // parity of the current pass over the ring
  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    unsigned long long keep_policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep_policy));
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const PipeTile p = pipe_tile(g, t, tiles_per_img);
      const int HW = g.hw[p.l];
      if (!g.vec[p.l]) continue;   // rows not 16 B aligned: the consumers load those tiles themselves
      const uint32_t row_bytes = (uint32_t)p.cnt * 4u;
      for (int ch = 0; ch < cls_chunks + 4; ++ch) {
        const bool is_cls = ch < cls_chunks;
        const int r0 = is_cls ? ch * kPipeRows : 0;
        const int rows = is_cls ? min(kPipeRows, ori - r0) : kPipeRows;
        const float* src = is_cls ? t_cls.p[p.l] + ((size_t)p.n * ori + r0) * HW + p.hw0
                                  : t_box.p[p.l] + ((size_t)p.n * kBoxCh + (ch - cls_chunks) * kBins) * HW + p.hw0;
        float* dst = s_ring + (size_t)slot * kPipeStageFloats;
        mbar_wait(&s_empty[slot], phase ^ 1u);   // slot free (first pass: passes immediately)
        if (lane == 0)
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_full[slot])),
                       "r"(row_bytes * (uint32_t)rows) : "memory");
        __syncwarp();
        if (lane < rows) {
          // the distillation kernels gather rows of these tensors later in the step: ask L2 to keep
          // them (evict_last) instead of letting the streaming traffic push them out
          if ((l2_keep == 2) || (l2_keep == 1 && !is_cls))
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                         ::"r"(smem_u32(dst + (size_t)lane * kPipeT)), "l"(src + (size_t)lane * HW), "r"(row_bytes),
                           "r"(smem_u32(&s_full[slot])), "l"(keep_policy) : "memory");
          else
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(dst + (size_t)lane * kPipeT)), "l"(src + (size_t)lane * HW), "r"(row_bytes),
                           "r"(smem_u32(&s_full[slot])) : "memory");
        }
        if (++slot == stages) { slot = 0; phase ^= 1u; }
      }
    }
    return;
  }
  // -------------------------------------------------------------------- consumers
  const int tid = threadIdx.x - 32;     // 0 .. kPipeT-1: anchor within the tile
  const int cwarp = tid >> 5;
  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    const PipeTile p = pipe_tile(g, t, tiles_per_img);
    const bool on = tid < p.cnt;
    float best = -INFINITY, u = -INFINITY;
    int arg = 0;
    float dist[4] = {0.f, 0.f, 0.f, 0.f};
    if (!g.vec[p.l]) {
      // unaligned level (a few % of the anchors): every consumer fetches its own column
      if (on) {
        const int HW = g.hw[p.l];
        const float* cb = t_cls.p[p.l] + (size_t)p.n * ori * HW + p.hw0 + tid;
#pragma unroll 8
        for (int c = 0; c < ori; ++c) {
          const float v = __ldg(cb + (size_t)c * HW);
          if (v > best) { best = v; arg = c; }
        }
        const float* bb = t_box.p[p.l] + (size_t)p.n * kBoxCh * HW + p.hw0 + tid;
#pragma unroll
        for (int sd = 0; sd < 4; ++sd) {
          float z[kBins];
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < kBins; ++j) {
            z[j] = __ldg(bb + (size_t)(sd * kBins + j) * HW);
            mx = fmaxf(mx, z[j]);
          }
          float sum = 0.f, num = 0.f;
#pragma unroll
          for (int j = 0; j < kBins; ++j) {
            const float e = __expf(z[j] - mx);
            sum += e;
            num = fmaf((float)j, e, num);
          }
          dist[sd] = __fdiv_rn(num, sum);
          u = fmaxf(u, mx);
        }
      }
    }
    for (int ch = 0; g.vec[p.l] && ch < cls_chunks + 4; ++ch) {
      const float* col = s_ring + (size_t)slot * kPipeStageFloats + tid;
      mbar_wait(&s_full[slot], phase);
      // rows of threads beyond the tile's last anchor hold stale but finite data: computing on
      // them unconditionally keeps the loops free of predicates (their results are discarded)
      if (ch < cls_chunks) {
        const int r0 = ch * kPipeRows;
        if (r0 + kPipeRows <= ori) {
#pragma unroll
          for (int r = 0; r < kPipeRows; ++r) {
            const float v = col[r * kPipeT];
            if (v > best) { best = v; arg = r0 + r; }
          }
        } else {
          for (int r = 0; r < ori - r0; ++r) {
            const float v = col[r * kPipeT];
            if (v > best) { best = v; arg = r0 + r; }
          }
        }
      } else {
        float z[kBins];
#pragma unroll
        for (int j = 0; j < kBins; ++j) z[j] = col[j * kPipeT];
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
        dist[ch - cls_chunks] = __fdiv_rn(num, sum);
        u = fmaxf(u, mx);
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_empty[slot])) : "memory");
      if (++slot == stages) { slot = 0; phase ^= 1u; }
    }
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (on) {
      const float m = sigmoid_ref(best);
      const size_t a = (size_t)p.n * g.A + g.start[p.l] + p.hw0 + tid;
      ws.t_m[a] = m;
      ws.t_arg[a] = arg;
      ws.t_u[a] = u;
      ws.t_dist[a] = make_float4(dist[0], dist[1], dist[2], dist[3]);
      acc[0] = (double)m;
      acc[1] = (double)m * (double)m;
      acc[2] = (double)u;
      acc[3] = (double)u * (double)u;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = warp_sum(acc[i]);
    asm volatile("bar.sync 1, %0;" ::"n"(kPipeT));   // consumers only: s_red free again
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) s_red[cwarp][i] = acc[i];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kPipeT));
    if (tid < 4) {
      double sm = 0.0;
      for (int w = 0; w < kPipeT / 32; ++w) sm += s_red[w][tid];
      ws.ers_part[((size_t)p.n * tiles_per_img + p.part) * 4 + tid] = sm;
    }
  }
}

static int launch_scan_pipe(const Geo& g, const Workspace& ws, const Ptr5& t_cls, const Ptr5& t_box, cudaStream_t st) {
  int tiles = 0;
  for (int l = 0; l < kLevels; ++l) tiles += (g.hw[l] + kPipeT - 1) / kPipeT;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(ers_scan_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * kPipeStageFloats * 4);
  }
  static int stages = 0;   // 2 CTAs/SM x stages x 17 KB of copies in flight per SM
  if (!stages) {
    const char* e = getenv("ERD_SCAN_STAGES");
    stages = e ? atoi(e) : 3;
    if (stages < 2 || stages > 6) stages = 3;
  }
  const int total = tiles * g.n_img;
  static int l2_keep = -1;
  if (l2_keep < 0) {
    const char* e = getenv("ERD_SCAN_L2");
    l2_keep = e ? atoi(e) : 2;   // 0 = no hint, 1 = box rows only, 2 = class and box rows
  }
  static int per_sm = 0;
  if (!per_sm) {
    const char* e = getenv("ERD_SCAN_CTAS");
    per_sm = e ? atoi(e) : 2;
    if (per_sm < 1 || per_sm > 4) per_sm = 2;
  }
  const int grid = total < per_sm * sms ? total : per_sm * sms;
  ERD_LAUNCH(kKErsScan, st,
             (ers_scan_pipe_kernel<<<grid, kPipeThreads, (size_t)stages * kPipeStageFloats * 4, st>>>(
                 g, ws, t_cls, t_box, tiles, total, stages, l2_keep)));
  return tiles;
}

template <int T>
static int launch_scan(const Geo& g, const Workspace& ws, const Ptr5& t_cls, const Ptr5& t_box, cudaStream_t st) {
  int tiles = 0;
  for (int l = 0; l < kLevels; ++l) tiles += (g.hw[l] + T - 1) / T;
  const size_t smem = (size_t)(g.ori + kBoxCh) * T * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(ers_scan_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  ERD_LAUNCH(kKErsScan, st, (ers_scan_kernel<T><<<dim3(tiles, g.n_img), T, smem, st>>>(g, ws, t_cls, t_box)));
  return tiles;
}

// ----------------------------------------------------------------------------- pass 2
// thr = mean + 2 * std (unbiased); rows strictly above it, in ascending anchor order (what
// nonzero() yields, gfl_increment_erd.py:150-151,158-159).  Each CTA owns a 2048-anchor chunk
// of one image; it recounts the flags of the anchors before its chunk from the L2-resident
// cache instead of waiting on a cross-CTA prefix, so one launch suffices.
constexpr int kSelThreads = 1024;
constexpr int kSelPer = 2;
constexpr int kSelChunk = kSelThreads * kSelPer;

__global__ void __launch_bounds__(kSelThreads) ers_select_kernel(Geo g, Workspace ws, int tiles, int32_t* cls_inds,
                                                                 int32_t* cls_count, int32_t* box_inds,
                                                                 int32_t* box_count, float* thr_out,
                                                                 uint8_t* __restrict__ sel_flags) {
  const int n = blockIdx.y;
  const int chunk0 = blockIdx.x * kSelChunk;
  __shared__ float s_thr[2];
  __shared__ int s_warp[2][kSelThreads / 32];
  __shared__ int s_base[2];
  {
    // image statistics from the per-CTA partial sums of pass 1: fixed thread -> partial mapping
    // and a fixed reduction tree, so the thresholds are bit-reproducible run to run
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const double* p = ws.ers_part + (size_t)n * tiles * 4;
    for (int t = threadIdx.x; t < tiles; t += kSelThreads) {
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] += p[t * 4 + i];
    }
    __shared__ double s_red[kSelThreads / 32][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = warp_sum(acc[i]);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) s_red[threadIdx.x >> 5][i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x < 2) {
      double s1 = 0.0, s2 = 0.0;
      for (int w = 0; w < kSelThreads / 32; ++w) { s1 += s_red[w][threadIdx.x * 2]; s2 += s_red[w][threadIdx.x * 2 + 1]; }
      const double A = (double)g.A;
      const double mean = s1 / A;
      double var = (s2 - s1 * s1 / A) / (A - 1.0);   // A == 1 -> NaN, as torch.std
      if (var < 0.0) var = 0.0;
      const float t = __fadd_rn((float)mean, __fmul_rn(2.0f, (float)sqrt(var)));
      s_thr[threadIdx.x] = t;
      if (blockIdx.x == 0) thr_out[n * 2 + threadIdx.x] = t;
    }
  }
  __syncthreads();
  const float thr_c = s_thr[0], thr_b = s_thr[1];
  const float* m = ws.t_m + (size_t)n * g.A;
  const float* u = ws.t_u + (size_t)n * g.A;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // selected rows before this chunk
  int pc = 0, pb = 0;
  if ((((uintptr_t)m | (uintptr_t)u) & 15) == 0) {   // chunk0 is a multiple of 2048: 16 B loads, 4 in flight
    const float4* m4 = reinterpret_cast<const float4*>(m);
    const float4* u4 = reinterpret_cast<const float4*>(u);
#pragma unroll 4
    for (int i = threadIdx.x; i < chunk0 / 4; i += kSelThreads) {
      const float4 a = m4[i], b = u4[i];
      pc += (a.x > thr_c) + (a.y > thr_c) + (a.z > thr_c) + (a.w > thr_c);
      pb += (b.x > thr_b) + (b.y > thr_b) + (b.z > thr_b) + (b.w > thr_b);
    }
  } else {
#pragma unroll 8
    for (int a = threadIdx.x; a < chunk0; a += kSelThreads) {
      pc += m[a] > thr_c;
      pb += u[a] > thr_b;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    pc += __shfl_xor_sync(0xffffffffu, pc, o);
    pb += __shfl_xor_sync(0xffffffffu, pb, o);
  }
  if (lane == 0) { s_warp[0][warp] = pc; s_warp[1][warp] = pb; }
  __syncthreads();
  if (threadIdx.x < 2) {
    int t = 0;
    for (int w = 0; w < kSelThreads / 32; ++w) t += s_warp[threadIdx.x][w];
    s_base[threadIdx.x] = t;
  }
  __syncthreads();
  // this chunk: 8 consecutive anchors per thread, block-wide exclusive scan of the counts
  const int a0 = chunk0 + threadIdx.x * kSelPer;
  unsigned fc = 0, fb = 0;
#pragma unroll
  for (int k = 0; k < kSelPer; ++k) {
    const bool in = a0 + k < g.A;
    const bool c = in && (m[a0 + k] > thr_c), b = in && (u[a0 + k] > thr_b);
    fc |= (unsigned)c << k;
    fb |= (unsigned)b << k;
    if (in) sel_flags[(size_t)n * g.A + a0 + k] = (uint8_t)((c ? 1 : 0) | (b ? 2 : 0));
  }
  const int nc = __popc(fc), nb = __popc(fb);
  int ic = nc, ib = nb;   // inclusive warp scans
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int tc = __shfl_up_sync(0xffffffffu, ic, o);
    const int tb = __shfl_up_sync(0xffffffffu, ib, o);
    if (lane >= o) { ic += tc; ib += tb; }
  }
  __syncthreads();
  if (lane == 31) { s_warp[0][warp] = ic; s_warp[1][warp] = ib; }
  __syncthreads();
  int oc = s_base[0] + ic - nc, ob = s_base[1] + ib - nb;
  int totc = 0, totb = 0;
  for (int w = 0; w < kSelThreads / 32; ++w) {
    if (w < warp) { oc += s_warp[0][w]; ob += s_warp[1][w]; }
    totc += s_warp[0][w];
    totb += s_warp[1][w];
  }
  int32_t* out_c = cls_inds + (size_t)n * g.sel_cap;
  int32_t* out_b = box_inds + (size_t)n * g.sel_cap;
#pragma unroll
  for (int k = 0; k < kSelPer; ++k) {
    if (fc & (1u << k)) out_c[oc++] = a0 + k;
    if (fb & (1u << k)) out_b[ob++] = a0 + k;
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
    cls_count[n] = s_base[0] + totc;
    box_count[n] = s_base[1] + totb;
  }
}

cudaError_t launch_ers(const Geo& g, const Workspace& ws, const Ptr5& t_cls, const Ptr5& t_box, int32_t* cls_inds,
                       int32_t* cls_count, int32_t* box_inds, int32_t* box_count, float* thr, uint8_t* sel_flags,
                       cudaStream_t st) {
  // ERD_SCAN_MODE=staged selects the non-persistent kernel (ERD_SCAN_TILE=64|128|256); default: pipelined
  static int mode = -1, forced = 0;
  if (mode < 0) {
    const char* m = getenv("ERD_SCAN_MODE");
    mode = (m && m[0] == 's') ? 1 : 0;
    const char* e = getenv("ERD_SCAN_TILE");
    forced = e ? atoi(e) : 0;
  }
  int tiles;
  if (mode == 0) {
    tiles = launch_scan_pipe(g, ws, t_cls, t_box, st);
  } else {
    const size_t row = (size_t)(g.ori + kBoxCh) * sizeof(float);
    const int T = forced ? forced : (row * 256 <= 110 * 1024 ? 256 : row * 128 <= 110 * 1024 ? 128 : 64);
    if (row * T > 220 * 1024) return cudaErrorInvalidValue;
    if (T == 256) tiles = launch_scan<256>(g, ws, t_cls, t_box, st);
    else if (T == 128) tiles = launch_scan<128>(g, ws, t_cls, t_box, st);
    else tiles = launch_scan<64>(g, ws, t_cls, t_box, st);
  }
  ERD_LAUNCH(kKErsSelect, st,
             (ers_select_kernel<<<dim3((g.A + kSelChunk - 1) / kSelChunk, g.n_img), kSelThreads, 0, st>>>(
                 g, ws, tiles, cls_inds, cls_count, box_inds, box_count, thr, sel_flags)));
  return cudaGetLastError();
}

}  // namespace erd
