// Elastic Response Selection: one streaming pass over the teacher head outputs, then
// per-image thresholds and ordered stream compaction.
// Reference: GFLIncrementERD.sel_pos / sel_pos_single
// (mmdet/models/detectors/gfl_increment_erd.py:143-200); the Integral decode fused into the
// pass is gfl_head_increment_erd.py:40-54,189-195.
#include "erd_common.cuh"

namespace erd {

// ----------------------------------------------------------------------------- pass 1
// Per anchor: m = max_c sigmoid(t_cls) (sigmoid is monotone, so sigmoid(max logit)), the
// first argmax class, u = max_j raw box logit, and the four softmax-integral distances.
// Per CTA: sums of m, m^2, u, u^2 in fp64 (deterministic two-level reduction).
template <bool VEC>
__device__ __forceinline__ void ers_tile(const Geo& g, const Workspace& ws, const float* __restrict__ cls,
                                         const float* __restrict__ box, int n, int l, int hw0, double (&acc)[4]) {
  const int HW = g.hw[l];
  const Quad<VEC> q(hw0, HW);
  float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int arg[4] = {0, 0, 0, 0};
  const float* cplane = cls + (size_t)n * g.ori * HW;
#pragma unroll 8
  for (int c = 0; c < g.ori; ++c) {
    float v[4];
    q.load(cplane + (size_t)c * HW, v, -INFINITY);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (v[k] > best[k]) { best[k] = v[k]; arg[k] = c; }
  }
  float u[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  float dist[4][4];
  const float* bplane = box + (size_t)n * kBoxCh * HW;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    // one streaming pass per side: exponentials are taken relative to the first bin, so no
    // 17-value register tile is needed; a non-finite sum (logit spread > 88) redoes the side
    // with the true maximum.
    const float* splane = bplane + (size_t)(s * kBins) * HW;
    float ref[4], sum[4], num[4], mx[4];
    q.load(splane, ref, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) { sum[k] = 1.f; num[k] = 0.f; mx[k] = ref[k]; }
#pragma unroll 4
    for (int j = 1; j < kBins; ++j) {
      float v[4];
      q.load(splane + (size_t)j * HW, v, 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float e = expf(v[k] - ref[k]);
        sum[k] += e;
        num[k] = fmaf((float)j, e, num[k]);
        mx[k] = fmaxf(mx[k], v[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!(sum[k] < 3.0e38f) || !(num[k] < 3.0e38f)) {   // overflow: exact two-pass fallback
        float s2 = 0.f, n2 = 0.f;
        for (int j = 0; j < kBins; ++j) {
          const float e = expf(__ldg(splane + (size_t)j * HW + q.hw[k]) - mx[k]);
          s2 += e;
          n2 = fmaf((float)j, e, n2);
        }
        sum[k] = s2;
        num[k] = n2;
      }
      dist[k][s] = __fdiv_rn(num[k], sum[k]);
      u[k] = fmaxf(u[k], mx[k]);
    }
  }
  const size_t base = (size_t)n * g.A + g.start[l];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (!q.ok[k]) continue;
    const float m = sigmoid_ref(best[k]);
    const size_t a = base + q.hw[k];
    ws.t_m[a] = m;
    ws.t_arg[a] = arg[k];
    ws.t_u[a] = u[k];
    ws.t_dist[a] = make_float4(dist[k][0], dist[k][1], dist[k][2], dist[k][3]);
    acc[0] += (double)m;
    acc[1] += (double)m * (double)m;
    acc[2] += (double)u[k];
    acc[3] += (double)u[k] * (double)u[k];
  }
}

__global__ void __launch_bounds__(kTileThreads) ers_scan_kernel(Geo g, Workspace ws, Ptr5 t_cls, Ptr5 t_box) {
  const int n = blockIdx.y;
  const int tile = blockIdx.x;
  const int l = level_of_tile(g, tile);
  const int hw0 = (tile - g.tile_start[l]) * kTile;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  if (g.vec[l])
    ers_tile<true>(g, ws, t_cls.p[l], t_box.p[l], n, l, hw0, acc);
  else
    ers_tile<false>(g, ws, t_cls.p[l], t_box.p[l], n, l, hw0, acc);
  __shared__ double red[kTileThreads / 32][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = warp_sum(acc[i]);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) red[threadIdx.x >> 5][i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int w = 0; w < kTileThreads / 32; ++w) s += red[w][threadIdx.x];
    ws.ers_part[((size_t)n * gridDim.x + tile) * 4 + threadIdx.x] = s;
  }
}

// ----------------------------------------------------------------------------- pass 2
// thr = mean + 2 * std (unbiased); rows strictly above it, in ascending anchor order (what
// nonzero() yields, gfl_increment_erd.py:150-151,158-159).  Each CTA owns a 2048-anchor chunk
// of one image; it recounts the flags of the anchors before its chunk from the L2-resident
// cache instead of waiting on a cross-CTA prefix, so one launch suffices.
constexpr int kSelThreads = 256;
constexpr int kSelPer = 8;
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
  if (threadIdx.x < 2) {
    double s1 = 0.0, s2 = 0.0;
    const double* p = ws.ers_part + (size_t)n * tiles * 4 + threadIdx.x * 2;
    for (int t = 0; t < tiles; ++t) { s1 += p[t * 4]; s2 += p[t * 4 + 1]; }
    const double A = (double)g.A;
    const double mean = s1 / A;
    double var = (s2 - s1 * s1 / A) / (A - 1.0);   // A == 1 -> NaN, as torch.std
    if (var < 0.0) var = 0.0;
    const float t = __fadd_rn((float)mean, __fmul_rn(2.0f, (float)sqrt(var)));
    s_thr[threadIdx.x] = t;
    if (blockIdx.x == 0) thr_out[n * 2 + threadIdx.x] = t;
  }
  __syncthreads();
  const float thr_c = s_thr[0], thr_b = s_thr[1];
  const float* m = ws.t_m + (size_t)n * g.A;
  const float* u = ws.t_u + (size_t)n * g.A;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // selected rows before this chunk
  int pc = 0, pb = 0;
  for (int a = threadIdx.x; a < chunk0; a += kSelThreads) {
    pc += m[a] > thr_c;
    pb += u[a] > thr_b;
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
  const int tiles = g.tile_start[kLevels];
  ERD_LAUNCH(kKErsScan, st, (ers_scan_kernel<<<dim3(tiles, g.n_img), kTileThreads, 0, st>>>(g, ws, t_cls, t_box)));
  ERD_LAUNCH(kKErsSelect, st,
             (ers_select_kernel<<<dim3((g.A + kSelChunk - 1) / kSelChunk, g.n_img), kSelThreads, 0, st>>>(
                 g, ws, tiles, cls_inds, cls_count, box_inds,
                                                                 box_count, thr, sel_flags)));
  return cudaGetLastError();
}

}  // namespace erd
