// The two normalisers of the GT losses, computed right after assignment and left on the
// device so that one 8-byte all-reduce serves both.
// Reference: avg_factor = sum_img max(num_pos, 1) (gfl_head.py:548-549,
// samplers/sampling_result.py:96-100) reduced at gfl_head_increment_erd.py:390-391, and
// sum of weight_targets = max_c sigmoid(student new-class logits) over positives
// (gfl_head_increment_erd.py:283-284,322,406-407).
#include "erd_common.cuh"

namespace erd {

constexpr int kAvgThreads = 256;

__global__ void __launch_bounds__(kAvgThreads) avg_kernel(Geo g, Workspace ws, Ptr5 s_cls,
                                                          const int64_t* __restrict__ gt_labels,
                                                          const int32_t* __restrict__ gt_offsets,
                                                          const int32_t* __restrict__ gt_inds,
                                                          const int32_t* __restrict__ num_pos,
                                                          float* __restrict__ avg) {
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int np = num_pos[n];
  double wsum = 0.0;
  for (int i = warp; i < np; i += kAvgThreads / 32) {
    const int a = ws.pos_list[(size_t)n * g.A + i];
    const int gi = gt_inds[(size_t)n * g.A + a];
    const long long label = gt_labels[gt_offsets[n] + gi - 1];
    if (label < 0 || label >= g.cn) continue;   // gfl_head_increment_erd.py:273-274
    const int l = level_of_anchor(g, a);
    const int hw = a - g.start[l];
    const float* plane = s_cls.p[l] + ((size_t)n * g.C + g.ori) * g.hw[l] + hw;
    float mx = -INFINITY;
    for (int c = lane; c < g.cn; c += 32) mx = fmaxf(mx, __ldg(plane + (size_t)c * g.hw[l]));
    mx = warp_max(mx);
    wsum += (double)sigmoid_ref(mx);   // identical on every lane
  }
  __shared__ double red[kAvgThreads / 32];
  __shared__ bool last;
  if (lane == 0) red[warp] = wsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kAvgThreads / 32; ++w) s += red[w];
    ws.avg_part[n] = s;
    __threadfence();
    last = atomicAdd(ws.counters, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double tot = 0.0;
    long long cnt = 0;
    for (int i = 0; i < g.n_img; ++i) {
      tot += ((volatile double*)ws.avg_part)[i];
      cnt += max(num_pos[i], 1);
    }
    avg[0] = (float)cnt;
    avg[1] = (float)tot;
    ws.counters[0] = 0;
  }
}

cudaError_t launch_avg(const Geo& g, const Workspace& ws, const Ptr5& s_cls, const int64_t* gt_labels,
                       const int32_t* gt_offsets, const int32_t* gt_inds, const int32_t* num_pos, float* avg,
                       cudaStream_t st) {
  ERD_LAUNCH(kKAvg, st,
             (avg_kernel<<<g.n_img, kAvgThreads, 0, st>>>(g, ws, s_cls, gt_labels, gt_offsets, gt_inds, num_pos, avg)));
  return cudaGetLastError();
}

}  // namespace erd
