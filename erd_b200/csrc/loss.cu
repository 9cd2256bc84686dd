// The sparse part of the loss path around the dense student pass (student.cu): the positives'
// prepass in front of the avg-factor all-reduce, the take-back pass after the teacher NMS and the
// final division into the reference's loss values.
// Reference: GFLHeadIncrementERD.loss_by_feat_single / distill_loss_by_image_single /
// loss_by_feat (dense_heads/gfl_head_increment_erd.py:142-454), distribution_focal_loss
// (losses/gfocal_loss.py:143-165), giou_loss (losses/iou_loss.py:110-126) over bbox_overlaps
// (structures/bbox/bbox_overlaps.py:151-199), weight_reduce_loss (losses/utils.py:30-65).
//
// Kernels (schedule: launch_loss at the end of this file)
//   pos_prepass_kernel     (erd_avg_factors) 4 threads per positive anchor: weight, softmax-integral
//                          decode, IoU score, GIoU/DFL loss sums, and the two avg factors.
//   assign_prepass_kernel  (erd_step_prepare) the ATSS decode and that prepass in one launch.
//   student_pass_kernel    (student.cu) every gradient element, written as if the NMS kept every
//                          ERS box candidate.
//   box_fix_kernel         after the NMS: takes the suppressed candidates back, sums the
//                          survivors' KL; its last block turns the fp64 accumulators into the
//                          reference's loss values (finalize_one).
#include <cstdlib>
#include "loss_math.cuh"

namespace erd {

// ----------------------------------------------------------------------------- positives
// Four threads per positive anchor, one per box side: weight, IoU score and loss sums (runs
// before the all-reduce; the gradients of the positives' rows are written by the student pass).
constexpr int kPosThreads = 256;

struct PosArgs {
  Ptr5 s_cls, s_box;
  const float* gt_boxes;
  const int64_t* gt_labels;
  const int32_t* gt_offsets;
  const int32_t* gt_inds;
  const int32_t* num_pos;
  float* avg;             // written by the last block
};

// One side of one positive anchor (a, assigned to GT row gidx); the four side threads are
// adjacent lanes.  Dead lanes (live == false) run along so the quad shuffles stay convergent.
__device__ __forceinline__ void pos_item(const Geo& g, const Workspace& ws, const PosArgs& A, int n, bool live, int a,
                                         int gidx, int side, double* s_acc) {
  // Dependent memory round trips are what this pass costs (it runs beside DRAM-saturating
  // kernels), so everything past the list entry is issued as one independent batch of loads.
  const int l = level_of_anchor(g, a);
  const int HW = g.hw[l];
  const int hw = a - g.start[l];
  const float* bplane = A.s_box.p[l] + ((size_t)n * kBoxCh + side * kBins) * HW + hw;
  float z[kBins];
#pragma unroll
  for (int j = 0; j < kBins; ++j) z[j] = live ? __ldg(bplane + (size_t)j * HW) : 0.f;
  // weight_targets: max_c sigmoid(new-class logits), detached (:283-284)
  float mx = -INFINITY;
  if (live) {
    const float* cplane = A.s_cls.p[l] + ((size_t)n * g.C + g.ori) * HW + hw;
    for (int c = side; c < g.cn; c += 4) mx = fmaxf(mx, __ldg(cplane + (size_t)c * HW));
  }
  const long long lab = live ? A.gt_labels[gidx] : -1;
  const float4 gb = live ? *reinterpret_cast<const float4*>(A.gt_boxes + (size_t)gidx * 4) : make_float4(0, 0, 1, 1);
  const bool on = live && lab >= 0 && lab < g.cn;                     // gfl_head_increment_erd.py:273-274
  mx = quad_max(mx);
  const float w = on ? sigmoid_ref(mx) : 0.f;
  const PosSide s = pos_side_decode(z, side, hw % g.w[l], hw / g.w[l], (float)g.stride[l], gb);
  if (live && side == 0) {   // what the student pass needs of this positive (pslot is filled in by whoever listed it)
    PosRec* rec = ws.pos_rec + (size_t)n * g.A + a;
    rec->gt = gb;
    rec->label = on ? (int)lab : -1;
    rec->score = s.g.iou;                                                                         // :289-292
    rec->w = w;
  }
  if (on) {
    const float giou = s.g.iou - (s.g.enc - s.g.uni) / s.g.enc;
    const float lse = s.zm + logf(s.sum);
    atomicAdd(&s_acc[kLevels + l], (double)(w * ((lse - s.zl) * s.t.wl + (lse - s.zr) * s.t.wr)));   // gfocal_loss.py:159-165
    if (side == 0) {
      atomicAdd(&s_acc[l], (double)(w * (1.0f - giou)));                                       // iou_loss.py:124-126
      atomicAdd(&s_acc[2 * kLevels], (double)w);
    }
  }
}

__global__ void __launch_bounds__(kPosThreads) pos_prepass_kernel(Geo g, Workspace ws, PosArgs A) {
  const int n = blockIdx.y;
  const int side = threadIdx.x & 3;
  const int np = A.num_pos[n];
  __shared__ double s_acc[2 * kLevels + 1];
  if (threadIdx.x < 2 * kLevels + 1) s_acc[threadIdx.x] = 0.0;
  __syncthreads();
  // whole warps stay together (8 positives per warp) so the quad shuffles are convergent
  for (int p = (blockIdx.x * kPosThreads + threadIdx.x) >> 2; p < ((np + 7) & ~7) && p < g.pos_cap;
       p += (gridDim.x * kPosThreads) >> 2) {
    const bool live = p < np;
    const int2 ent = live ? ws.pos_list[(size_t)n * g.A + p] : make_int2(0, 0);
    if (live && side == 0) ws.pos_rec[(size_t)n * g.A + ent.x].pslot = p;
    pos_item(g, ws, A, n, live, ent.x, ent.y, side, s_acc);
  }
  __syncthreads();
  if (threadIdx.x < 2 * kLevels + 1 && s_acc[threadIdx.x] != 0.0)
    atomicAdd(ws.pre_acc + threadIdx.x, s_acc[threadIdx.x]);
  // last block: avg[0] = sum_img max(num_pos, 1) (sampling_result.py:96-100), avg[1] = sum of weights
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(ws.counters, 1u) == gridDim.x * gridDim.y - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // publish the sums for finalize and leave the accumulators / ticket clean for the next call
  if (threadIdx.x < 2 * kLevels + 1) {
    const double v = ((volatile double*)ws.pre_acc)[threadIdx.x];
    ws.pre_pub[threadIdx.x] = v;
    ws.pre_acc[threadIdx.x] = 0.0;
    if (threadIdx.x == 2 * kLevels) A.avg[1] = (float)v;
  }
  if (threadIdx.x == 32) {
    long long cnt = 0;
    for (int i = 0; i < g.n_img; ++i) cnt += max(A.num_pos[i], 1);
    A.avg[0] = (float)cnt;
    ws.counters[0] = 0u;
  }
}

// ATSS decode + positives prepass in ONE launch (erd_step_prepare): the student-side chain in
// front of the sweeps is a sequence of short latency-bound kernels, so every launch boundary
// and every list round trip removed from it moves the sweeps earlier.  Per anchor like
// atss_finalize_kernel; a warp that found positives then works them off eight at a time
// (lane = positive-in-batch x side) with the body of pos_prepass_kernel.  The last block
// publishes num_pos and both avg factors.
constexpr int kAssignPer = 4;   // anchors per thread: the whole grid is one wave

__global__ void __launch_bounds__(256) assign_prepass_kernel(Geo g, Workspace ws, PosArgs A,
                                                             const int32_t* __restrict__ pad_hw,
                                                             int32_t* __restrict__ gt_inds,
                                                             int32_t* __restrict__ num_pos, ExchangeInfo xchg) {
  const int n = blockIdx.y;
  const int lane = threadIdx.x & 31;
  __shared__ double s_acc[2 * kLevels + 1];
  if (threadIdx.x < 2 * kLevels + 1) s_acc[threadIdx.x] = 0.0;
  __syncthreads();
  // one batch of independent loads: the argmax-table entries of this thread's anchors.  With few GT boxes a CTA takes
  // one contiguous 1024-anchor range.  With many (more than 24 per image on average) a warp's four 32-anchor chunks
  // are dealt round-robin over the image's CTAs instead (chunk c -> CTA c mod gridDim.x): positives then cluster --
  // nearly every anchor of the coarse levels is one -- and the CTA that owned those levels alone took several times as
  // long as the rest of the grid (dense config: 107 -> 41 us); the scattered chunks cost the sparse case 2-3 us.
  const int warp = threadIdx.x >> 5;
  const bool spread = A.gt_offsets[g.n_img] > 24 * g.n_img;   // grid-uniform
  int anchor[kAssignPer];
  unsigned long long key[kAssignPer];
#pragma unroll
  for (int i = 0; i < kAssignPer; ++i) {
    anchor[i] = spread ? ((i * 8 + warp) * (int)gridDim.x + (int)blockIdx.x) * 32 + lane
                       : (int)blockIdx.x * (256 * kAssignPer) + (int)threadIdx.x + i * 256;
    key[i] = anchor[i] < g.A ? ws.atss_key[(size_t)n * g.A + anchor[i]] : 0ull;
  }
  const int pad_h = pad_hw[n * 2], pad_w = pad_hw[n * 2 + 1], first_gt = A.gt_offsets[n];
  // decode all of the warp's 128 anchors first and collect its positives in a shared-memory list: the list is then
  // worked off eight positives at a time, so a warp pays one round of dependent loads per eight positives (not one per
  // 32-anchor row that happens to hold a positive) and one returning atomic for its slots in the image's list
  __shared__ int2 s_list[256 / 32][32 * kAssignPer];
  int cnt = 0;   // warp-uniform
#pragma unroll
  for (int i = 0; i < kAssignPer; ++i) {
    const int a = anchor[i];
    const int gidx = a < g.A ? atss_decode_only(g, ws, pad_h, pad_w, first_gt, gt_inds, n, a, key[i]) : -1;
    const unsigned m = __ballot_sync(0xffffffffu, gidx >= 0);
    if (gidx >= 0) s_list[warp][cnt + __popc(m & ((1u << lane) - 1u))] = make_int2(a, gidx);
    cnt += __popc(m);
  }
  if (cnt) {   // warp-uniform
    int base = 0;
    if (lane == 0) base = atomicAdd(ws.pos_counter + n, cnt);
    base = __shfl_sync(0xffffffffu, base, 0);
    __syncwarp();
    for (int e = lane; e < cnt; e += 32) {   // the image's positives list and the slot of each positive
      const int2 ent = s_list[warp][e];
      ws.pos_list[(size_t)n * g.A + base + e] = ent;
      ws.pos_rec[(size_t)n * g.A + ent.x].pslot = base + e;
    }
    for (int e0 = 0; e0 < cnt; e0 += 8) {   // lane = positive-in-round x side
      const int e = e0 + (lane >> 2);
      const bool live = e < cnt;
      const int2 ent = live ? s_list[warp][e] : make_int2(0, 0);
      pos_item(g, ws, A, n, live, ent.x, ent.y, lane & 3, s_acc);
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * kLevels + 1 && s_acc[threadIdx.x] != 0.0)
    atomicAdd(ws.pre_acc + threadIdx.x, s_acc[threadIdx.x]);
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(ws.counters + 3, 1u) == gridDim.x * gridDim.y - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < 2 * kLevels + 1) {
    const double v = ((volatile double*)ws.pre_acc)[threadIdx.x];
    ws.pre_pub[threadIdx.x] = v;
    ws.pre_acc[threadIdx.x] = 0.0;
    if (threadIdx.x == 2 * kLevels) A.avg[1] = (float)v;
  }
  __shared__ float s_avg[2];
  __shared__ unsigned int s_epoch;
  if (threadIdx.x == 2 * kLevels) s_avg[1] = (float)((volatile double*)ws.pre_pub)[2 * kLevels];
  if (threadIdx.x >= 32 && threadIdx.x < 64) {   // avg[0] = sum_img max(num_pos, 1) (sampling_result.py:96-100)
    long long cnt = 0;
    for (int i = lane; i < g.n_img; i += 32) {
      const int np = ((volatile int*)ws.pos_counter)[i];
      num_pos[i] = np;
      ws.pos_counter[i] = 0;
      cnt += max(np, 1);
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) {
      A.avg[0] = (float)cnt;
      s_avg[0] = (float)cnt;
      ws.counters[3] = 0u;
    }
  }
  // Data-parallel runs: post the two local factors to the peers right here -- the student pass waits for theirs in
  // its prologue, so the exchange costs neither a launch nor a place on the caller's stream (reduce_mean,
  // mmdet/utils/dist_utils.py:59-65; call sites gfl_head_increment_erd.py:390-391,406-407)
  if (xchg.world > 1) {
    __syncthreads();
    exchange_post(xchg, s_avg[0], s_avg[1], &s_epoch);
  }
}


// The distillation only counts for the candidates the teacher NMS keeps, which is not known while
// the student pass runs (it runs BESIDE the NMS and writes every candidate as if kept).  After the
// NMS: sum the weighted KL of the survivors (flag bit 2 set by the resolve pass) and rewrite the
// gradient rows of the suppressed candidates without the distillation term.
constexpr int kLateThreads = 256;

__device__ __forceinline__ void finalize_one(const Geo& g, const Workspace& ws, const LossArgs& A, int i);

__global__ void __launch_bounds__(kLateThreads) box_fix_kernel(Geo g, Workspace ws, LossArgs A) {
  if (A.skip_flag && *A.skip_flag == 0u) return;
  const int n = blockIdx.y;
  const int side = threadIdx.x & 3;
  const int K = A.box_count[n];
  float kd = 0.f;
  for (int r = (blockIdx.x * kLateThreads + threadIdx.x) >> 2; r < K; r += (gridDim.x * kLateThreads) >> 2) {
    const int a = A.box_inds[(size_t)n * g.sel_cap + r];
    const size_t ga = (size_t)n * g.A + a;
    if (A.sel_flags[ga] & 4) {
      if (side == 0) kd += ws.kd_loss[ga];
      continue;
    }
    const int l = level_of_anchor(g, a);
    const int HW = g.hw[l];
    const int hw = a - g.start[l];
    const float* prow = nullptr;
    if (A.gt_inds[ga] > 0) prow = ws.pos_rows + ((size_t)n * g.pos_cap + ws.pos_rec[ga].pslot) * kBoxCh + side * kBins;
    float* gp = A.g_box.p[l] + ((size_t)n * kBoxCh + side * kBins) * HW + hw;
#pragma unroll
    for (int j = 0; j < kBins; ++j) gp[(size_t)j * HW] = prow ? prow[j] : 0.f;
  }
  __shared__ float red[kLateThreads / 32];
  __shared__ bool last;
  kd = warp_sum(kd);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = kd;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s2 = 0.0;
    for (int w = 0; w < kLateThreads / 32; ++w) s2 += (double)red[w];
    if (s2 != 0.0) atomicAdd(ws.loss_acc + acc_dbox(g, n), s2);
    // This is the last launch of the step (every other kernel is ordered before it): its last
    // block turns the accumulators into the loss vector, which saves a launch at the very end
    // of the critical path.
    __threadfence();
    last = atomicAdd(ws.counters + 2, 1u) == gridDim.x * gridDim.y - 1;
  }
  __syncthreads();
  if (last) {
    __threadfence();
    for (int i = threadIdx.x; i < 3 * kLevels + 2 * g.n_img; i += kLateThreads) finalize_one(g, ws, A, i);
    if (threadIdx.x == 0) ws.counters[2] = 0u;
  }
}

// ws.counters[1] = 1 when any upstream gradient differs from 1
__global__ void upstream_check_kernel(Workspace ws, const float* __restrict__ up, int total) {
  bool nonunit = false;
  for (int i = threadIdx.x; i < total; i += blockDim.x) nonunit |= (up[i] != 1.0f);
  const int any = __syncthreads_or(nonunit ? 1 : 0);
  if (threadIdx.x == 0) ws.counters[1] = any ? 1u : 0u;
}

// Accumulators -> the reference's loss values, with its division order.
__device__ __forceinline__ void finalize_one(const Geo& g, const Workspace& ws, const LossArgs& A, int i) {
  const int total = 3 * kLevels + 2 * g.n_img;
  if (i >= total) return;
  const float avg2 = fmaxf(A.avg[1], 1.0f);                                        // :407 clamp_(min=1)
  float out;
  if (i < kLevels) {
    out = g.w_cls * ((float)((volatile double*)ws.loss_acc)[i] / (float)((double)A.avg[0] + (double)kEps32));
  } else if (i < 2 * kLevels) {
    out = g.w_bbox * ((float)ws.pre_pub[i - kLevels] / (1.0f + kEps32)) / avg2;   // :299-303,408
  } else if (i < 3 * kLevels) {
    out = g.w_dfl * ((float)ws.pre_pub[i - kLevels] / 4.0f) / avg2;               // :306-310,409
  } else if (i < 3 * kLevels + g.n_img) {
    const int n = i - 3 * kLevels;
    out = A.dlw * (float)(((volatile double*)ws.loss_acc)[i] / ((double)A.cls_count[n] * (double)g.ori));   // mean over K*ori; 0/0 -> NaN
  } else {
    out = A.dlw * (g.w_ld * ((float)((volatile double*)ws.loss_acc)[i] / 4.0f));
  }
  A.losses[i] = out;
  ws.loss_acc[i] = 0.0;   // clean for the next step
}

static int pos_grid_x(const Geo& g) {
  long long blocks = ((long long)g.pos_cap * 4 + kPosThreads - 1) / kPosThreads;
  return (int)(blocks < 1 ? 1 : (blocks > 32 ? 32 : blocks));
}

cudaError_t launch_avg(const Geo& g, const Workspace& ws, const Ptr5& s_cls, const Ptr5& s_box,
                       const float* gt_boxes, const int64_t* gt_labels, const int32_t* gt_offsets,
                       const int32_t* gt_inds, const int32_t* num_pos, float* avg, cudaStream_t st) {
  PosArgs a;
  a.s_cls = s_cls;
  a.s_box = s_box;
  a.gt_boxes = gt_boxes;
  a.gt_labels = gt_labels;
  a.gt_offsets = gt_offsets;
  a.gt_inds = gt_inds;
  a.num_pos = num_pos;
  a.avg = avg;
  ERD_LAUNCH(kKAvg, st, (pos_prepass_kernel<<<dim3(pos_grid_x(g), g.n_img), kPosThreads, 0, st>>>(g, ws, a)));
  return cudaGetLastError();
}

cudaError_t launch_assign_avg(const Geo& g, const Workspace& ws, const Ptr5& s_cls, const Ptr5& s_box,
                              const float* gt_boxes, const int64_t* gt_labels, const int32_t* gt_offsets,
                              const int32_t* pad_hw, int32_t* gt_inds, int32_t* num_pos, float* avg,
                              const ExchangeInfo* xchg, cudaStream_t st) {
  cudaError_t e = launch_atss_candidates(g, ws, gt_boxes, gt_offsets, pad_hw, st);
  if (e != cudaSuccess) return e;
  PosArgs a;
  a.s_cls = s_cls;
  a.s_box = s_box;
  a.gt_boxes = gt_boxes;
  a.gt_labels = gt_labels;
  a.gt_offsets = gt_offsets;
  a.gt_inds = gt_inds;
  a.num_pos = num_pos;
  a.avg = avg;
  ExchangeInfo x;
  if (xchg) x = *xchg;
  else x.world = 0;
  ERD_LAUNCH(kKAvg, st,
             (assign_prepass_kernel<<<dim3((g.A + 256 * kAssignPer - 1) / (256 * kAssignPer), g.n_img), 256, 0, st>>>(
                 g, ws, a, pad_hw, gt_inds, num_pos, x)));
  return cudaGetLastError();
}
cudaError_t launch_student(const Geo& g, const Workspace& ws, const LossArgs& a, cudaStream_t st);   // student.cu

// Schedule of the loss call.  Everything dense is the student pass (caller's stream); it needs the
// ERS selection (sel_ready) and the reduced avg factors, not the NMS.  The take-back pass joins
// the NMS and, as the last launch of the step, also finalizes the loss vector.
cudaError_t launch_loss(const Geo& g, const Workspace& ws, const LossArgs& a, cudaStream_t st, const LossStreams* ls) {
  const int total = 3 * kLevels + 2 * g.n_img;
  cudaError_t e = cudaSuccess;   // accumulators are left clean by the previous finalize (erd_workspace_init once)
  if (a.skip_flag) ERD_LAUNCH(kKUpCheck, st, (upstream_check_kernel<<<1, 128, 0, st>>>(ws, a.upstream, total)));
  if (ls && ls->sel_ready) {
    e = cudaStreamWaitEvent(st, ls->sel_ready, 0);
    if (e != cudaSuccess) return e;
  }
  e = launch_student(g, ws, a, st);
  if (e != cudaSuccess) return e;
  if (ls && ls->nms_done) {
    e = cudaStreamWaitEvent(st, ls->nms_done, 0);
    if (e != cudaSuccess) return e;
  }
  const dim3 late_grid((g.sel_cap * 4 + kLateThreads - 1) / kLateThreads, g.n_img);
  ERD_LAUNCH(kKBoxFix, st, (box_fix_kernel<<<late_grid, kLateThreads, 0, st>>>(g, ws, a)));
  return cudaGetLastError();
}

}  // namespace erd
