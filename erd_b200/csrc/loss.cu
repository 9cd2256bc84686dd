// Fused forward + backward of the GT losses (QFL / GIoU / DFL) and the two distillation
// losses, writing dense NCHW gradients of the student head outputs.
// Reference: GFLHeadIncrementERD.loss_by_feat_single / distill_loss_by_image_single /
// loss_by_feat (dense_heads/gfl_head_increment_erd.py:142-454), quality_focal_loss and
// distribution_focal_loss (losses/gfocal_loss.py:12-53,143-165), giou_loss
// (losses/iou_loss.py:110-126) over bbox_overlaps (structures/bbox/bbox_overlaps.py:151-199),
// knowledge_distillation_kl_div_loss (losses/kd_loss.py:12-37), weight_reduce_loss
// (losses/utils.py:30-65).  Closed-form gradients: SURVEY.md Appendix A.
//
// Kernels (schedule: launch_loss at the end of this file)
//   pos_kernel<false>      (erd_avg_factors) 4 threads per positive anchor: weight, softmax-integral
//                          decode, IoU score, GIoU/DFL loss sums, and the two avg factors.
//   assign_prepass_kernel  (erd_step_prepare) the ATSS decode and that prepass in one launch.
//   zero_fill_kernel       the structurally-zero part of the gradient (old-class channels, all
//                          box channels) with streaming stores, at the start of the step.
//   cls_sweep_kernel       the one dense read-modify-write pass: QFL on the new-class channels.
//   cls_kd_kernel          class-response L2 rows of the ERS set, list driven.
//   pos_kernel<true>       box-logit gradient rows of the positives (needs the reduced avg
//                          factor): compact copy + scattered into the gradient tensor.
//   box_kd_kernel          rows of every ERS box candidate incl. the DFL-distribution KL, written
//                          beside the NMS as if kept.
//   box_fix_kernel         after the NMS: takes the suppressed candidates back, sums the
//                          survivors' KL; its last block turns the fp64 accumulators into the
//                          reference's loss values (finalize_one).
#include <cstdlib>
#include "erd_common.cuh"

namespace erd {

// accumulator layout == loss vector layout
__device__ __forceinline__ int acc_cls(int l) { return l; }
__device__ __forceinline__ int acc_bbox(int l) { return kLevels + l; }
__device__ __forceinline__ int acc_dfl(int l) { return 2 * kLevels + l; }
__device__ __forceinline__ int acc_dcls(int n) { return 3 * kLevels + n; }
__device__ __forceinline__ int acc_dbox(const Geo& g, int n) { return 3 * kLevels + g.n_img + n; }

__device__ __forceinline__ float upstream_of(const float* up, int i) { return up ? up[i] : 1.0f; }

struct QflTerm {
  float loss, grad;
};

// sigma and softplus from one exp: e = exp(-|x|) in (0, 1].
//   sigma    = 1/(1+e) for x >= 0, e/(1+e) otherwise
//   softplus = max(x, 0) + log1p(e), with log1p(e) = 2 atanh(s), s = e / (2 + e) in [0, 1/3]:
//              the odd series through s^13 is exact to ~1e-7 relative on the whole range,
//              without the cancellation a log(1+e) has for the small e of background anchors.
__device__ __forceinline__ void sig_sp(float x, float& sig, float& sp) {
  const float e = ex2_approx(-1.4426950408889634f * fabsf(x));
  const float r = __fdividef(1.0f, 1.0f + e);
  sig = x >= 0.f ? r : e * r;
  const float s = __fdividef(e, 2.0f + e);
  const float s2 = s * s;
  float p = fmaf(s2, 1.0f / 13.0f, 1.0f / 11.0f);
  p = fmaf(p, s2, 1.0f / 9.0f);
  p = fmaf(p, s2, 1.0f / 7.0f);
  p = fmaf(p, s2, 1.0f / 5.0f);
  p = fmaf(p, s2, 1.0f / 3.0f);
  p = fmaf(p, s2, 1.0f);
  sp = fmaf(2.0f * s, p, fmaxf(x, 0.f));
}

// negatives: BCE(x, 0) * sigma^2 (gfocal_loss.py:36-41)
__device__ __forceinline__ QflTerm qfl_neg(float x) {
  float sig, sp;
  sig_sp(x, sig, sp);
  const float s2 = sig * sig;
  return {sp * s2, s2 * (sig + 2.0f * sp * (1.0f - sig))};
}

// the label channel of a positive: BCE(x, score) * |score - sigma|^2 (gfocal_loss.py:47-50)
__device__ __forceinline__ QflTerm qfl_pos(float x, float score) {
  float sig, sp;
  sig_sp(x, sig, sp);
  const float bce = sp - score * x;
  const float d = score - sig;
  return {bce * d * d, (sig - score) * d * d - 2.0f * bce * d * sig * (1.0f - sig)};
}

// torch autograd of elementwise max/min routes the gradient to the selected operand and
// splits it evenly on exact ties.
__device__ __forceinline__ float pick_gt(float a, float b) { return a > b ? 1.0f : (a == b ? 0.5f : 0.0f); }

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}

// ----------------------------------------------------------------------------- positives
// Four threads per positive anchor, one per box side.  GRAD=false: weight, IoU score and
// loss sums (runs before the all-reduce); GRAD=true: gradients w.r.t. the 68 box logits.
constexpr int kPosThreads = 256;

struct PosArgs {
  Ptr5 s_cls, s_box;
  MPtr5 g_box;
  const float* gt_boxes;
  const int64_t* gt_labels;
  const int32_t* gt_offsets;
  const int32_t* gt_inds;
  const int32_t* num_pos;
  float* avg;             // GRAD=false: written by the last block; GRAD=true: read
  const float* upstream;
  const unsigned int* skip_flag;
};

// One side of one positive anchor (a, assigned to GT row gidx); the four side threads are
// adjacent lanes.  Dead lanes (live == false) run along so the quad shuffles stay convergent.
template <bool GRAD>
__device__ __forceinline__ void pos_item(const Geo& g, const Workspace& ws, const PosArgs& A, int n, bool live, int a,
                                         int gidx, int p, int side, double* s_acc, float avg2) {
  // Dependent memory round trips are what this kernel costs (it runs beside DRAM-saturating
  // sweeps), so everything past the list entry is issued as one independent batch of loads.
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
  // DFL target of this side: bbox2distance clamped to [0, reg_max - 0.1] (transforms.py:221-230)
  const float fs = (float)g.stride[l];
  const float cx = (float)(hw % g.w[l]), cy = (float)(hw / g.w[l]);
  const float tx1 = gb.x / fs, ty1 = gb.y / fs, tx2 = gb.z / fs, ty2 = gb.w / fs;   // :288
  const float tgt = side == 0 ? cx - tx1 : side == 1 ? cy - ty1 : side == 2 ? tx2 - cx : ty2 - cy;
  const float y = fminf(fmaxf(tgt, 0.f), (float)(kBins - 1) - 0.1f);
  const int yl = (int)y;
  const float wl = (float)(yl + 1) - y, wr = y - (float)yl;
  float zl = 0.f, zr = 0.f;   // the two logits the DFL cross-entropy reads
#pragma unroll
  for (int j = 0; j < kBins; ++j) {
    zl = j == yl ? z[j] : zl;
    zr = j == yl + 1 ? z[j] : zr;
  }
  // Integral of this thread's side: softmax expectation (:40-54,285)
  float zm = z[0];
#pragma unroll
  for (int j = 1; j < kBins; ++j) zm = fmaxf(zm, z[j]);
  float sum = 0.f, num = 0.f;
#pragma unroll
  for (int j = 0; j < kBins; ++j) {
    z[j] = expf(z[j] - zm);
    sum += z[j];
    num = fmaf((float)j, z[j], num);
  }
  const float inv = 1.0f / sum;
  const float dmine = num * inv;
  float d[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) d[s] = __shfl_sync(0xffffffffu, dmine, (threadIdx.x & 28) | s, 32);
  // anchor centre / stride is the grid coordinate itself (gfl_head.py:232-243, :281)
  const float px1 = cx - d[0], py1 = cy - d[1], px2 = cx + d[2], py2 = cy + d[3];   // distance2bbox
  // aligned IoU / GIoU (bbox_overlaps.py:151-169,189-199), eps 1e-6
  const float area_p = (px2 - px1) * (py2 - py1);
  const float area_t = (tx2 - tx1) * (ty2 - ty1);
  const float iw_raw = fminf(px2, tx2) - fmaxf(px1, tx1), ih_raw = fminf(py2, ty2) - fmaxf(py1, ty1);
  const float iw = fmaxf(iw_raw, 0.f), ih = fmaxf(ih_raw, 0.f);
  const float inter = iw * ih;
  const float uni_raw = area_p + area_t - inter;
  const float uni = fmaxf(uni_raw, 1e-6f);
  const float iou = inter / uni;
  const float ew_raw = fmaxf(px2, tx2) - fminf(px1, tx1), eh_raw = fmaxf(py2, ty2) - fminf(py1, ty1);
  const float ew = fmaxf(ew_raw, 0.f), eh = fmaxf(eh_raw, 0.f);
  const float enc_raw = ew * eh;
  const float enc = fmaxf(enc_raw, 1e-6f);
  if (!GRAD) {
    if (on) {
      const float giou = iou - (enc - uni) / enc;
      const float lse = zm + logf(sum);
      atomicAdd(&s_acc[kLevels + l], (double)(w * ((lse - zl) * wl + (lse - zr) * wr)));   // gfocal_loss.py:159-165
      if (side == 0) {
        ws.pos_score[(size_t)n * g.A + a] = iou;                                              // :289-292
        atomicAdd(&s_acc[l], (double)(w * (1.0f - giou)));                                   // iou_loss.py:124-126
        atomicAdd(&s_acc[2 * kLevels], (double)w);
      }
    }
  } else if (on) {
    // d(1 - giou) / d(px1, py1, px2, py2), then through distance2bbox to this side's distance
    const float g_uni = (inter / (uni * uni) - 1.0f / enc) * pick_gt(uni_raw, 1e-6f);
    const float g_int = -1.0f / uni - g_uni;
    const float g_enc = (uni / (enc * enc)) * pick_gt(enc_raw, 1e-6f);
    const float g_iw = g_int * ih * (iw_raw >= 0.f ? 1.f : 0.f);
    const float g_ih = g_int * iw * (ih_raw >= 0.f ? 1.f : 0.f);
    const float g_ew = g_enc * eh * (ew_raw >= 0.f ? 1.f : 0.f);
    const float g_eh = g_enc * ew * (eh_raw >= 0.f ? 1.f : 0.f);
    const float hgt = py2 - py1, wid = px2 - px1;
    float gd;
    if (side == 0) gd = g_uni * hgt + g_iw * pick_gt(px1, tx1) + g_ew * pick_gt(tx1, px1);        // -d/dx1
    else if (side == 1) gd = g_uni * wid + g_ih * pick_gt(py1, ty1) + g_eh * pick_gt(ty1, py1);  // -d/dy1
    else if (side == 2) gd = g_uni * hgt + g_iw * pick_gt(tx2, px2) + g_ew * pick_gt(px2, tx2);  // d/dx2
    else gd = g_uni * wid + g_ih * pick_gt(ty2, py2) + g_eh * pick_gt(py2, ty2);                  // d/dy2
    const float cb = upstream_of(A.upstream, acc_bbox(l)) * g.w_bbox / (1.0f + kEps32) / avg2 * w * gd;
    const float cd = upstream_of(A.upstream, acc_dfl(l)) * g.w_dfl / 4.0f / avg2 * w;
    // compact copy (the candidates' kernels merge it) + the gradient tensor itself
    float* row = ws.pos_rows + ((size_t)n * g.pos_cap + p) * kBoxCh + side * kBins;
    float* gp = A.g_box.p[l] + ((size_t)n * kBoxCh + side * kBins) * HW + hw;
#pragma unroll
    for (int j = 0; j < kBins; ++j) {
      const float pj = z[j] * inv;
      float gr = cb * pj * ((float)j - dmine);
      gr += cd * (wl * (pj - (j == yl ? 1.f : 0.f)) + wr * (pj - (j == yl + 1 ? 1.f : 0.f)));
      row[j] = gr;
      gp[(size_t)j * HW] = gr;
    }
  } else if (GRAD && live) {   // assigned to a GT whose label lies outside the new-class range: no box loss
    float* row = ws.pos_rows + ((size_t)n * g.pos_cap + p) * kBoxCh + side * kBins;
#pragma unroll
    for (int j = 0; j < kBins; ++j) row[j] = 0.f;
  }
}

template <bool GRAD>
__global__ void __launch_bounds__(kPosThreads) pos_kernel(Geo g, Workspace ws, PosArgs A) {
  if (GRAD && A.skip_flag && *A.skip_flag == 0u) return;
  const int n = blockIdx.y;
  const int side = threadIdx.x & 3;
  const int np = A.num_pos[n];
  __shared__ double s_acc[2 * kLevels + 1];
  if (!GRAD) {
    if (threadIdx.x < 2 * kLevels + 1) s_acc[threadIdx.x] = 0.0;
    __syncthreads();
  }
  const float avg2 = GRAD ? fmaxf(A.avg[1], 1.0f) : 1.0f;
  // whole warps stay together (8 positives per warp) so the quad shuffles are convergent
  for (int p = (blockIdx.x * kPosThreads + threadIdx.x) >> 2; p < ((np + 7) & ~7) && p < g.pos_cap;
       p += (gridDim.x * kPosThreads) >> 2) {
    const bool live = p < np;
    const int2 ent = live ? ws.pos_list[(size_t)n * g.A + p] : make_int2(0, 0);
    pos_item<GRAD>(g, ws, A, n, live, ent.x, ent.y, p, side, s_acc, avg2);
  }
  if (GRAD) return;
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
// (lane = positive-in-batch x side) with the body of pos_kernel<false>.  The last block
// publishes num_pos and both avg factors.
constexpr int kAssignPer = 4;   // anchors per thread: the whole grid is one wave

__global__ void __launch_bounds__(256) assign_prepass_kernel(Geo g, Workspace ws, PosArgs A,
                                                             const int32_t* __restrict__ pad_hw,
                                                             int32_t* __restrict__ gt_inds,
                                                             int32_t* __restrict__ num_pos) {
  const int n = blockIdx.y;
  const int lane = threadIdx.x & 31;
  __shared__ double s_acc[2 * kLevels + 1];
  if (threadIdx.x < 2 * kLevels + 1) s_acc[threadIdx.x] = 0.0;
  __syncthreads();
  // one batch of independent loads: the argmax-table entries of this thread's anchors
  const int a0 = blockIdx.x * (256 * kAssignPer) + threadIdx.x;
  unsigned long long key[kAssignPer];
#pragma unroll
  for (int i = 0; i < kAssignPer; ++i) {
    const int a = a0 + i * 256;
    key[i] = a < g.A ? ws.atss_key[(size_t)n * g.A + a] : 0ull;
  }
  const int pad_h = pad_hw[n * 2], pad_w = pad_hw[n * 2 + 1], first_gt = A.gt_offsets[n];
#pragma unroll
  for (int i = 0; i < kAssignPer; ++i) {
    const int a = a0 + i * 256;
    const int gidx = a < g.A ? atss_decode_key(g, ws, pad_h, pad_w, first_gt, gt_inds, n, a, key[i]) : -1;
    unsigned todo = __ballot_sync(0xffffffffu, gidx >= 0);
    while (todo) {   // warp-uniform
      // the (lane >> 2)-th positive still to do, if there is one
      unsigned m = todo;
      for (int k = 0; k < (lane >> 2); ++k) m &= m - 1;
      const bool live = m != 0;
      const int src = live ? __ffs(m) - 1 : 0;
      const int pa = __shfl_sync(0xffffffffu, a, src);
      const int pg = __shfl_sync(0xffffffffu, gidx, src);
      pos_item<false>(g, ws, A, n, live, live ? pa : 0, live ? pg : 0, 0, lane & 3, s_acc, 1.0f);
      for (int k = 0; k < 8 && todo; ++k) todo &= todo - 1;
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
      ws.counters[3] = 0u;
    }
  }
}

// ----------------------------------------------------------------------------- class sweep
// QFL over the new-class channels.  grid (tile, image, part): a part is a group of kSweepCh
// channels; part numbers start behind the (ori + kSweepCh - 1) / kSweepCh old-class groups.
constexpr int kSweepCh = 8;

template <bool VEC>
__device__ __forceinline__ void cls_tile(const Geo& g, const Workspace& ws, const LossArgs& A, int n, int l,
                                         int hw0, int part, float& out_loss) {
  const int HW = g.hw[l];
  const Quad<VEC> q(hw0, HW);
  const size_t abase = (size_t)n * g.A + g.start[l];
  const int parts_old = (g.ori + kSweepCh - 1) / kSweepCh;
  const float* scls = A.s_cls.p[l] + (size_t)n * g.C * HW;
  float* gcls = A.g_cls.p[l] + (size_t)n * g.C * HW;
  // QFL over new-class channels of every anchor (:260-261,317-320)
  const int c0 = (part - parts_old) * kSweepCh, c1 = min(c0 + kSweepCh, g.cn);
  int label[4];
  float score[4], lw[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int gi = q.ok[k] ? A.gt_inds[abase + q.hw[k]] : -1;
    lw[k] = gi >= 0 ? 1.0f : 0.0f;                                  // label_weights, gfl_head.py:650-655,663
    label[k] = -1;
    score[k] = 0.f;
    if (gi > 0) {
      const long long lab = A.gt_labels[A.gt_offsets[n] + gi - 1];
      if (lab >= c0 && lab < c1) {
        label[k] = (int)lab;
        score[k] = ws.pos_score[abase + q.hw[k]];
      }
    }
  }
  const float inv_avg1 = 1.0f / (float)((double)A.avg[0] + (double)kEps32);            // losses/utils.py:60-61
  const float scale_cls = upstream_of(A.upstream, acc_cls(l)) * g.w_cls * inv_avg1;
  float loss_cls = 0.f;
  float gs[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) gs[k] = lw[k] * scale_cls;
  // every element as a negative first: the hot loop stays branch-free ...
#pragma unroll 4
  for (int c = c0; c < c1; ++c) {
    float x[4], gr[4];
    q.load(scls + (size_t)(g.ori + c) * HW, x, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const QflTerm t = qfl_neg(x[k]);
      loss_cls = fmaf(lw[k], t.loss, loss_cls);
      gr[k] = gs[k] * t.grad;
    }
    q.store(gcls + (size_t)(g.ori + c) * HW, gr);
  }
  // ... then the label channel of the (rare) positives is redone with its soft target
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (label[k] < 0) continue;
    const size_t off = (size_t)(g.ori + label[k]) * HW + q.hw[k];
    const float x = scls[off];
    const QflTerm tp = qfl_pos(x, score[k]), tn = qfl_neg(x);
    loss_cls += lw[k] * (tp.loss - tn.loss);
    gcls[off] = gs[k] * tp.grad;
  }
  out_loss += loss_cls;
}

__global__ void __launch_bounds__(kTileThreads) cls_sweep_kernel(Geo g, Workspace ws, LossArgs A, int part_offset) {
  if (A.skip_flag && *A.skip_flag == 0u) return;
  const int n = blockIdx.y;
  const int tile = blockIdx.x;
  const int part = blockIdx.z + part_offset;
  const int l = level_of_tile(g, tile);
  const int hw0 = (tile - g.tile_start[l]) * kTile;
  float out = 0.f;
  if (g.vec[l])
    cls_tile<true>(g, ws, A, n, l, hw0, part, out);
  else
    cls_tile<false>(g, ws, A, n, l, hw0, part, out);
  __shared__ float red[kTileThreads / 32];
  out = warp_sum(out);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = out;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kTileThreads / 32; ++w) s += (double)red[w];
    if (s != 0.0) atomicAdd(ws.loss_acc + acc_cls(l), s);
  }
}

// ----------------------------------------------------------------------------- box distillation
// DFL-distribution distillation (KL at temperature T between student and teacher box
// distributions, weighted by the student's max old-class score; :204-221, kd_loss.py:12-37).
// Only the NMS survivors carry it (a few dozen anchors per image at iou_threshold 0.005), so it
// is computed where their gradient rows are written instead of for every ERS candidate.
//
// One side (17 bins) of one anchor: returns sum_j p_t (log p_t - log p_s) and fills
// row[j] = w * (p_s - p_t); the constant factor (upstream, dist_loss_weight, loss weight, T) is
// applied by the caller.
__device__ __forceinline__ float kd_side(const Geo& g, const LossArgs& A, int n, int l, int hw, int side, float w,
                                         float* row) {
  const int HW = g.hw[l];
  const float inv_T = 1.0f / g.T;
  const size_t off = ((size_t)n * kBoxCh + side * kBins) * HW + hw;
  const float* sp = A.s_box.p[l] + off;
  const float* tp = A.t_box.p[l] + off;
  float zs[kBins], zt[kBins];
#pragma unroll
  for (int j = 0; j < kBins; ++j) {
    zs[j] = __ldg(sp + (size_t)j * HW) * inv_T;
    zt[j] = __ldg(tp + (size_t)j * HW) * inv_T;
  }
  float ms = zs[0], mt = zt[0];
#pragma unroll
  for (int j = 1; j < kBins; ++j) { ms = fmaxf(ms, zs[j]); mt = fmaxf(mt, zt[j]); }
  float ss = 0.f, st = 0.f;
#pragma unroll
  for (int j = 0; j < kBins; ++j) {
    zs[j] -= ms;
    zt[j] -= mt;
    ss += expf(zs[j]);
    st += expf(zt[j]);
  }
  const float lss = logf(ss), lst = logf(st);
  float kl = 0.f;
#pragma unroll
  for (int j = 0; j < kBins; ++j) {
    const float lps = zs[j] - lss, lpt = zt[j] - lst;
    const float ps = expf(lps), pt = expf(lpt);
    if (pt > 0.f) kl += pt * (lpt - lps);
    row[j] = w * (ps - pt);
  }
  return kl;
}

// this thread's share (channels first, first + step, ...) of max_c over the old-class logits
__device__ __forceinline__ float kd_weight_part(const Geo& g, const LossArgs& A, int n, int l, int hw, int first,
                                                int step) {
  const int HW = g.hw[l];
  const float* cplane = A.s_cls.p[l] + (size_t)n * g.C * HW + hw;
  float mx = -INFINITY;
  for (int c = first; c < g.ori; c += step) mx = fmaxf(mx, __ldg(cplane + (size_t)c * HW));
  return mx;
}

// ----------------------------------------------------------------------------- zero fill
// Most of the gradient is structurally zero: the old-class channels off the ERS rows and the
// box channels of every anchor that is neither a positive nor an ERS box candidate.  One
// launch clears both families (10 strided regions), with streaming 16 B stores; it depends on
// nothing, so erd_step_prepare runs it at the very start of the step beside the ERS scan.
// The sparse non-zero rows are written over it afterwards (cls_kd / pos_kernel<true> / box_kd).
struct ZeroArgs {
  float* base[2 * kLevels];
  long long pitch[2 * kLevels];    // floats between consecutive rows
  long long width[2 * kLevels];    // floats to clear per row
  int rows[2 * kLevels];
  int chunk0[2 * kLevels + 1];     // prefix of kZeroChunk-sized chunks
  int vec[2 * kLevels];
  const unsigned int* skip_flag;   // non-NULL: no-op when *skip_flag == 0 (see LossArgs)
};
constexpr int kZeroThreads = 256;
constexpr int kZeroChunk = kZeroThreads * 16;   // floats per chunk

__global__ void __launch_bounds__(kZeroThreads) zero_fill_kernel(ZeroArgs z) {
  if (z.skip_flag && *z.skip_flag == 0u) return;
  const int total = z.chunk0[2 * kLevels];
  for (int c = blockIdx.x; c < total; c += gridDim.x) {
    int reg = 0;
    while (c >= z.chunk0[reg + 1]) ++reg;
    const long long per_row = (z.width[reg] + kZeroChunk - 1) / kZeroChunk;
    const long long local = c - z.chunk0[reg];
    const long long row = local / per_row;
    const long long off = (local - row * per_row) * kZeroChunk;
    float* p = z.base[reg] + row * z.pitch[reg] + off;
    const long long len = min((long long)kZeroChunk, z.width[reg] - off);
    if (z.vec[reg]) {
      for (long long i = (long long)threadIdx.x * 4; i < len; i += kZeroThreads * 4)
        __stcs(reinterpret_cast<float4*>(p + i), make_float4(0.f, 0.f, 0.f, 0.f));
    } else {
      for (long long i = threadIdx.x; i < len; i += kZeroThreads) __stcs(p + i, 0.f);
    }
  }
}

cudaError_t launch_zero_fill(const Geo& g, const MPtr5& g_cls, const MPtr5& g_box, const unsigned int* skip_flag,
                             cudaStream_t st) {
  ZeroArgs z;
  z.skip_flag = skip_flag;
  int chunks = 0;
  for (int i = 0; i < 2 * kLevels; ++i) {
    const int l = i % kLevels;
    const bool cls = i < kLevels;
    z.base[i] = cls ? g_cls.p[l] : g_box.p[l];
    z.pitch[i] = (long long)(cls ? g.C : kBoxCh) * g.hw[l];
    z.width[i] = (long long)(cls ? g.ori : kBoxCh) * g.hw[l];
    z.rows[i] = g.n_img;
    z.vec[i] = g.vec[l];
    if (!cls) {   // contiguous over the images: one long row
      z.width[i] *= g.n_img;
      z.rows[i] = 1;
    }
    z.chunk0[i] = chunks;
    chunks += (int)((z.width[i] + kZeroChunk - 1) / kZeroChunk) * z.rows[i];
  }
  z.chunk0[2 * kLevels] = chunks;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  static int per_sm = 0;
  if (!per_sm) {
    const char* v = getenv("ERD_ZERO_CTAS");
    per_sm = v ? atoi(v) : 1;
    if (per_sm < 1 || per_sm > 8) per_sm = 1;
  }
  const int grid = chunks < per_sm * sms ? chunks : per_sm * sms;
  if (grid > 0) ERD_LAUNCH(kKZero, st, (zero_fill_kernel<<<grid, kZeroThreads, 0, st>>>(z)));
  return cudaGetLastError();
}

// ----------------------------------------------------------------------------- class-response distillation
// 2 (x_s - x_t) / (K ori) on the old-class channels of the ERS rows (gfl_head_increment_erd.py:
// 181-186,324-332), list driven: eight threads per selected anchor, each a strided eighth of
// the channels, all gathers in flight at once.  Overwrites the zero fill on those rows.
constexpr int kClsKdThreads = 256;

__global__ void __launch_bounds__(kClsKdThreads) cls_kd_kernel(Geo g, Workspace ws, LossArgs A) {
  if (A.skip_flag && *A.skip_flag == 0u) return;
  const int n = blockIdx.y;
  const int sub = threadIdx.x & 7;
  const int K = A.cls_count[n];
  const float kc = (float)K * (float)g.ori;
  const float scale_dc = upstream_of(A.upstream, acc_dcls(n)) * A.dlw * 2.0f / kc;
  float sq = 0.f;
  for (int r = (blockIdx.x * kClsKdThreads + threadIdx.x) >> 3; r < K; r += (gridDim.x * kClsKdThreads) >> 3) {
    const int a = A.cls_inds[(size_t)n * g.sel_cap + r];
    const int l = level_of_anchor(g, a);
    const int HW = g.hw[l];
    const int hw = a - g.start[l];
    const float* sp = A.s_cls.p[l] + (size_t)n * g.C * HW + hw;
    const float* tp = A.t_cls.p[l] + (size_t)n * g.ori * HW + hw;
    float* gp = A.g_cls.p[l] + (size_t)n * g.C * HW + hw;
#pragma unroll 4
    for (int c = sub; c < g.ori; c += 8) {
      const float df = __ldg(sp + (size_t)c * HW) - __ldg(tp + (size_t)c * HW);
      sq = fmaf(df, df, sq);
      gp[(size_t)c * HW] = scale_dc * df;
    }
  }
  __shared__ float red[kClsKdThreads / 32];
  sq = warp_sum(sq);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s2 = 0.0;
    for (int w = 0; w < kClsKdThreads / 32; ++w) s2 += (double)red[w];
    if (s2 != 0.0) atomicAdd(ws.loss_acc + acc_dcls(n), s2);
  }
}

// Box-logit gradients of the ERS box candidates, list driven: four threads per candidate, one
// per box side, all loads in flight at once.  The distillation only counts for the candidates
// the teacher NMS keeps, which is not known yet: this kernel runs BESIDE the NMS and writes
// every candidate as if kept -- (positive's row, if it is one) + the KL gradient, and the
// candidate's weighted KL into ws.kd_loss -- and box_fix_kernel takes the suppressed ones back
// afterwards.  (Its ~64 B-per-element NCHW gathers are the expensive part and overlap the
// latency-bound NMS chain this way, whatever fraction the NMS ends up keeping.)
constexpr int kLateThreads = 256;

__device__ __forceinline__ void finalize_one(const Geo& g, const Workspace& ws, const LossArgs& A, int i);

__global__ void __launch_bounds__(kLateThreads) box_kd_kernel(Geo g, Workspace ws, LossArgs A) {
  if (A.skip_flag && *A.skip_flag == 0u) return;
  const int n = blockIdx.y;
  const int side = threadIdx.x & 3;
  const int K = A.box_count[n];
  const float kT = g.T;
  const float scale = upstream_of(A.upstream, acc_dbox(g, n)) * A.dlw * g.w_ld / 4.0f * (kT * kT / (float)kBins) / kT;
  // whole warps stay together (8 candidates per warp) so the quad shuffles are convergent
  for (int r = (blockIdx.x * kLateThreads + threadIdx.x) >> 2; r < ((K + 7) & ~7);
       r += (gridDim.x * kLateThreads) >> 2) {
    const bool on = r < K;
    const int a = on ? A.box_inds[(size_t)n * g.sel_cap + r] : 0;
    const int l = level_of_anchor(g, a);
    const int HW = g.hw[l];
    const int hw = a - g.start[l];
    const size_t ga = (size_t)n * g.A + a;
    const bool pos = on && A.gt_inds[ga] > 0;
    const int slot = pos ? ws.pos_slot[ga] : 0;
    float mx = on ? kd_weight_part(g, A, n, l, hw, side, 4) : 0.f;
    mx = quad_max(mx);
    const float w = sigmoid_ref(mx);                                               // :217-218
    float row[kBins];
    float kl = 0.f;
    if (on) kl = kd_side(g, A, n, l, hw, side, w, row);
    kl += __shfl_xor_sync(0xffffffffu, kl, 1);
    kl += __shfl_xor_sync(0xffffffffu, kl, 2);
    if (!on) continue;
    if (side == 0) ws.kd_loss[(size_t)n * g.sel_cap + r] = w * (kl / (float)kBins * (kT * kT));   // .mean(1) * T*T
    const float* prow = pos ? ws.pos_rows + ((size_t)n * g.pos_cap + slot) * kBoxCh + side * kBins : nullptr;
    float* gp = A.g_box.p[l] + ((size_t)n * kBoxCh + side * kBins) * HW + hw;
#pragma unroll
    for (int j = 0; j < kBins; ++j) gp[(size_t)j * HW] = fmaf(scale, row[j], prow ? prow[j] : 0.f);
  }
}

// After the NMS: sum the weighted KL of the survivors (flag bit 2 set by the resolve pass) and
// rewrite the gradient of the suppressed candidates without the distillation term.
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
      if (side == 0) kd += ws.kd_loss[(size_t)n * g.sel_cap + r];
      continue;
    }
    const int l = level_of_anchor(g, a);
    const int HW = g.hw[l];
    const int hw = a - g.start[l];
    const float* prow = nullptr;
    if (A.gt_inds[ga] > 0) prow = ws.pos_rows + ((size_t)n * g.pos_cap + ws.pos_slot[ga]) * kBoxCh + side * kBins;
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
  for (int l = 0; l < kLevels; ++l) a.g_box.p[l] = nullptr;
  a.gt_boxes = gt_boxes;
  a.gt_labels = gt_labels;
  a.gt_offsets = gt_offsets;
  a.gt_inds = gt_inds;
  a.num_pos = num_pos;
  a.avg = avg;
  a.upstream = nullptr;
  a.skip_flag = nullptr;
  ERD_LAUNCH(kKAvg, st, (pos_kernel<false><<<dim3(pos_grid_x(g), g.n_img), kPosThreads, 0, st>>>(g, ws, a)));
  return cudaGetLastError();
}

cudaError_t launch_assign_avg(const Geo& g, const Workspace& ws, const Ptr5& s_cls, const Ptr5& s_box,
                              const float* gt_boxes, const int64_t* gt_labels, const int32_t* gt_offsets,
                              const int32_t* pad_hw, int32_t* gt_inds, int32_t* num_pos, float* avg, cudaStream_t st) {
  cudaError_t e = launch_atss_candidates(g, ws, gt_boxes, gt_offsets, pad_hw, st);
  if (e != cudaSuccess) return e;
  PosArgs a;
  a.s_cls = s_cls;
  a.s_box = s_box;
  for (int l = 0; l < kLevels; ++l) a.g_box.p[l] = nullptr;
  a.gt_boxes = gt_boxes;
  a.gt_labels = gt_labels;
  a.gt_offsets = gt_offsets;
  a.gt_inds = gt_inds;
  a.num_pos = num_pos;
  a.avg = avg;
  a.upstream = nullptr;
  a.skip_flag = nullptr;
  ERD_LAUNCH(kKAvg, st,
             (assign_prepass_kernel<<<dim3((g.A + 256 * kAssignPer - 1) / (256 * kAssignPer), g.n_img), 256, 0, st>>>(
                 g, ws, a, pad_hw, gt_inds, num_pos)));
  return cudaGetLastError();
}

cudaError_t launch_loss(const Geo& g, const Workspace& ws, const LossArgs& a, cudaStream_t st, const LossStreams* ls) {
  const int total = 3 * kLevels + 2 * g.n_img;
  cudaError_t e = cudaSuccess;   // accumulators are left clean by the previous finalize (erd_workspace_init once)
  if (a.skip_flag) ERD_LAUNCH(kKUpCheck, st, (upstream_check_kernel<<<1, 128, 0, st>>>(ws, a.upstream, total)));
  PosArgs p;
  p.s_cls = a.s_cls;
  p.s_box = a.s_box;
  p.g_box = a.g_box;
  p.gt_boxes = a.gt_boxes;
  p.gt_labels = a.gt_labels;
  p.gt_offsets = a.gt_offsets;
  p.gt_inds = a.gt_inds;
  p.num_pos = a.num_pos;
  p.avg = const_cast<float*>(a.avg);
  p.upstream = a.upstream;
  p.skip_flag = a.skip_flag;
  const dim3 late_grid((g.sel_cap * 4 + kLateThreads - 1) / kLateThreads, g.n_img);
  const dim3 cls_kd_grid((g.sel_cap * 8 + kClsKdThreads - 1) / kClsKdThreads, g.n_img);
  const int parts_old = (g.ori + kSweepCh - 1) / kSweepCh, parts_new = (g.cn + kSweepCh - 1) / kSweepCh;
  const dim3 tile_grid_new(g.tile_start[kLevels], g.n_img, parts_new);
  // Schedule.  The only dense pass is the QFL sweep over the new-class channels (caller's
  // stream).  Everything else is a zero fill (done by erd_step_prepare when it was given the
  // gradient pointers, else here) overwritten by sparse, list-driven launches:
  //   lo: [zero fill] -> wait(select) -> class-response rows of the ERS set
  //   hi: positives' box rows -> wait(select) -> box candidates' rows incl. distillation
  //       -> wait(NMS) -> take-back of the suppressed candidates
  // whose last block also turns the accumulators into the loss vector.
  cudaStream_t hi = ls ? ls->late : st, lo = ls ? ls->early : st;
  if (ls) {
    e = cudaEventRecord(ls->fork, st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(hi, ls->fork, 0);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(lo, ls->fork, 0);
    if (e != cudaSuccess) return e;
  }
  if (ls && ls->cleared) {
    e = cudaStreamWaitEvent(hi, ls->cleared, 0);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(lo, ls->cleared, 0);
    if (e != cudaSuccess) return e;
  } else {
    e = launch_zero_fill(g, a.g_cls, a.g_box, a.skip_flag, lo);
    if (e != cudaSuccess) return e;
    if (ls) {
      e = cudaEventRecord(ls->pos_done, lo);   // (event reused: "zero fill done")
      if (e == cudaSuccess) e = cudaStreamWaitEvent(hi, ls->pos_done, 0);
      if (e != cudaSuccess) return e;
    }
  }
  ERD_LAUNCH(kKPosGrad, hi, (pos_kernel<true><<<dim3(pos_grid_x(g), g.n_img), kPosThreads, 0, hi>>>(g, ws, p)));
  if (ls && ls->sel_ready) {
    e = cudaStreamWaitEvent(lo, ls->sel_ready, 0);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(hi, ls->sel_ready, 0);
    if (e != cudaSuccess) return e;
  }
  ERD_LAUNCH(kKClsOld, lo, (cls_kd_kernel<<<cls_kd_grid, kClsKdThreads, 0, lo>>>(g, ws, a)));
  ERD_LAUNCH(kKBoxKd, hi, (box_kd_kernel<<<late_grid, kLateThreads, 0, hi>>>(g, ws, a)));
  ERD_LAUNCH(kKLossMain, st, (cls_sweep_kernel<<<tile_grid_new, kTileThreads, 0, st>>>(g, ws, a, parts_old)));
  if (ls) {   // the take-back pass also finalizes: order every other launch of the step before it
    e = cudaEventRecord(ls->early_done, lo);
    if (e == cudaSuccess) e = cudaEventRecord(ls->main_done, st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(hi, ls->early_done, 0);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(hi, ls->main_done, 0);
    if (e == cudaSuccess && ls->nms_done) e = cudaStreamWaitEvent(hi, ls->nms_done, 0);
    if (e != cudaSuccess) return e;
  }
  ERD_LAUNCH(kKBoxSweep, hi, (box_fix_kernel<<<late_grid, kLateThreads, 0, hi>>>(g, ws, a)));
  if (ls) {
    e = cudaEventRecord(ls->late_done, hi);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ls->late_done, 0);
    if (e != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

}  // namespace erd
