// Fused forward + backward of the GT losses (QFL / GIoU / DFL) and the two distillation
// losses, writing dense NCHW gradients in one sweep over the student head outputs.
// Reference: GFLHeadIncrementERD.loss_by_feat_single / distill_loss_by_image_single /
// loss_by_feat (dense_heads/gfl_head_increment_erd.py:142-454), quality_focal_loss and
// distribution_focal_loss (losses/gfocal_loss.py:12-53,143-165), giou_loss
// (losses/iou_loss.py:110-126) over bbox_overlaps (structures/bbox/bbox_overlaps.py:151-199),
// knowledge_distillation_kl_div_loss (losses/kd_loss.py:12-37), weight_reduce_loss
// (losses/utils.py:30-65).  Closed-form gradients: SURVEY.md Appendix A.
//
// Kernels
//   pos_kernel<false>  (prepare phase) 4 threads per positive anchor: weight, softmax-integral
//                      decode, IoU score, GIoU/DFL loss sums, and the two avg factors.
//   pos_kernel<true>   box-logit gradient rows of the positives (needs the reduced avg factor),
//                      written to a compact row buffer.
//   cls_sweep_kernel   streaming sweep over the class logits: QFL on the new-class channels of
//                      every anchor, class-response L2 on the old-class channels (zero off
//                      the ERS rows), gradients written once, densely.
//   kd_rows_kernel     (prepare phase, beside the NMS) DFL-distribution KL rows of every ERS
//                      box candidate.
//   box_sweep_kernel   dense write of the box-logit gradients: zero, the compact rows of the
//                      positives, and the distillation rows of the NMS survivors.
//   finalize_kernel    accumulators -> the reference's loss values.
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
  const float e = __expf(-fabsf(x));
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
    // Dependent memory round trips are what this kernel costs (it runs beside DRAM-saturating
    // sweeps), so everything past the list entry is issued as one independent batch of loads.
    const bool live = p < np;
    const int2 ent = live ? ws.pos_list[(size_t)n * g.A + p] : make_int2(0, 0);
    const int a = ent.x, gidx = ent.y;
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
      float* row = ws.pos_rows + ((size_t)n * g.pos_cap + p) * kBoxCh + side * kBins;
#pragma unroll
      for (int j = 0; j < kBins; ++j) {
        const float pj = z[j] * inv;
        float gr = cb * pj * ((float)j - dmine);
        gr += cd * (wl * (pj - (j == yl ? 1.f : 0.f)) + wr * (pj - (j == yl + 1 ? 1.f : 0.f)));
        row[j] = gr;
      }
    } else if (GRAD && live) {   // assigned to a GT whose label lies outside the new-class range: no box loss
      float* row = ws.pos_rows + ((size_t)n * g.pos_cap + p) * kBoxCh + side * kBins;
#pragma unroll
      for (int j = 0; j < kBins; ++j) row[j] = 0.f;
    }
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

// ----------------------------------------------------------------------------- class sweep
// grid (tile, image, part): a part is a group of kSweepCh class channels that lies entirely
// in the old-class range [0, ori) or in the new-class range [ori, C).
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
  if (part < parts_old) {
    // classification-response distillation: 2 (x_s - x_t) / (K ori) on the ERS rows, zero elsewhere
    // (gfl_head_increment_erd.py:181-186,324-332).  A thread whose four anchors are all
    // unselected issues no loads at all.
    const int c0 = part * kSweepCh, c1 = min(c0 + kSweepCh, g.ori);
    float sel[4];
    bool any = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      sel[k] = (q.ok[k] && (A.sel_flags[abase + q.hw[k]] & 1)) ? 1.0f : 0.0f;   // gfl_increment_erd.py:149-151
      any |= sel[k] != 0.f;
    }
    if (!any) {
      for (int c = c0; c < c1; ++c) q.store_zero(gcls + (size_t)c * HW);
      return;
    }
    const float kc = (float)A.cls_count[n] * (float)g.ori;
    const float scale_dc = upstream_of(A.upstream, acc_dcls(n)) * A.dlw * 2.0f / kc;
    const float* tcls = A.t_cls.p[l] + (size_t)n * g.ori * HW;
    float sq = 0.f;
#pragma unroll 4
    for (int c = c0; c < c1; ++c) {
      float xs[4], xt[4], gr[4];
      q.load(scls + (size_t)c * HW, xs, 0.f);
      q.load(tcls + (size_t)c * HW, xt, 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float df = sel[k] * (xs[k] - xt[k]);
        sq = fmaf(df, df, sq);
        gr[k] = scale_dc * df;
      }
      q.store(gcls + (size_t)c * HW, gr);
    }
    out_loss += sq;
    return;
  }
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
#pragma unroll 4
  for (int c = c0; c < c1; ++c) {
    float x[4], gr[4];
    q.load(scls + (size_t)(g.ori + c) * HW, x, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const QflTerm t = (c == label[k]) ? qfl_pos(x[k], score[k]) : qfl_neg(x[k]);
      loss_cls += lw[k] * t.loss;
      gr[k] = lw[k] * scale_cls * t.grad;
    }
    q.store(gcls + (size_t)(g.ori + c) * HW, gr);
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
    const bool old_part = part < (g.ori + kSweepCh - 1) / kSweepCh;
    if (s != 0.0) atomicAdd(ws.loss_acc + (old_part ? acc_dcls(n) : acc_cls(l)), s);
  }
}

// ----------------------------------------------------------------------------- box distillation rows
// DFL-distribution distillation (KL at temperature T between student and teacher box
// distributions, weighted by the student's max old-class score; :204-221, kd_loss.py:12-37)
// for EVERY ERS box candidate, four threads per candidate (one per side), all loads in flight
// at once.  It depends only on the selection, so it runs beside the NMS; the box sweep later
// merges the rows of the candidates the NMS kept.  Rows hold w * (p_s - p_t); the constant
// factor (upstream, dist_loss_weight, loss weight, T) is applied at merge time.
constexpr int kKdThreads = 256;

__global__ void __launch_bounds__(kKdThreads) kd_rows_kernel(Geo g, Workspace ws, Ptr5 s_cls, Ptr5 s_box, Ptr5 t_box,
                                                             const int32_t* __restrict__ box_inds,
                                                             const int32_t* __restrict__ box_count) {
  const int n = blockIdx.y;
  const int side = threadIdx.x & 3;
  const int K = box_count[n];
  const float inv_T = 1.0f / g.T;
  for (int r = (blockIdx.x * kKdThreads + threadIdx.x) >> 2; r < ((K + 7) & ~7); r += (gridDim.x * kKdThreads) >> 2) {
    const bool on = r < K;
    const int a = on ? box_inds[(size_t)n * g.sel_cap + r] : 0;
    const int l = level_of_anchor(g, a);
    const int HW = g.hw[l];
    const int hw = a - g.start[l];
    float mx = -INFINITY;
    if (on) {
      const float* cplane = s_cls.p[l] + (size_t)n * g.C * HW + hw;
      for (int c = side; c < g.ori; c += 4) mx = fmaxf(mx, __ldg(cplane + (size_t)c * HW));
    }
    mx = quad_max(mx);
    const float w = sigmoid_ref(mx);                                               // :217-218
    const size_t off = ((size_t)n * kBoxCh + side * kBins) * HW + hw;
    const float* sp = s_box.p[l] + off;
    const float* tp = t_box.p[l] + off;
    float zs[kBins], zt[kBins];
#pragma unroll
    for (int j = 0; j < kBins; ++j) {
      zs[j] = on ? __ldg(sp + (size_t)j * HW) * inv_T : 0.f;
      zt[j] = on ? __ldg(tp + (size_t)j * HW) * inv_T : 0.f;
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
    float* row = ws.kd_rows + ((size_t)n * g.sel_cap + r) * kBoxCh + side * kBins;
#pragma unroll
    for (int j = 0; j < kBins; ++j) {
      const float lps = zs[j] - lss, lpt = zt[j] - lst;
      const float ps = expf(lps), pt = expf(lpt);
      if (pt > 0.f) kl += pt * (lpt - lps);
      if (on) row[j] = w * (ps - pt);
    }
    kl += __shfl_xor_sync(0xffffffffu, kl, 1);
    kl += __shfl_xor_sync(0xffffffffu, kl, 2);
    if (on && side == 0) ws.kd_loss[(size_t)n * g.sel_cap + r] = w * (kl / (float)kBins * (g.T * g.T));   // .mean(1) * T*T
  }
}

// ----------------------------------------------------------------------------- box sweep
// grid (tile, image, side).  Every box-logit gradient element is written exactly once, densely:
// zero, plus the compact row of a positive (pos_kernel<true>), plus the distillation row of an
// NMS survivor (kd_rows_kernel, marked by the NMS resolve pass).
//
// MODE 0: everything in one launch.  MODE 1 ("early", runs beside the NMS): only the groups
// that hold no ERS box candidate, i.e. whose result cannot depend on the NMS; the remaining
// groups are written after the NMS by box_late_kernel, one warp per candidate.  On 16 B aligned
// levels a group is 8 consecutive anchors = one 32 B sector per channel, so the two launches
// never share a sector; on the other levels (a few % of the anchors) a group is one anchor.
template <bool VEC, int MODE>
__device__ __forceinline__ void box_tile(const Geo& g, const Workspace& ws, const LossArgs& A, int n, int l,
                                         int hw0, int side, float& kd_loss) {
  const int HW = g.hw[l];
  Quad<VEC> q(hw0, HW);
  const size_t abase = (size_t)n * g.A + g.start[l];
  const float* prow[4];
  const float* krow[4];
  bool any = false, cand = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    prow[k] = nullptr;
    krow[k] = nullptr;
    if (!q.ok[k]) continue;
    const size_t a = abase + q.hw[k];
    const unsigned flags = A.sel_flags[a];
    cand |= (flags & 2) != 0;
    if (!VEC && MODE == 1 && (flags & 2)) { q.ok[k] = false; continue; }   // scalar levels: groups are single anchors
    if (MODE != 1 && (flags & 4)) {
      const size_t slot = (size_t)n * g.sel_cap + ws.kd_slot[a];
      krow[k] = ws.kd_rows + slot * kBoxCh + side * kBins;
      if (side == 0) kd_loss += ws.kd_loss[slot];
    }
    if (A.gt_inds[a] > 0)
      prow[k] = ws.pos_rows + ((size_t)n * g.pos_cap + ws.pos_slot[a]) * kBoxCh + side * kBins;
    any |= prow[k] != nullptr || krow[k] != nullptr;
  }
  if (VEC && MODE != 0) {
    const int neighbour = __shfl_xor_sync(0xffffffffu, cand ? 1 : 0, 1);   // every lane takes part
    const bool group = cand || neighbour != 0;
    if ((MODE == 1) == group) return;
  }
  float* gbox = A.g_box.p[l] + ((size_t)n * kBoxCh + side * kBins) * HW;
  if (!any) {
#pragma unroll
    for (int j = 0; j < kBins; ++j) q.store_zero(gbox + (size_t)j * HW);
    return;
  }
  const float kT = g.T;
  const float scale = upstream_of(A.upstream, acc_dbox(g, n)) * A.dlw * g.w_ld / 4.0f * (kT * kT / (float)kBins) / kT;
  for (int j = 0; j < kBins; ++j) {
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = 0.f;
      if (prow[k]) v[k] += prow[k][j];
      if (krow[k]) v[k] = fmaf(scale, krow[k][j], v[k]);
    }
    q.store(gbox + (size_t)j * HW, v);
  }
}

template <int MODE>
__global__ void __launch_bounds__(kTileThreads) box_sweep_kernel(Geo g, Workspace ws, LossArgs A) {
  if (A.skip_flag && *A.skip_flag == 0u) return;
  const int n = blockIdx.y;
  const int tile = blockIdx.x;
  const int l = level_of_tile(g, tile);
  const int hw0 = (tile - g.tile_start[l]) * kTile;
  float kd = 0.f;
  if (g.vec[l])
    box_tile<true, MODE>(g, ws, A, n, l, hw0, blockIdx.z, kd);
  else
    box_tile<false, MODE>(g, ws, A, n, l, hw0, blockIdx.z, kd);
  if (MODE == 1) return;
  __shared__ float red[kTileThreads / 32];
  kd = warp_sum(kd);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = kd;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kTileThreads / 32; ++w) s += (double)red[w];
    if (s != 0.0) atomicAdd(ws.loss_acc + acc_dbox(g, n), s);
  }
}

// Late box gradients, list driven: one warp per ERS box candidate writes the whole group the
// candidate lives in (lane = anchor-in-group x side), merging positives' rows and, where the
// NMS kept the anchor, its distillation row.  A group shared by several candidates is written
// by the warp of its first candidate only.
constexpr int kLateThreads = 256;

__device__ __forceinline__ void finalize_one(const Geo& g, const Workspace& ws, const LossArgs& A, int i);

__global__ void __launch_bounds__(kLateThreads) box_late_kernel(Geo g, Workspace ws, LossArgs A,
                                                                const int32_t* __restrict__ box_count) {
  if (A.skip_flag && *A.skip_flag == 0u) return;
  const int n = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int K = box_count[n];
  const float kT = g.T;
  const float scale = upstream_of(A.upstream, acc_dbox(g, n)) * A.dlw * g.w_ld / 4.0f * (kT * kT / (float)kBins) / kT;
  float kd = 0.f;
  for (int r = blockIdx.x * (kLateThreads / 32) + warp; r < K; r += gridDim.x * (kLateThreads / 32)) {
    const int a = A.box_inds[(size_t)n * g.sel_cap + r];
    const int l = level_of_anchor(g, a);
    const int HW = g.hw[l];
    const int hw = a - g.start[l];
    const size_t abase = (size_t)n * g.A + g.start[l];
    const bool vec = g.vec[l] != 0;
    const int g0 = vec ? (hw & ~7) : hw;
    const int k = vec ? (lane & 7) : 0;
    const int side = vec ? (lane >> 3) : lane;
    const int hwk = g0 + k;
    const bool live = hwk < HW && side < 4;
    const unsigned flags = live ? A.sel_flags[abase + hwk] : 0u;
    if (vec) {   // the first candidate of the group owns it
      const unsigned cands = __ballot_sync(0xffffffffu, (flags & 2) != 0) & 0xffu;
      if (g0 + __ffs(cands) - 1 != hw) continue;
    }
    if (!live) continue;
    const float* prow = nullptr;
    const float* krow = nullptr;
    if (A.gt_inds[abase + hwk] > 0)
      prow = ws.pos_rows + ((size_t)n * g.pos_cap + ws.pos_slot[abase + hwk]) * kBoxCh + side * kBins;
    if (flags & 4) {
      const size_t slot = (size_t)n * g.sel_cap + ws.kd_slot[abase + hwk];
      krow = ws.kd_rows + slot * kBoxCh + side * kBins;
      if (side == 0) kd += ws.kd_loss[slot];
    }
    float* gp = A.g_box.p[l] + ((size_t)n * kBoxCh + side * kBins) * HW + hwk;
#pragma unroll
    for (int j = 0; j < kBins; ++j) {
      float v = 0.f;
      if (prow) v = prow[j];
      if (krow) v = fmaf(scale, krow[j], v);
      gp[(size_t)j * HW] = v;
    }
  }
  __shared__ float red[kLateThreads / 32];
  kd = warp_sum(kd);
  if (lane == 0) red[warp] = kd;
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    double s2 = 0.0;
    for (int w = 0; w < kLateThreads / 32; ++w) s2 += (double)red[w];
    if (s2 != 0.0) atomicAdd(ws.loss_acc + acc_dbox(g, n), s2);
    // the last block to finish turns the accumulators into the loss vector (saves a launch at
    // the very end of the step's critical path)
    __threadfence();
    last = atomicAdd(ws.counters + 2, 1u) == gridDim.x * gridDim.y - 1;
  }
  __syncthreads();
  if (last) {
    __threadfence();
    finalize_one(g, ws, A, threadIdx.x);
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

__global__ void finalize_kernel(Geo g, Workspace ws, LossArgs A) {
  if (A.skip_flag && *A.skip_flag == 0u) return;
  finalize_one(g, ws, A, threadIdx.x);
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

cudaError_t launch_kd_rows(const Geo& g, const Workspace& ws, const Ptr5& s_cls, const Ptr5& s_box, const Ptr5& t_box,
                           const int32_t* box_inds, const int32_t* box_count, cudaStream_t st) {
  ERD_LAUNCH(kKKdRows, st,
             (kd_rows_kernel<<<dim3(16, g.n_img), kKdThreads, 0, st>>>(g, ws, s_cls, s_box, t_box, box_inds, box_count)));
  return cudaGetLastError();
}

// Sequence on the caller's stream `st`.  With helper streams (erd_step_prepare's context):
//   st    : QFL sweep (needs only assignment + avg factors) -> wait selection -> class-response
//           sweep -> wait early/NMS/KD -> late box kernel -> finalize
//   early : wait selection -> positives' rows -> box sectors that cannot depend on the NMS
// so the bandwidth-bound sweeps overlap the latency-bound ERS / NMS chain instead of queueing
// behind it.
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
  const dim3 box_grid(g.tile_start[kLevels], g.n_img, 4);
  const int parts_old = (g.ori + kSweepCh - 1) / kSweepCh, parts_new = (g.cn + kSweepCh - 1) / kSweepCh;
  const dim3 tile_grid_new(g.tile_start[kLevels], g.n_img, parts_new), tile_grid_old(g.tile_start[kLevels], g.n_img, parts_old);
  if (ls) {
    e = cudaEventRecord(ls->fork, st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ls->early, ls->fork, 0);
    if (e == cudaSuccess && ls->sel_ready) e = cudaStreamWaitEvent(ls->early, ls->sel_ready, 0);
    if (e != cudaSuccess) return e;
    ERD_LAUNCH(kKPosGrad, ls->early,
               (pos_kernel<true><<<dim3(pos_grid_x(g), g.n_img), kPosThreads, 0, ls->early>>>(g, ws, p)));
    ERD_LAUNCH(kKBoxEarly, ls->early, (box_sweep_kernel<1><<<box_grid, kTileThreads, 0, ls->early>>>(g, ws, a)));
    e = cudaEventRecord(ls->early_done, ls->early);
    if (e != cudaSuccess) return e;
  }
  ERD_LAUNCH(kKLossMain, st, (cls_sweep_kernel<<<tile_grid_new, kTileThreads, 0, st>>>(g, ws, a, parts_old)));
  if (ls && ls->sel_ready) {
    e = cudaStreamWaitEvent(st, ls->sel_ready, 0);
    if (e != cudaSuccess) return e;
  }
  ERD_LAUNCH(kKClsOld, st, (cls_sweep_kernel<<<tile_grid_old, kTileThreads, 0, st>>>(g, ws, a, 0)));
  if (ls) {
    e = cudaStreamWaitEvent(st, ls->early_done, 0);
    if (e == cudaSuccess && ls->nms_done) e = cudaStreamWaitEvent(st, ls->nms_done, 0);
    if (e == cudaSuccess && ls->kd_done) e = cudaStreamWaitEvent(st, ls->kd_done, 0);
    if (e != cudaSuccess) return e;
    ERD_LAUNCH(kKBoxSweep, st, (box_late_kernel<<<dim3(128, g.n_img), kLateThreads, 0, st>>>(g, ws, a, a.box_count)));
    return cudaGetLastError();   // box_late's last block wrote the loss vector
  } else {
    ERD_LAUNCH(kKPosGrad, st, (pos_kernel<true><<<dim3(pos_grid_x(g), g.n_img), kPosThreads, 0, st>>>(g, ws, p)));
    ERD_LAUNCH(kKBoxSweep, st, (box_sweep_kernel<0><<<box_grid, kTileThreads, 0, st>>>(g, ws, a)));
  }
  ERD_LAUNCH(kKFinalize, st, (finalize_kernel<<<1, ((total + 31) / 32) * 32, 0, st>>>(g, ws, a)));
  return cudaGetLastError();
}

}  // namespace erd
