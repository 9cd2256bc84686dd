// Fused forward + backward of the GT losses (QFL / GIoU / DFL) and the two distillation
// losses, writing dense NCHW gradients in one sweep over the student head outputs.
// Reference: GFLHeadIncrementERD.loss_by_feat_single / distill_loss_by_image_single /
// loss_by_feat (dense_heads/gfl_head_increment_erd.py:142-454), quality_focal_loss and
// distribution_focal_loss (losses/gfocal_loss.py:12-53,143-165), giou_loss
// (losses/iou_loss.py:110-126) over bbox_overlaps (structures/bbox/bbox_overlaps.py:151-199),
// knowledge_distillation_kl_div_loss (losses/kd_loss.py:12-37), weight_reduce_loss
// (losses/utils.py:30-65).  Closed-form gradients: SURVEY.md Appendix A.
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

// sigma and softplus from one exp: e = exp(-|x|)
__device__ __forceinline__ void sig_sp(float x, float& sig, float& sp) {
  const float e = __expf(-fabsf(x));
  const float r = __fdividef(1.0f, 1.0f + e);
  sig = x >= 0.f ? r : e * r;
  sp = fmaxf(x, 0.f) + log1pf(e);
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

struct PosOut {
  float score, w;
};

// Everything a positive anchor contributes: weight, decode, IoU score, GIoU and DFL losses
// and the gradient w.r.t. its 68 box logits (written with scalar stores over the zero fill).
__device__ PosOut positive_anchor(const Geo& g, const LossArgs& A, int n, int l, int hw, float gx1, float gy1,
                                  float gx2, float gy2, float scale_bbox, float scale_dfl, float& loss_bbox,
                                  float& loss_dfl) {
  const int HW = g.hw[l];
  const float fs = (float)g.stride[l];
  // weight_targets: max_c sigmoid(new-class logits), detached (:283-284)
  const float* cplane = A.s_cls.p[l] + ((size_t)n * g.C + g.ori) * HW + hw;
  float mx = -INFINITY;
  for (int c = 0; c < g.cn; ++c) mx = fmaxf(mx, __ldg(cplane + (size_t)c * HW));
  const float w = sigmoid_ref(mx);
  // Integral: softmax expectation per side (:40-54,285)
  const float* bplane = A.s_box.p[l] + (size_t)n * kBoxCh * HW + hw;
  float zmax[4], inv[4], d[4], lse[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    float z[kBins];
#pragma unroll
    for (int j = 0; j < kBins; ++j) z[j] = __ldg(bplane + (size_t)(s * kBins + j) * HW);
    float m = z[0];
#pragma unroll
    for (int j = 1; j < kBins; ++j) m = fmaxf(m, z[j]);
    float sum = 0.f, num = 0.f;
#pragma unroll
    for (int j = 0; j < kBins; ++j) {
      const float e = expf(z[j] - m);
      sum += e;
      num = fmaf((float)j, e, num);
    }
    zmax[s] = m;
    inv[s] = 1.0f / sum;
    d[s] = num * inv[s];
    lse[s] = m + logf(sum);
  }
  // anchor centre / stride is the grid coordinate itself (gfl_head.py:232-243, :281)
  const float cx = (float)(hw % g.w[l]), cy = (float)(hw / g.w[l]);
  const float px1 = cx - d[0], py1 = cy - d[1], px2 = cx + d[2], py2 = cy + d[3];   // distance2bbox
  const float tx1 = gx1 / fs, ty1 = gy1 / fs, tx2 = gx2 / fs, ty2 = gy2 / fs;       // :288
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
  const float giou = iou - (enc - uni) / enc;
  loss_bbox += w * (1.0f - giou);
  // d(1 - giou) / d(px1, py1, px2, py2)
  const float g_uni = (inter / (uni * uni) - 1.0f / enc) * pick_gt(uni_raw, 1e-6f);
  const float g_int = -1.0f / uni - g_uni;
  const float g_enc = (uni / (enc * enc)) * pick_gt(enc_raw, 1e-6f);
  const float g_iw = g_int * ih * (iw_raw >= 0.f ? 1.f : 0.f);
  const float g_ih = g_int * iw * (ih_raw >= 0.f ? 1.f : 0.f);
  const float g_ew = g_enc * eh * (ew_raw >= 0.f ? 1.f : 0.f);
  const float g_eh = g_enc * ew * (eh_raw >= 0.f ? 1.f : 0.f);
  const float hgt = py2 - py1, wid = px2 - px1;
  const float g_x1 = -g_uni * hgt - g_iw * pick_gt(px1, tx1) - g_ew * pick_gt(tx1, px1);
  const float g_y1 = -g_uni * wid - g_ih * pick_gt(py1, ty1) - g_eh * pick_gt(ty1, py1);
  const float g_x2 = g_uni * hgt + g_iw * pick_gt(tx2, px2) + g_ew * pick_gt(px2, tx2);
  const float g_y2 = g_uni * wid + g_ih * pick_gt(ty2, py2) + g_eh * pick_gt(py2, ty2);
  const float g_d[4] = {-g_x1, -g_y1, g_x2, g_y2};
  // DFL targets: bbox2distance clamped to [0, reg_max - 0.1] (transforms.py:221-230)
  const float tgt[4] = {cx - tx1, cy - ty1, tx2 - cx, ty2 - cy};
  const float cb = scale_bbox * w, cd = scale_dfl * w;
  float* gplane = A.g_box.p[l] + (size_t)n * kBoxCh * HW + hw;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const float y = fminf(fmaxf(tgt[s], 0.f), (float)(kBins - 1) - 0.1f);
    const int yl = (int)y;
    const float wl = (float)(yl + 1) - y, wr = y - (float)yl;
    const float zl = __ldg(bplane + (size_t)(s * kBins + yl) * HW);
    const float zr = __ldg(bplane + (size_t)(s * kBins + yl + 1) * HW);
    loss_dfl += w * ((lse[s] - zl) * wl + (lse[s] - zr) * wr);   // gfocal_loss.py:159-165
#pragma unroll
    for (int j = 0; j < kBins; ++j) {
      const float p = expf(__ldg(bplane + (size_t)(s * kBins + j) * HW) - zmax[s]) * inv[s];
      float gr = cb * g_d[s] * p * ((float)j - d[s]);
      gr += cd * (wl * (p - (j == yl ? 1.f : 0.f)) + wr * (p - (j == yl + 1 ? 1.f : 0.f)));
      gplane[(size_t)(s * kBins + j) * HW] = gr;
    }
  }
  return {iou, w};
}

template <bool VEC>
__device__ __forceinline__ void loss_tile(const Geo& g, const Workspace& ws, const LossArgs& A, int n, int l,
                                          int hw0, float (&part)[4]) {
  const int HW = g.hw[l];
  const Quad<VEC> q(hw0, HW);
  const size_t abase = (size_t)n * g.A + g.start[l];
  int gi[4];
  bool sel[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    gi[k] = q.ok[k] ? A.gt_inds[abase + q.hw[k]] : -1;
    sel[k] = q.ok[k] && (A.sel_flags[abase + q.hw[k]] & 1);   // ERS rows, gfl_increment_erd.py:149-151
  }
  const float avg1 = A.avg[0];
  const float avg2 = fmaxf(A.avg[1], 1.0f);                                        // :407 clamp_(min=1)
  const float inv_avg1 = 1.0f / (float)((double)avg1 + (double)kEps32);            // losses/utils.py:60-61
  const float scale_cls = upstream_of(A.upstream, acc_cls(l)) * g.w_cls * inv_avg1;
  const float scale_bbox = upstream_of(A.upstream, acc_bbox(l)) * g.w_bbox / (1.0f + kEps32) / avg2;
  const float scale_dfl = upstream_of(A.upstream, acc_dfl(l)) * g.w_dfl / 4.0f / avg2;

  // 1. box gradients: zero everywhere ...
  float* gbox = A.g_box.p[l] + (size_t)n * kBoxCh * HW;
#pragma unroll 4
  for (int c = 0; c < kBoxCh; ++c) q.store_zero(gbox + (size_t)c * HW);
  // ... except at positives (:273-310)
  int label[4];
  float score[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    label[k] = -1;
    score[k] = 0.f;
    if (gi[k] > 0) {
      const int gidx = A.gt_offsets[n] + gi[k] - 1;
      const long long lab = A.gt_labels[gidx];
      if (lab >= 0 && lab < g.cn) {
        const float4 gb = *reinterpret_cast<const float4*>(A.gt_boxes + (size_t)gidx * 4);
        const PosOut po = positive_anchor(g, A, n, l, q.hw[k], gb.x, gb.y, gb.z, gb.w, scale_bbox, scale_dfl,
                                          part[1], part[2]);
        label[k] = (int)lab;
        score[k] = po.score;
      }
    }
  }
  // 2. QFL over the new-class channels of every anchor (:260-261,317-320)
  const float* scls = A.s_cls.p[l] + (size_t)n * g.C * HW;
  float* gcls = A.g_cls.p[l] + (size_t)n * g.C * HW;
  float lw[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) lw[k] = gi[k] >= 0 ? 1.0f : 0.0f;   // label_weights, gfl_head.py:650-655,663
  float loss_cls = 0.f;
#pragma unroll 4
  for (int c = 0; c < g.cn; ++c) {
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
  part[0] += loss_cls;
  // 3. old-class channels: zero, except the classification-response L2 on ERS rows (:181-186,324-332)
#pragma unroll 4
  for (int c = 0; c < g.ori; ++c) q.store_zero(gcls + (size_t)c * HW);
  const float kc = (float)A.cls_count[n] * (float)g.ori;
  const float scale_dc = upstream_of(A.upstream, acc_dcls(n)) * A.dlw * 2.0f / kc;
  const float* tcls = A.t_cls.p[l] + (size_t)n * g.ori * HW;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (!sel[k]) continue;
    float sq = 0.f;
    for (int c = 0; c < g.ori; ++c) {
      const float df = __ldg(scls + (size_t)c * HW + q.hw[k]) - __ldg(tcls + (size_t)c * HW + q.hw[k]);
      sq = fmaf(df, df, sq);
      gcls[(size_t)c * HW + q.hw[k]] = scale_dc * df;
    }
    part[3] += sq;
  }
}

// ws.counters[1] = 1 when any upstream gradient differs from 1
__global__ void upstream_check_kernel(Workspace ws, const float* __restrict__ up, int total) {
  bool nonunit = false;
  for (int i = threadIdx.x; i < total; i += blockDim.x) nonunit |= (up[i] != 1.0f);
  const int any = __syncthreads_or(nonunit ? 1 : 0);
  if (threadIdx.x == 0) ws.counters[1] = any ? 1u : 0u;
}

__global__ void __launch_bounds__(kTileThreads) loss_main_kernel(Geo g, Workspace ws, LossArgs A) {
  if (A.skip_flag && *A.skip_flag == 0u) return;
  const int n = blockIdx.y;
  const int tile = blockIdx.x;
  const int l = level_of_tile(g, tile);
  const int hw0 = (tile - g.tile_start[l]) * kTile;
  float part[4] = {0.f, 0.f, 0.f, 0.f};   // cls, bbox, dfl, dist_cls
  if (g.vec[l])
    loss_tile<true>(g, ws, A, n, l, hw0, part);
  else
    loss_tile<false>(g, ws, A, n, l, hw0, part);
  __shared__ float red[kTileThreads / 32][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) part[i] = warp_sum(part[i]);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) red[threadIdx.x >> 5][i] = part[i];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int w = 0; w < kTileThreads / 32; ++w) s += (double)red[w][threadIdx.x];
    const int slot = threadIdx.x == 0 ? acc_cls(l) : threadIdx.x == 1 ? acc_bbox(l)
                     : threadIdx.x == 2 ? acc_dfl(l) : acc_dcls(n);
    if (s != 0.0) atomicAdd(ws.loss_acc + slot, s);
  }
}

// DFL-distribution distillation on the NMS survivors: KL(T=10) between student and teacher
// box distributions, weighted by the student's max old-class score (:204-221, kd_loss.py:12-37).
// One warp per kept row; gradients are added onto what loss_main_kernel wrote.
constexpr int kKdThreads = 256;

__global__ void __launch_bounds__(kKdThreads) kd_kernel(Geo g, Workspace ws, LossArgs A) {
  if (A.skip_flag && *A.skip_flag == 0u) return;
  const int n = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int M = A.keep_count[n];
  const float kT = g.T;
  const float scale = upstream_of(A.upstream, acc_dbox(g, n)) * A.dlw * g.w_ld / 4.0f * (kT * kT / (float)kBins) / kT;
  float lsum = 0.f;
  for (int r = blockIdx.x * (kKdThreads / 32) + warp; r < M; r += gridDim.x * (kKdThreads / 32)) {
    const int a = A.box_inds[(size_t)n * g.sel_cap + A.keep[(size_t)n * g.sel_cap + r]];
    const int l = level_of_anchor(g, a);
    const int HW = g.hw[l];
    const int hw = a - g.start[l];
    const float* cplane = A.s_cls.p[l] + (size_t)n * g.C * HW + hw;
    float mx = -INFINITY;
    for (int c = lane; c < g.ori; c += 32) mx = fmaxf(mx, __ldg(cplane + (size_t)c * HW));
    const float w = sigmoid_ref(warp_max(mx));                                    // :217-218
    const float* sp = A.s_box.p[l] + (size_t)n * kBoxCh * HW + hw;
    const float* tp = A.t_box.p[l] + (size_t)n * kBoxCh * HW + hw;
    float* gp = A.g_box.p[l] + (size_t)n * kBoxCh * HW + hw;
    float row = 0.f;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const bool on = lane < kBins;
      const size_t off = (size_t)(s * kBins + lane) * HW;
      const float zs = on ? __fdiv_rn(__ldg(sp + off), kT) : -INFINITY;
      const float zt = on ? __fdiv_rn(__ldg(tp + off), kT) : -INFINITY;
      const float ms = warp_max(zs), mt = warp_max(zt);
      const float es = on ? expf(zs - ms) : 0.f, et = on ? expf(zt - mt) : 0.f;
      const float ss = warp_sum(es), st = warp_sum(et);
      const float ps = es / ss, pt = et / st;
      const float logps = zs - ms - logf(ss), logpt = zt - mt - logf(st);
      const float kl = (on && pt > 0.f) ? pt * (logpt - logps) : 0.f;
      row += warp_sum(kl) / (float)kBins * (kT * kT);                              // .mean(1) * T*T
      if (on) gp[off] += scale * w * (ps - pt);
    }
    lsum += w * row;
  }
  __shared__ float red[kKdThreads / 32];
  if (lane == 0) red[warp] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kKdThreads / 32; ++w) s += (double)red[w];
    if (s != 0.0) atomicAdd(ws.loss_acc + acc_dbox(g, n), s);
  }
}

// Accumulators -> the reference's loss values, with its division order.
__global__ void finalize_kernel(Geo g, Workspace ws, LossArgs A) {
  if (A.skip_flag && *A.skip_flag == 0u) return;
  const int i = threadIdx.x;
  const int total = 3 * kLevels + 2 * g.n_img;
  if (i >= total) return;
  const float v = (float)ws.loss_acc[i];
  const float avg2 = fmaxf(A.avg[1], 1.0f);
  float out;
  if (i < kLevels) {
    out = g.w_cls * (v / (float)((double)A.avg[0] + (double)kEps32));
  } else if (i < 2 * kLevels) {
    out = g.w_bbox * (v / (1.0f + kEps32)) / avg2;
  } else if (i < 3 * kLevels) {
    out = g.w_dfl * (v / 4.0f) / avg2;
  } else if (i < 3 * kLevels + g.n_img) {
    const int n = i - 3 * kLevels;
    out = A.dlw * (float)(ws.loss_acc[i] / ((double)A.cls_count[n] * (double)g.ori));   // mean over K*ori; 0/0 -> NaN
  } else {
    out = A.dlw * (g.w_ld * (v / 4.0f));
  }
  A.losses[i] = out;
}

cudaError_t launch_loss(const Geo& g, const Workspace& ws, const LossArgs& a, cudaStream_t st) {
  const int total = 3 * kLevels + 2 * g.n_img;
  cudaError_t e = cudaMemsetAsync(ws.loss_acc, 0, sizeof(double) * total, st);
  if (e != cudaSuccess) return e;
  if (a.skip_flag) ERD_LAUNCH(kKUpCheck, st, (upstream_check_kernel<<<1, 128, 0, st>>>(ws, a.upstream, total)));
  ERD_LAUNCH(kKLossMain, st,
             (loss_main_kernel<<<dim3(g.tile_start[kLevels], g.n_img), kTileThreads, 0, st>>>(g, ws, a)));
  ERD_LAUNCH(kKKd, st, (kd_kernel<<<dim3(32, g.n_img), kKdThreads, 0, st>>>(g, ws, a)));
  ERD_LAUNCH(kKFinalize, st, (finalize_kernel<<<1, ((total + 31) / 32) * 32, 0, st>>>(g, ws, a)));
  return cudaGetLastError();
}

}  // namespace erd
