// Closed-form loss / gradient arithmetic shared by the positives prepass (loss.cu) and the dense
// student pass (student.cu).  References: quality_focal_loss (losses/gfocal_loss.py:12-53),
// giou_loss (losses/iou_loss.py:110-126) over bbox_overlaps (structures/bbox/bbox_overlaps.py:
// 151-199), SURVEY.md Appendix A for the gradients.
#pragma once
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
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// ----- positives: decode geometry, DFL target, GIoU gradient -----------------------------------
// Geometry of one positive: decoded box (anchor centre -/+ the four softmax-integral distances, in
// stride units), target box, aligned IoU / GIoU pieces (bbox_overlaps.py:151-169,189-199, eps 1e-6).
struct PosGeom {
  float iou, uni, uni_raw, enc, enc_raw, inter, iw, ih, iw_raw, ih_raw, ew, eh, ew_raw, eh_raw;
  float px1, py1, px2, py2, tx1, ty1, tx2, ty2, cx, cy;
};

__device__ __forceinline__ PosGeom pos_geom(const float (&d)[4], int x, int yy, float fs, float4 gb) {
  PosGeom s;
  s.cx = (float)x;   // anchor centre / stride is the grid coordinate itself (gfl_head.py:232-243, :281)
  s.cy = (float)yy;
  s.tx1 = gb.x / fs; s.ty1 = gb.y / fs; s.tx2 = gb.z / fs; s.ty2 = gb.w / fs;   // gfl_head_increment_erd.py:288
  s.px1 = s.cx - d[0]; s.py1 = s.cy - d[1]; s.px2 = s.cx + d[2]; s.py2 = s.cy + d[3];   // distance2bbox
  const float area_p = (s.px2 - s.px1) * (s.py2 - s.py1);
  const float area_t = (s.tx2 - s.tx1) * (s.ty2 - s.ty1);
  s.iw_raw = fminf(s.px2, s.tx2) - fmaxf(s.px1, s.tx1);
  s.ih_raw = fminf(s.py2, s.ty2) - fmaxf(s.py1, s.ty1);
  s.iw = fmaxf(s.iw_raw, 0.f);
  s.ih = fmaxf(s.ih_raw, 0.f);
  s.inter = s.iw * s.ih;
  s.uni_raw = area_p + area_t - s.inter;
  s.uni = fmaxf(s.uni_raw, 1e-6f);
  s.iou = s.inter / s.uni;
  s.ew_raw = fmaxf(s.px2, s.tx2) - fminf(s.px1, s.tx1);
  s.eh_raw = fmaxf(s.py2, s.ty2) - fminf(s.py1, s.ty1);
  s.ew = fmaxf(s.ew_raw, 0.f);
  s.eh = fmaxf(s.eh_raw, 0.f);
  s.enc_raw = s.ew * s.eh;
  s.enc = fmaxf(s.enc_raw, 1e-6f);
  return s;
}

// DFL target of one side: bbox2distance clamped to [0, reg_max - 0.1] (transforms.py:221-230) and
// the two interpolation weights of distribution_focal_loss (gfocal_loss.py:159-165)
struct DflTarget {
  float y, wl, wr;
  int yl;
};
__device__ __forceinline__ DflTarget dfl_target(const PosGeom& s, int side) {
  const float tgt = side == 0 ? s.cx - s.tx1 : side == 1 ? s.cy - s.ty1 : side == 2 ? s.tx2 - s.cx : s.ty2 - s.cy;
  DflTarget t;
  t.y = fminf(fmaxf(tgt, 0.f), (float)(kBins - 1) - 0.1f);
  t.yl = (int)t.y;
  t.wl = (float)(t.yl + 1) - t.y;
  t.wr = t.y - (float)t.yl;
  return t;
}

// d(1 - giou) / d(this side's distance), through distance2bbox
__device__ __forceinline__ float pos_side_giou_grad(const PosGeom& s, int side) {
  const float g_uni = (s.inter / (s.uni * s.uni) - 1.0f / s.enc) * pick_gt(s.uni_raw, 1e-6f);
  const float g_int = -1.0f / s.uni - g_uni;
  const float g_enc = (s.uni / (s.enc * s.enc)) * pick_gt(s.enc_raw, 1e-6f);
  const float g_iw = g_int * s.ih * (s.iw_raw >= 0.f ? 1.f : 0.f);
  const float g_ih = g_int * s.iw * (s.ih_raw >= 0.f ? 1.f : 0.f);
  const float g_ew = g_enc * s.eh * (s.ew_raw >= 0.f ? 1.f : 0.f);
  const float g_eh = g_enc * s.ew * (s.eh_raw >= 0.f ? 1.f : 0.f);
  const float hgt = s.py2 - s.py1, wid = s.px2 - s.px1;
  if (side == 0) return g_uni * hgt + g_iw * pick_gt(s.px1, s.tx1) + g_ew * pick_gt(s.tx1, s.px1);   // -d/dx1
  if (side == 1) return g_uni * wid + g_ih * pick_gt(s.py1, s.ty1) + g_eh * pick_gt(s.ty1, s.py1);   // -d/dy1
  if (side == 2) return g_uni * hgt + g_iw * pick_gt(s.tx2, s.px2) + g_ew * pick_gt(s.px2, s.tx2);   // d/dx2
  return g_uni * wid + g_ih * pick_gt(s.ty2, s.py2) + g_eh * pick_gt(s.py2, s.ty2);                  // d/dy2
}

// Quad form used by the positives' prepass: four adjacent lanes hold the four sides, each with the
// 17 logits z[] of its side (overwritten with exp(z - max)); every lane of the warp must call it
// (quad shuffles), dead lanes with finite dummy inputs.
struct PosSide {
  float zm, sum, inv, dmine;      // softmax of this side: max, sum exp, 1/sum, expectation (Integral, :40-54,285)
  float zl, zr;                   // the two raw logits the DFL cross-entropy reads
  DflTarget t;
  PosGeom g;
};

__device__ __forceinline__ PosSide pos_side_decode(float (&z)[kBins], int side, int x, int yy, float fs, float4 gb) {
  PosSide s;
  s.zm = z[0];
#pragma unroll
  for (int j = 1; j < kBins; ++j) s.zm = fmaxf(s.zm, z[j]);
  float raw[kBins];
#pragma unroll
  for (int j = 0; j < kBins; ++j) raw[j] = z[j];
  float sum = 0.f, num = 0.f;
#pragma unroll
  for (int j = 0; j < kBins; ++j) {
    z[j] = expf(z[j] - s.zm);
    sum += z[j];
    num = fmaf((float)j, z[j], num);
  }
  s.sum = sum;
  s.inv = 1.0f / sum;
  s.dmine = num * s.inv;
  float d[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) d[k] = __shfl_sync(0xffffffffu, s.dmine, (threadIdx.x & 28) | k, 32);
  s.g = pos_geom(d, x, yy, fs, gb);
  s.t = dfl_target(s.g, side);
  s.zl = 0.f;
  s.zr = 0.f;
#pragma unroll
  for (int j = 0; j < kBins; ++j) {
    s.zl = j == s.t.yl ? raw[j] : s.zl;
    s.zr = j == s.t.yl + 1 ? raw[j] : s.zr;
  }
  return s;
}

// reductions over the 8 lanes that share one box side in the warp-wide item layout (lane = side * 8 + b)
__device__ __forceinline__ float oct_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
}
__device__ __forceinline__ float oct_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v + __shfl_xor_sync(0xffffffffu, v, 4);
}

}  // namespace erd
