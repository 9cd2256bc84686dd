// C ABI of liberd_b200.so (include/erd_b200.h): argument checking, workspace carving and
// kernel sequencing.  No torch types; every call is asynchronous on the caller's stream.
#include <stdio.h>
#include <string.h>

#include <cstdlib>
#include "erd_common.cuh"

using namespace erd;

static thread_local char g_err[256] = "";

static int fail(ErdStatus s, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s", what);
  return (int)s;
}
static int fail_cuda(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return (int)ERD_ERR_CUDA;
}

static int make_geo(const ErdShape* s, Geo* g) {
  if (!s) return fail(ERD_ERR_NULL, "shape is NULL");
  if (s->num_levels != kLevels) return fail(ERD_ERR_BAD_SHAPE, "num_levels must be 5");
  if (s->reg_max != kBins - 1) return fail(ERD_ERR_BAD_SHAPE, "only reg_max=16 is compiled");
  if (s->num_imgs < 1 || s->num_classes < 2 || s->ori_classes < 1 || s->ori_classes >= s->num_classes)
    return fail(ERD_ERR_BAD_SHAPE, "need num_imgs>=1 and 1 <= ori_classes < num_classes");
  if (s->total_gt < 0) return fail(ERD_ERR_BAD_SHAPE, "total_gt < 0");
  memset(g, 0, sizeof(*g));
  g->n_img = s->num_imgs;
  g->C = s->num_classes;
  g->ori = s->ori_classes;
  g->cn = s->num_classes - s->ori_classes;
  g->total_gt = s->total_gt;
  g->w_cls = s->loss_weight_cls;
  g->w_bbox = s->loss_weight_bbox;
  g->w_dfl = s->loss_weight_dfl;
  g->w_ld = s->loss_weight_ld;
  g->T = s->kd_temperature;
  if (!(g->T > 0.f)) return fail(ERD_ERR_BAD_SHAPE, "kd_temperature must be > 0");
  int a = 0, t = 0;
  for (int l = 0; l < kLevels; ++l) {
    if (s->level_h[l] < 1 || s->level_w[l] < 1 || s->stride[l] < 1) return fail(ERD_ERR_BAD_SHAPE, "bad level size");
    g->h[l] = s->level_h[l];
    g->w[l] = s->level_w[l];
    g->hw[l] = s->level_h[l] * s->level_w[l];
    g->stride[l] = s->stride[l];
    g->start[l] = a;
    g->tile_start[l] = t;
    g->half[l] = 0.5f * (float)s->stride[l] * (s->anchor_scale > 0.f ? s->anchor_scale : 8.0f);
    g->vec[l] = (g->hw[l] % 4 == 0) ? 1 : 0;
    g->vec2[l] = (g->hw[l] % 2 == 0) ? 1 : 0;
    a += g->hw[l];
    t += (g->hw[l] + kTile - 1) / kTile;
  }
  g->tile_start[kLevels] = t;
  g->A = a;
  g->sel_cap = a / 5 + 1;
  {
    const long long mg = s->max_gt_per_img > 0 ? s->max_gt_per_img : 128;
    const long long cap = (long long)kTopK * kLevels * mg;
    g->pos_cap = (int)(cap < a ? cap : a);
  }
  return ERD_OK;
}

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static void carve(const Geo& g, void* base, Workspace* ws) {
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align_up(bytes);
    return r;
  };
  const size_t NA = (size_t)g.n_img * g.A, NS = (size_t)g.n_img * g.sel_cap;
  ws->t_m = (float*)take(NA * 4);
  ws->t_arg = (int*)take(NA * 4);
  ws->t_u = (float*)take(NA * 4);
  ws->t_dist = (float4*)take(NA * 16);
  size_t tiles32 = (size_t)g.A / 32 + kLevels;   // >= sum_l ceil(hw_l / 32)
  if (tiles32 < (size_t)head_partials_per_img(g)) tiles32 = (size_t)head_partials_per_img(g);   // tiny levels, fused teacher head
  ws->ers_part = (double*)take((size_t)g.n_img * tiles32 * 4 * 8);
  ws->t_stash = (float*)take((size_t)g.n_img * kStashRows * stash_pitch(g.ori) * 4);
  ws->t_slot = (unsigned short*)take(NA * 2);
  ws->pthr_state = (unsigned int*)take(4 * 4);
  ws->atss_key = (unsigned long long*)take(NA * 8);
  ws->pos_list = (int2*)take(NA * 8);
  ws->pos_counter = (int*)take((size_t)g.n_img * 4);
  ws->pos_rec = (PosRec*)take(NA * sizeof(PosRec));
  ws->pre_acc = (double*)take((size_t)(2 * kLevels + 1) * 8);
  ws->pre_pub = (double*)take((size_t)(2 * kLevels + 1) * 8);
  ws->keep_raw = (int*)take(NS * 4);
  ws->kd_loss = (float*)take(NA * 4);
  ws->pos_rows = (float*)take((size_t)g.n_img * g.pos_cap * kBoxCh * 4);
  ws->nms_nz = (unsigned long long*)take(NS * nms_nz_words(g.sel_cap) * 8);
  ws->counters = (unsigned int*)take(8 * 4);
  ws->nms_score = (float*)take(NS * 4);
  ws->nms_cls = (int*)take(NS * 4);
  ws->nms_box = (float4*)take(NS * 16);
  ws->nms_orig = (int*)take(NS * 4);
  ws->nms_tbox = (float4*)take(NS * 16);
  ws->nms_tscore = (float*)take(NS * 4);
  ws->nms_tcls = (int*)take(NS * 4);
  ws->nms_mask = (unsigned long long*)take(NS * nms_words(g.sel_cap) * 8);
  ws->loss_acc = (double*)take((size_t)(3 * kLevels + 2 * g.n_img) * 8);
  ws->bytes = off;
}

static void set_vec(Geo* g, const float* const* a, const float* const* b = nullptr, float* const* c = nullptr,
                    float* const* d = nullptr) {
  for (int l = 0; l < kLevels; ++l) {
    uintptr_t bits = 0;
    if (a) bits |= (uintptr_t)a[l];
    if (b) bits |= (uintptr_t)b[l];
    if (c) bits |= (uintptr_t)c[l];
    if (d) bits |= (uintptr_t)d[l];
    if (bits & 15) g->vec[l] = 0;
    if (bits & 7) g->vec2[l] = 0;
  }
}

static bool any_null(const void* const* p) {
  if (!p) return true;
  for (int l = 0; l < kLevels; ++l)
    if (!p[l]) return true;
  return false;
}
#define NULLS(x) any_null((const void* const*)(x))

static Ptr5 ptr5(const float* const* p) {
  Ptr5 r;
  for (int l = 0; l < kLevels; ++l) r.p[l] = p[l];
  return r;
}
static MPtr5 mptr5(float* const* p) {
  MPtr5 r;
  for (int l = 0; l < kLevels; ++l) r.p[l] = p[l];
  return r;
}

struct ErdContext {
  cudaStream_t side;             // teacher chain: ERS scan + select -> NMS
  cudaEvent_t fork, sel_done, nms_resolved, nms_all;
  bool nms_pending;              // nms_resolved recorded by erd_step_prepare, not yet waited on
  ExchangeInfo xchg;             // peer buffers of the avg-factor exchange (world <= 1: none)
  bool xchg_pending;             // erd_step_prepare posted this rank's factors; the next loss call waits for the peers'
};

extern "C" {

int erd_abi_version(void) { return ERD_ABI_VERSION; }
const char* erd_last_error(void) { return g_err; }

int erd_sizes(const ErdShape* shape, ErdSizes* out) {
  Geo g;
  int rc = make_geo(shape, &g);
  if (rc) return rc;
  if (!out) return fail(ERD_ERR_NULL, "out is NULL");
  Workspace ws;
  carve(g, nullptr, &ws);
  out->anchors_per_img = g.A;
  out->sel_cap = g.sel_cap;
  out->num_losses = 3 * kLevels + 2 * g.n_img;
  out->workspace_bytes = ws.bytes;
  return ERD_OK;
}

int erd_workspace_init(const ErdShape* shape, void* wsp, void* stream) {
  Geo g;
  int rc = make_geo(shape, &g);
  if (rc) return rc;
  if (!wsp) return fail(ERD_ERR_NULL, "erd_workspace_init: NULL workspace");
  Workspace ws;
  carve(g, wsp, &ws);
  cudaError_t e = cudaMemsetAsync(wsp, 0, ws.bytes, (cudaStream_t)stream);
  return e == cudaSuccess ? ERD_OK : fail_cuda(e, "erd_workspace_init");
}

int erd_workspace_field(const ErdShape* shape, void* wsp, const char* name, void** ptr, size_t* bytes) {
  Geo g;
  int rc = make_geo(shape, &g);
  if (rc) return rc;
  if (!wsp || !name || !ptr || !bytes) return fail(ERD_ERR_NULL, "erd_workspace_field: NULL argument");
  Workspace ws;
  carve(g, wsp, &ws);
  const size_t NA = (size_t)g.n_img * g.A;
  if (!strcmp(name, "t_slot")) { *ptr = ws.t_slot; *bytes = NA * 2; }
  else if (!strcmp(name, "pthr_state")) { *ptr = ws.pthr_state; *bytes = 16; }
  else if (!strcmp(name, "t_m")) { *ptr = ws.t_m; *bytes = NA * 4; }
  else if (!strcmp(name, "t_u")) { *ptr = ws.t_u; *bytes = NA * 4; }
  else if (!strcmp(name, "t_arg")) { *ptr = ws.t_arg; *bytes = NA * 4; }
  else if (!strcmp(name, "t_dist")) { *ptr = ws.t_dist; *bytes = NA * 16; }
  else return fail(ERD_ERR_BAD_SHAPE, "erd_workspace_field: unknown field");
  return ERD_OK;
}

int erd_create(ErdContext** ctx) {
  if (!ctx) return fail(ERD_ERR_NULL, "ctx is NULL");
  ErdContext* c = new ErdContext();
  // the teacher chain is short kernels behind one streaming pass: give it the highest priority so
  // its few CTAs are not queued behind the bandwidth-bound student pass running beside it
  int prio_least = 0, prio_greatest = 0;
  cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
  cudaError_t e = cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, prio_greatest);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->sel_done, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->nms_resolved, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->nms_all, cudaEventDisableTiming);
  c->nms_pending = false;
  c->xchg.world = 0;
  c->xchg.rank = 0;
  c->xchg_pending = false;
  if (e != cudaSuccess) {
    delete c;
    return fail_cuda(e, "erd_create");
  }
  *ctx = c;
  return ERD_OK;
}

int erd_context_set_exchange(ErdContext* c, void* const* peer_bufs, int32_t rank, int32_t world) {
  if (!c) return fail(ERD_ERR_NULL, "erd_context_set_exchange: NULL context");
  c->xchg.world = 0;
  c->xchg_pending = false;
  if (!peer_bufs || world <= 1) return ERD_OK;   // disabled
  if (world > kMaxRanks || rank < 0 || rank >= world) return fail(ERD_ERR_BAD_SHAPE, "erd_context_set_exchange: bad rank / world");
  for (int i = 0; i < kMaxRanks; ++i) c->xchg.peers.buf[i] = i < world ? (unsigned char*)peer_bufs[i] : nullptr;
  for (int i = 0; i < world; ++i)
    if (!c->xchg.peers.buf[i]) return fail(ERD_ERR_NULL, "erd_context_set_exchange: NULL peer buffer");
  c->xchg.rank = rank;
  c->xchg.world = world;
  return ERD_OK;
}

int erd_destroy(ErdContext* c) {
  if (!c) return ERD_OK;
  cudaStreamDestroy(c->side);
  cudaEventDestroy(c->fork);
  cudaEventDestroy(c->sel_done);
  cudaEventDestroy(c->nms_resolved);
  cudaEventDestroy(c->nms_all);
  delete c;
  return ERD_OK;
}

int erd_ers_select(const ErdShape* shape, const float* const* t_cls, const float* const* t_box, int32_t* cls_inds,
                   int32_t* cls_count, int32_t* box_inds, int32_t* box_count, float* thr, uint8_t* sel_flags,
                   void* wsp, void* stream) {
  Geo g;
  int rc = make_geo(shape, &g);
  if (rc) return rc;
  if (NULLS(t_cls) || NULLS(t_box) || !cls_inds || !cls_count || !box_inds || !box_count || !thr || !sel_flags || !wsp)
    return fail(ERD_ERR_NULL, "erd_ers_select: NULL argument");
  set_vec(&g, t_cls, t_box);
  Workspace ws;
  carve(g, wsp, &ws);
  cudaError_t e = launch_ers(g, ws, ptr5(t_cls), ptr5(t_box), cls_inds, cls_count, box_inds, box_count, thr,
                             sel_flags, (cudaStream_t)stream);
  return e == cudaSuccess ? ERD_OK : fail_cuda(e, "erd_ers_select");
}

size_t erd_teacher_head_packed_floats(int32_t out_channels) {
  return out_channels > 0 ? (size_t)head_ncls_pad(out_channels) * 256 * 9 : 0;
}

int erd_teacher_head_pack(const float* w_oihw, int32_t out_channels, float* packed, void* stream) {
  if (!w_oihw || !packed) return fail(ERD_ERR_NULL, "erd_teacher_head_pack: NULL argument");
  if (out_channels < 1 || out_channels > 256) return fail(ERD_ERR_BAD_SHAPE, "erd_teacher_head_pack: 1 <= out_channels <= 256");
  cudaError_t e = launch_head_pack(w_oihw, out_channels, packed, (cudaStream_t)stream);
  return e == cudaSuccess ? ERD_OK : fail_cuda(e, "erd_teacher_head_pack");
}

int erd_teacher_head_fused(const ErdShape* shape, const ErdTeacherHead* head, const float* const* cls_feat,
                           const float* const* reg_feat, float* const* t_cls_out, float* const* t_box_out,
                           int32_t* cls_count, int32_t* box_count, void* wsp, void* stream) {
  Geo g;
  int rc = make_geo(shape, &g);
  if (rc) return rc;
  if (!head || !head->w_cls || !head->w_reg || !head->b_cls || !head->b_reg || NULLS(cls_feat) || NULLS(reg_feat) ||
      !cls_count || !box_count || !wsp)
    return fail(ERD_ERR_NULL, "erd_teacher_head_fused: NULL argument");
  const bool emit = t_cls_out || t_box_out;
  if (emit && (!t_cls_out || !t_box_out)) return fail(ERD_ERR_NULL, "erd_teacher_head_fused: both logit outputs or none");
  MPtr5 oc, ob;
  for (int l = 0; l < kLevels; ++l) {
    if (((uintptr_t)cls_feat[l] | (uintptr_t)reg_feat[l]) & 15)
      return fail(ERD_ERR_BAD_SHAPE, "erd_teacher_head_fused: features must be 16 B aligned");
    if (emit && (!t_cls_out[l] || !t_box_out[l])) return fail(ERD_ERR_NULL, "erd_teacher_head_fused: NULL logit output");
    oc.p[l] = emit ? t_cls_out[l] : nullptr;
    ob.p[l] = emit ? t_box_out[l] : nullptr;
  }
  if (((uintptr_t)head->w_cls | (uintptr_t)head->w_reg) & 15)
    return fail(ERD_ERR_BAD_SHAPE, "erd_teacher_head_fused: packed weights must be 16 B aligned");
  if ((long long)g.n_img * g.h[0] * g.w[0] * 64 >= (1ll << 32))
    return fail(ERD_ERR_BAD_SHAPE, "erd_teacher_head_fused: level 0 exceeds 64 GB of features");
  Workspace ws;
  carve(g, wsp, &ws);
  cudaError_t e = launch_teacher_head(g, ws, ptr5(cls_feat), ptr5(reg_feat), head->w_cls, head->w_reg, head->b_cls,
                                      head->b_reg, head->scale, emit ? &oc : nullptr, emit ? &ob : nullptr, cls_count,
                                      box_count, (cudaStream_t)stream);
  return e == cudaSuccess ? ERD_OK : fail_cuda(e, "erd_teacher_head_fused");
}

int erd_ers_select_cached(const ErdShape* shape, int32_t* cls_inds, int32_t* cls_count, int32_t* box_inds,
                          int32_t* box_count, float* thr, uint8_t* sel_flags, void* wsp, void* stream) {
  Geo g;
  int rc = make_geo(shape, &g);
  if (rc) return rc;
  if (!cls_inds || !cls_count || !box_inds || !box_count || !thr || !sel_flags || !wsp)
    return fail(ERD_ERR_NULL, "erd_ers_select_cached: NULL argument");
  Workspace ws;
  carve(g, wsp, &ws);
  cudaError_t e = launch_ers_flags(g, ws, head_partials_per_img(g), thr, sel_flags, cls_count, box_count, (cudaStream_t)stream);
  if (e == cudaSuccess) e = launch_ers_lists(g, ws, cls_inds, cls_count, box_inds, box_count, thr, sel_flags, (cudaStream_t)stream);
  return e == cudaSuccess ? ERD_OK : fail_cuda(e, "erd_ers_select_cached");
}

int erd_atss_assign(const ErdShape* shape, const float* gt_boxes, const int64_t* gt_labels,
                    const int32_t* gt_offsets, const int32_t* pad_hw, int32_t* gt_inds, int32_t* num_pos, void* wsp,
                    void* stream) {
  Geo g;
  int rc = make_geo(shape, &g);
  if (rc) return rc;
  if (!gt_offsets || !pad_hw || !gt_inds || !num_pos || !wsp || (g.total_gt > 0 && (!gt_boxes || !gt_labels)))
    return fail(ERD_ERR_NULL, "erd_atss_assign: NULL argument");
  Workspace ws;
  carve(g, wsp, &ws);
  cudaError_t e = launch_atss(g, ws, gt_boxes, gt_labels, gt_offsets, pad_hw, gt_inds, num_pos, (cudaStream_t)stream);
  return e == cudaSuccess ? ERD_OK : fail_cuda(e, "erd_atss_assign");
}

int erd_avg_factors(const ErdShape* shape, const float* const* s_cls, const float* const* s_box,
                    const float* gt_boxes, const int64_t* gt_labels, const int32_t* gt_offsets,
                    const int32_t* gt_inds, const int32_t* num_pos, float* avg, void* wsp, void* stream) {
  Geo g;
  int rc = make_geo(shape, &g);
  if (rc) return rc;
  if (NULLS(s_cls) || NULLS(s_box) || !gt_offsets || !gt_inds || !num_pos || !avg || !wsp ||
      (g.total_gt > 0 && (!gt_labels || !gt_boxes)))
    return fail(ERD_ERR_NULL, "erd_avg_factors: NULL argument");
  if (g.total_gt > 0 && ((uintptr_t)gt_boxes & 15)) return fail(ERD_ERR_BAD_SHAPE, "gt_boxes must be 16 B aligned");
  Workspace ws;
  carve(g, wsp, &ws);
  cudaError_t e = launch_avg(g, ws, ptr5(s_cls), ptr5(s_box), gt_boxes, gt_labels, gt_offsets, gt_inds, num_pos, avg,
                             (cudaStream_t)stream);
  return e == cudaSuccess ? ERD_OK : fail_cuda(e, "erd_avg_factors");
}

int erd_teacher_nms(const ErdShape* shape, const int32_t* box_inds, const int32_t* box_count, const int32_t* pad_hw,
                    float iou_thr, int32_t* keep, int32_t* keep_count, uint8_t* sel_flags, void* wsp, void* stream) {
  Geo g;
  int rc = make_geo(shape, &g);
  if (rc) return rc;
  if (!box_inds || !box_count || !pad_hw || !keep || !keep_count || !sel_flags || !wsp)
    return fail(ERD_ERR_NULL, "erd_teacher_nms: NULL argument");
  Workspace ws;
  carve(g, wsp, &ws);
  cudaError_t e = launch_nms(g, ws, box_inds, box_count, pad_hw, iou_thr, keep, keep_count, sel_flags,
                             (cudaStream_t)stream, nullptr, nullptr);
  return e == cudaSuccess ? ERD_OK : fail_cuda(e, "erd_teacher_nms");
}

// erd_teacher_nms on the context's NMS stream: `resolved` fires as soon as the survivors are
// marked (all the loss needs), `all_done` after the keep list has been put in score order.
static int nms_on_side_stream(ErdContext* ctx, const ErdShape* shape, const ErdStepBuffers* b, const int32_t* pad_hw,
                              float iou_thr, void* wsp) {
  Geo g;
  int rc = make_geo(shape, &g);
  if (rc) return rc;
  Workspace ws;
  carve(g, wsp, &ws);
  cudaError_t e = launch_nms(g, ws, b->box_inds, b->box_count, pad_hw, iou_thr, b->keep, b->keep_count, b->sel_flags,
                             ctx->side, nullptr, ctx->nms_resolved);
  if (e == cudaSuccess) e = cudaEventRecord(ctx->nms_all, ctx->side);
  return e == cudaSuccess ? ERD_OK : fail_cuda(e, "erd_step_prepare nms");
}

int erd_loss_fwd_bwd(ErdContext* ctx, const ErdShape* shape, const float* const* s_cls, const float* const* s_box,
                     const float* const* t_cls, const float* const* t_box, const float* gt_boxes,
                     const int64_t* gt_labels, const int32_t* gt_offsets, const int32_t* pad_hw,
                     const int32_t* gt_inds, const int32_t* num_pos, const int32_t* cls_inds,
                     const int32_t* cls_count, const uint8_t* sel_flags, const int32_t* box_inds, const int32_t* box_count,
                     const int32_t* keep,
                     const int32_t* keep_count, const float* avg, float dist_loss_weight, const float* upstream,
                     int32_t skip_if_unit_upstream, float* losses, float* const* g_cls, float* const* g_box,
                     void* wsp, void* stream) {
  Geo g;
  int rc = make_geo(shape, &g);
  if (rc) return rc;
  if (NULLS(s_cls) || NULLS(s_box) || NULLS(t_cls) || NULLS(t_box) || NULLS(g_cls) || NULLS(g_box) || !gt_offsets ||
      !pad_hw || !gt_inds || !num_pos || !cls_inds || !cls_count || !sel_flags || !box_inds || !box_count || !keep || !keep_count ||
      !avg ||
      !losses || !wsp || (g.total_gt > 0 && (!gt_boxes || !gt_labels)))
    return fail(ERD_ERR_NULL, "erd_loss_fwd_bwd: NULL argument");
  set_vec(&g, s_cls, s_box, g_cls, g_box);
  if (g.total_gt > 0 && ((uintptr_t)gt_boxes & 15)) return fail(ERD_ERR_BAD_SHAPE, "gt_boxes must be 16 B aligned");
  Workspace ws;
  carve(g, wsp, &ws);
  LossArgs a;
  a.s_cls = ptr5(s_cls);
  a.s_box = ptr5(s_box);
  a.t_cls = ptr5(t_cls);
  a.t_box = ptr5(t_box);
  a.g_cls = mptr5(g_cls);
  a.g_box = mptr5(g_box);
  a.gt_boxes = gt_boxes;
  a.gt_labels = gt_labels;
  a.gt_offsets = gt_offsets;
  a.pad_hw = pad_hw;
  a.gt_inds = gt_inds;
  a.num_pos = num_pos;
  a.cls_inds = cls_inds;
  a.cls_count = cls_count;
  a.sel_flags = sel_flags;
  a.box_inds = box_inds;
  a.box_count = box_count;
  a.keep = keep;
  a.keep_count = keep_count;
  a.avg = avg;
  a.upstream = upstream;
  a.skip_flag = (upstream && skip_if_unit_upstream) ? ws.counters + 1 : nullptr;
  a.losses = losses;
  a.dlw = dist_loss_weight;
  a.xchg.world = 0;
  if (ctx && ctx->xchg_pending) {   // the fused prepare posted this rank's factors: this call's student pass averages
    a.xchg = ctx->xchg;
    ctx->xchg_pending = false;
  }
  // With a context the teacher chain forked by erd_step_prepare is joined where its results are
  // consumed: the ERS selection in front of the student pass, the NMS in front of the take-back.
  cudaError_t e;
  if (ctx) {
    LossStreams ls;
    ls.sel_ready = ctx->nms_pending ? ctx->sel_done : nullptr;
    ls.nms_done = ctx->nms_pending ? ctx->nms_resolved : nullptr;
    const bool had_nms = ctx->nms_pending;
    e = launch_loss(g, ws, a, (cudaStream_t)stream, &ls);
    // the score-ordered keep list is an output only: join it last
    if (e == cudaSuccess && had_nms) e = cudaStreamWaitEvent((cudaStream_t)stream, ctx->nms_all, 0);
    ctx->nms_pending = false;
  } else {
    e = launch_loss(g, ws, a, (cudaStream_t)stream, nullptr);
  }
  return e == cudaSuccess ? ERD_OK : fail_cuda(e, "erd_loss_fwd_bwd");
}

int erd_step_prepare(ErdContext* ctx, const ErdShape* shape, const float* const* t_cls, const float* const* t_box,
                     const float* const* s_cls, const float* const* s_box, const float* gt_boxes,
                     const int64_t* gt_labels, const int32_t* gt_offsets, const int32_t* pad_hw, float iou_thr,
                     const ErdStepBuffers* b, void* wsp, void* stream, uint32_t flags) {
  if (!ctx || !b) return fail(ERD_ERR_NULL, "erd_step_prepare: NULL ctx/buffers");
  cudaStream_t main = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
  if (ctx->nms_pending) {   // a previous prepare nobody consumed: do not race its teacher cache
    e = cudaStreamWaitEvent(main, ctx->nms_all, 0);
    ctx->nms_pending = false;
  }
  // Stream layout.  The caller's stream only carries what the avg-factor all-reduce needs
  // (ATSS + positives prepass).  The teacher side runs beside it on ctx->side: ERS scan + select
  // (sel_done) -> NMS (nms_resolved when the survivors are marked, nms_all when the keep list is
  // ordered).  erd_loss_fwd_bwd(ctx, ...) joins them exactly where their results are consumed.
  if (e == cudaSuccess) e = cudaEventRecord(ctx->fork, main);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->side, ctx->fork, 0);
  if (e != cudaSuccess) return fail_cuda(e, "erd_step_prepare fork");
  int rc = 0;
  if (!(flags & ERD_PREPARE_ERS_DONE)) {
    // the student pass only needs the teacher pass and the flags; the ordered lists are for the NMS
    Geo gt;
    rc = make_geo(shape, &gt);
    if (rc) return rc;
    if ((!(flags & ERD_PREPARE_TEACHER_CACHED) && (NULLS(t_cls) || NULLS(t_box))) || !b->cls_inds || !b->cls_count || !b->box_inds || !b->box_count || !b->thr ||
        !b->sel_flags || !wsp)
      return fail(ERD_ERR_NULL, "erd_step_prepare: NULL ERS argument");
    if (!(flags & ERD_PREPARE_TEACHER_CACHED)) set_vec(&gt, t_cls, t_box);
    Workspace wt;
    carve(gt, wsp, &wt);
    int tiles = 0;
    if (flags & ERD_PREPARE_TEACHER_CACHED) tiles = head_partials_per_img(gt);   // erd_teacher_head_fused wrote cache and sums
    else
    e = launch_teacher_pass(gt, wt, ptr5(t_cls), ptr5(t_box), b->cls_count, b->box_count, &tiles, ctx->side);
    if (e == cudaSuccess) e = launch_ers_flags(gt, wt, tiles, b->thr, b->sel_flags, b->cls_count, b->box_count, ctx->side);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->sel_done, ctx->side);
    if (e == cudaSuccess)
      e = launch_ers_lists(gt, wt, b->cls_inds, b->cls_count, b->box_inds, b->box_count, b->thr, b->sel_flags, ctx->side);
  } else {
    e = cudaEventRecord(ctx->sel_done, ctx->side);
  }
  if (e != cudaSuccess) return fail_cuda(e, "erd_step_prepare sel");
  if (!b->box_inds || !b->box_count || !pad_hw || !b->keep || !b->keep_count || !b->sel_flags || !wsp)
    return fail(ERD_ERR_NULL, "erd_step_prepare: NULL NMS buffer");
  rc = nms_on_side_stream(ctx, shape, b, pad_hw, iou_thr, wsp);
  if (rc) return rc;
  ctx->nms_pending = true;
  // erd_atss_assign + erd_avg_factors, with the decode and the positives prepass in one launch
  Geo g;
  rc = make_geo(shape, &g);
  if (rc) return rc;
  if (NULLS(s_cls) || NULLS(s_box) || !gt_offsets || !b->gt_inds || !b->num_pos || !b->avg ||
      (g.total_gt > 0 && (!gt_labels || !gt_boxes)))
    return fail(ERD_ERR_NULL, "erd_step_prepare: NULL assignment argument");
  if (g.total_gt > 0 && ((uintptr_t)gt_boxes & 15)) return fail(ERD_ERR_BAD_SHAPE, "gt_boxes must be 16 B aligned");
  Workspace ws;
  carve(g, wsp, &ws);
  const bool fused_xchg = ctx->xchg.world > 1 && !(flags & ERD_PREPARE_NO_EXCHANGE);
  e = launch_assign_avg(g, ws, ptr5(s_cls), ptr5(s_box), gt_boxes, gt_labels, gt_offsets, pad_hw, b->gt_inds,
                        b->num_pos, b->avg, fused_xchg ? &ctx->xchg : nullptr, main);
  ctx->xchg_pending = fused_xchg;
  return e == cudaSuccess ? ERD_OK : fail_cuda(e, "erd_step_prepare assignment");
}

}  // extern "C"
