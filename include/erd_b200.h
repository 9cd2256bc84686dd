/*
 * erd_b200 -- C ABI of the B200-native GFL + Elastic Response Distillation loss path.
 *
 * The reference (Hi-FT/ERD, an MMDetection 3.0.0 fork) is pure Python and has no FFI of
 * its own; each entry point below names the reference Python interface (file:line under
 * the reference root) whose work it replaces.  INTEGRATION.md shows the ctypes binding a
 * maintainer adds on the reference side.
 *
 * Conventions
 *  - Plain C: raw device pointers, sizes, an opaque cudaStream_t passed as void*.
 *  - All tensors are borrowed for the duration of the call and must live on the device
 *    that is current when the call is made.  Nothing is allocated inside the library
 *    except in erd_create(); outputs and the workspace are caller-allocated.
 *  - Every call is asynchronous on `stream`; no call synchronises the device.
 *  - Return value: ERD_OK (0) or a negative ErdStatus.  Nothing throws.
 *  - Head outputs are contiguous fp32 NCHW per pyramid level, exactly what
 *    GFLHead.forward emits (mmdet/models/dense_heads/gfl_head.py:205-230).
 *  - Anchor numbering inside an image: level 0 first, row-major (y * W_l + x), as produced
 *    by AnchorGenerator.grid_priors (task_modules/prior_generators/anchor_generator.py:230-301).
 */
#ifndef ERD_B200_H_
#define ERD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ERD_MAX_LEVELS 5
#define ERD_ABI_VERSION 6

typedef enum ErdStatus {
  ERD_OK = 0,
  ERD_ERR_BAD_SHAPE = -1,     /* unsupported level count / reg_max / class split          */
  ERD_ERR_NULL = -2,          /* a required pointer is NULL                               */
  ERD_ERR_WORKSPACE = -3,     /* workspace smaller than erd_workspace_bytes()             */
  ERD_ERR_CUDA = -4,          /* a CUDA runtime call or launch failed (see erd_last_error) */
  ERD_ERR_NO_VALID_ANCHOR = -5 /* gfl_head.py:613-617 ValueError, detected on the host from pad_hw */
} ErdStatus;

/* Static description of one batch.  Mirrors the config keys of
 * configs/gfl_increment/gfl_r50_fpn_1x_coco_first_40_incre_last_40_cats.py:57-90. */
typedef struct ErdShape {
  int32_t num_imgs;                   /* N images on this rank                              */
  int32_t num_levels;                 /* must be 5 (strides 8..128)                         */
  int32_t num_classes;                /* C  student classes (80)                            */
  int32_t ori_classes;                /* ori teacher classes (40 or 70); new = C - ori      */
  int32_t reg_max;                    /* 16 -> 17 bins per side; only 16 is compiled        */
  int32_t level_h[ERD_MAX_LEVELS];    /* feature-map heights                                */
  int32_t level_w[ERD_MAX_LEVELS];    /* feature-map widths                                 */
  int32_t stride[ERD_MAX_LEVELS];     /* 8,16,32,64,128 (square strides)                    */
  int32_t total_gt;                   /* rows of gt_boxes a launch may cover: gt_offsets[N] of them are in use, so a
                                       * fixed capacity works too (and keeps a captured CUDA graph valid)        */
  float anchor_scale;                 /* octave_base_scale * 2**0 = 8                       */
  float loss_weight_cls;              /* QualityFocalLoss loss_weight (1.0), beta fixed at 2 */
  float loss_weight_bbox;             /* GIoULoss loss_weight (2.0), eps 1e-6               */
  float loss_weight_dfl;              /* DistributionFocalLoss loss_weight (0.25)           */
  float loss_weight_ld;               /* KnowledgeDistillationKLDivLoss loss_weight (0.25)  */
  float kd_temperature;               /* T (10)                                             */
  int32_t max_gt_per_img;             /* capacity: most GT boxes any image may hold (0 -> 128);
                                         sizes the positives' gradient-row buffer (45 rows / GT) */
} ErdShape;

/* Derived sizes a caller needs to allocate outputs. */
typedef struct ErdSizes {
  int64_t anchors_per_img;  /* A = sum_l H_l * W_l                                          */
  int64_t sel_cap;          /* per-image capacity of every ERS / NMS index list: A/5 + 1
                               (mean + 2 std selects at most A/5 rows, Cantelli)            */
  int64_t num_losses;       /* 3 * L + 2 * N: loss_cls[L], loss_bbox[L], loss_dfl[L],
                               loss_dist_cls[N], loss_dist_bbox[N]                          */
  size_t workspace_bytes;
} ErdSizes;

typedef struct ErdContext ErdContext; /* per-device helper streams/events (erd_create) */

int erd_abi_version(void);
const char* erd_last_error(void);

/* Sizes and workspace requirement for a shape. */
int erd_sizes(const ErdShape* shape, ErdSizes* out);

/* Zero the workspace.  Required once after allocation (and after a failed call): the kernels
 * leave every accumulator / counter / table clean for the next step themselves, so no memset
 * sits on a step's critical path. */
int erd_workspace_init(const ErdShape* shape, void* ws, void* stream);

int erd_create(ErdContext** ctx);
int erd_destroy(ErdContext* ctx);

/* Elastic Response Selection.
 * Replaces GFLIncrementERD.sel_pos / sel_pos_single
 * (mmdet/models/detectors/gfl_increment_erd.py:143-200).
 * t_cls[l]: (N, ori, H_l, W_l); t_box[l]: (N, 4*(reg_max+1), H_l, W_l).
 * Outputs, ascending anchor indices per image: cls_inds/box_inds (N, sel_cap) int32 with
 * device-side counts (N,) int32.  Also fills the per-anchor teacher cache in `ws`
 * (max sigmoid score, argmax class, max raw box logit, softmax-integral distances) that
 * erd_teacher_nms and erd_loss_fwd_bwd consume, `thr` (N,2) fp32 thresholds and
 * `sel_flags` (N, A) uint8: bit 0 = row in cls_inds, bit 1 = row in box_inds. */
int erd_ers_select(const ErdShape* shape, const float* const* t_cls, const float* const* t_box,
                   int32_t* cls_inds, int32_t* cls_count, int32_t* box_inds, int32_t* box_count,
                   float* thr, uint8_t* sel_flags, void* ws, void* stream);

/* Anchors, valid flags, ATSS assignment, pseudo sampling.
 * Replaces AnchorHead.get_anchors (dense_heads/anchor_head.py:164-199),
 * GFLHead.get_targets/_get_targets_single (dense_heads/gfl_head.py:504-679),
 * ATSSAssigner.assign (task_modules/assigners/atss_assigner.py:74-254) and
 * PseudoSampler.sample (task_modules/samplers/pseudo_sampler.py:26-60).
 * gt_boxes (total_gt,4) fp32 xyxy px; gt_labels (total_gt,) int64; gt_offsets (N+1,) int32
 * CSR offsets; pad_hw (N,2) int32 = img_meta['pad_shape'][:2].
 * Outputs: gt_inds (N, A) int32: -1 anchor outside pad_shape (label_weight 0), 0 background,
 * k>0 assigned to the k-th GT of its image (1-based); num_pos (N,) int32. */
int erd_atss_assign(const ErdShape* shape, const float* gt_boxes, const int64_t* gt_labels,
                    const int32_t* gt_offsets, const int32_t* pad_hw, int32_t* gt_inds,
                    int32_t* num_pos, void* ws, void* stream);

/* The two normalisers of the GT losses, left on the device for one 8-byte all-reduce.
 * Replaces the reduce_mean operands of gfl_head_increment_erd.py:390-391 and :406-407
 * (mmdet/utils/dist_utils.py:59-65): avg[0] = sum_img max(num_pos,1),
 * avg[1] = sum over positives of max_c sigmoid(student new-class logits).
 * The caller divides by world size, all-reduces (SUM) and hands the buffer to
 * erd_loss_fwd_bwd, which applies clamp(min=1) to avg[1].  The same pass decodes the positives
 * (softmax integral, gfl_head_increment_erd.py:285-292) and keeps their IoU quality scores and
 * the GIoU / DFL loss sums in `ws` for erd_loss_fwd_bwd. */
int erd_avg_factors(const ErdShape* shape, const float* const* s_cls, const float* const* s_box,
                    const float* gt_boxes, const int64_t* gt_labels, const int32_t* gt_offsets,
                    const int32_t* gt_inds, const int32_t* num_pos, float* avg, void* ws,
                    void* stream);

/* Teacher box decode + class-aware greedy IoU-NMS (threshold `iou_thr`, 0.005 in the
 * reference) over the ERS-selected rows.  Replaces the mmcv.ops.batched_nms call and the
 * decode in GFLHeadIncrementERD.distill_loss_by_image_single
 * (dense_heads/gfl_head_increment_erd.py:189-202).  Needs the teacher cache written by
 * erd_ers_select.  keep (N, sel_cap) int32 holds positions into the image's box_inds list
 * in descending-score order (what batched_nms returns); keep_count (N,) int32.  Survivors are
 * also marked in sel_flags (bit 2) for erd_loss_fwd_bwd. */
int erd_teacher_nms(const ErdShape* shape, const int32_t* box_inds, const int32_t* box_count,
                    const int32_t* pad_hw, float iou_thr, int32_t* keep, int32_t* keep_count,
                    uint8_t* sel_flags, void* ws, void* stream);

/* Fused forward + backward of QFL / GIoU / DFL and both distillation losses.
 * Replaces GFLHeadIncrementERD.loss_by_feat_single, distill_loss_by_image_single and the
 * glue of loss_by_feat (dense_heads/gfl_head_increment_erd.py:142-454) together with
 * quality_focal_loss / distribution_focal_loss (losses/gfocal_loss.py:12-53,143-165),
 * giou_loss (losses/iou_loss.py:110-126), knowledge_distillation_kl_div_loss
 * (losses/kd_loss.py:12-37) and weight_reduce_loss (losses/utils.py:30-65), and their
 * autograd backward.
 * upstream: NULL (every loss term has upstream gradient 1, what mmengine parse_losses
 * produces) or (num_losses,) fp32 per-term upstream gradients on the device.
 * skip_if_unit_upstream: when non-zero (and upstream != NULL) the call is a device-side
 * no-op if every upstream value equals 1 -- the autograd backward uses this to re-derive
 * gradients only when the caller weighted the loss terms, without a host sync.
 * losses: (num_losses,) fp32 in the order of ErdSizes.num_losses.
 * g_cls[l] (N,C,H_l,W_l) / g_box[l] (N,4*(reg_max+1),H_l,W_l): dense NCHW gradients,
 * fully overwritten.
 * ctx: NULL, or the context erd_step_prepare ran on -- the teacher NMS it forked is then
 * joined in front of the one kernel that needs the keep lists.  Must follow erd_avg_factors
 * (or erd_step_prepare) on the same workspace. */
int erd_loss_fwd_bwd(ErdContext* ctx, const ErdShape* shape, const float* const* s_cls,
                     const float* const* s_box,
                     const float* const* t_cls, const float* const* t_box, const float* gt_boxes,
                     const int64_t* gt_labels, const int32_t* gt_offsets, const int32_t* pad_hw,
                     const int32_t* gt_inds, const int32_t* num_pos, const int32_t* cls_inds, const int32_t* cls_count,
                     const uint8_t* sel_flags,
                     const int32_t* box_inds, const int32_t* box_count, const int32_t* keep, const int32_t* keep_count,
                     const float* avg, float dist_loss_weight, const float* upstream,
                     int32_t skip_if_unit_upstream, float* losses, float* const* g_cls, float* const* g_box, void* ws,
                     void* stream);

/* One training-step worth of the path in two calls around the caller's all-reduce:
 * erd_step_prepare = erd_ers_select + erd_atss_assign + erd_avg_factors + erd_teacher_nms,
 * erd_step_loss = erd_loss_fwd_bwd(ctx, ...).  Replaces GFLIncrementERD.loss
 * (detectors/gfl_increment_erd.py:202-220) minus the conv stacks.
 * Only assignment and avg factors are ordered on `stream` when erd_step_prepare returns (that is
 * what the all-reduce needs); the teacher side (selection, NMS) keeps running
 * on the context's helper streams and is joined by erd_loss_fwd_bwd(ctx, ...) where its results
 * are consumed.  A caller that reads ERS / NMS outputs itself must synchronise the device (or
 * call erd_loss_fwd_bwd first). */
typedef struct ErdStepBuffers {
  int32_t* cls_inds;
  int32_t* cls_count;
  int32_t* box_inds;
  int32_t* box_count;
  float* thr;
  uint8_t* sel_flags;
  int32_t* gt_inds;
  int32_t* num_pos;
  int32_t* keep;
  int32_t* keep_count;
  float* avg;
} ErdStepBuffers;

int erd_step_prepare(ErdContext* ctx, const ErdShape* shape, const float* const* t_cls,
                     const float* const* t_box, const float* const* s_cls, const float* const* s_box,
                     const float* gt_boxes, const int64_t* gt_labels, const int32_t* gt_offsets, const int32_t* pad_hw,
                     float iou_thr, const ErdStepBuffers* buf, void* ws, void* stream,
                     uint32_t flags);
/* flags for erd_step_prepare */
#define ERD_PREPARE_ERS_DONE 1u /* erd_ers_select already ran on these teacher tensors (sel_pos);
                                   its lists, counts, sel_flags and the teacher cache are reused */
#define ERD_PREPARE_TEACHER_CACHED 4u /* erd_teacher_head_fused already filled the teacher cache, the threshold sums and the
                                         stash on this workspace (it was ordered on `stream` before this call): the
                                         teacher pass is skipped; t_cls / t_box are only passed on (may hold NULLs) */
#define ERD_PREPARE_NO_EXCHANGE 2u /* do not post the avg factors to the peers even if the context has an
                                      exchange (the caller reduces buf->avg itself before erd_loss_fwd_bwd) */

/* reduce_mean of the two avg factors (mmdet/utils/dist_utils.py:59-65, call sites
 * gfl_head_increment_erd.py:390-391,406-407) over NVLink peer memory, for one process per GPU
 * on one node: avg[0..1] <- sum_r avg_r / world, identical bits on every rank, one small
 * kernel on `stream` (CUDA-graph capturable) instead of an NCCL launch.
 * peer_bufs[r]: rank r's exchange buffer of erd_avg_exchange_bytes() bytes as mapped into THIS
 * process (the host maps them, e.g. with torch symmetric memory or CUDA IPC), zeroed once by
 * its owner before the first call.  Collective: every rank calls it once per step.
 * A peer that never arrives (spin bound: minutes) is fatal, not silent: avg[0..1] become NaN, so every loss
 * of the step is NaN, and word [erd_avg_exchange_bytes()/4 - 3] of the own buffer becomes 1. */
size_t erd_avg_exchange_bytes(void);
int erd_avg_exchange(float* avg, void* const* peer_bufs, int32_t rank, int32_t world, void* stream);
/* The same exchange WITHOUT a launch of its own, for the fused step: once the context knows the peer buffers,
 * erd_step_prepare's assignment kernel posts this rank's factors to the peers from its last block, and the student
 * pass of the next erd_loss_fwd_bwd(ctx, ...) waits for the peers' factors in its prologue, averages them
 * (identical bits on every rank) and also WRITES the means into avg[0..1] for the caller.  The exchange then sits on
 * no stream's critical path.  peer_bufs as above; world <= 1 or NULL disables it.  Collective per step like
 * erd_avg_exchange; do not mix the two on the same buffers within a step. */
int erd_context_set_exchange(ErdContext* ctx, void* const* peer_bufs, int32_t rank, int32_t world);

/* --- teacher head-output producer fusion (next row of the scope table) -----------------
 * Replaces the teacher head's last convolutions -- gfl_cls / gfl_reg (3x3, 256 -> ori / 4*(reg_max+1)) + bias and the
 * box branch's Scale, mmdet/models/dense_heads/gfl_head.py:228-230, run under no_grad by
 * GFLIncrementERD.loss (mmdet/models/detectors/gfl_increment_erd.py:205) -- TOGETHER with the streaming half of
 * erd_ers_select: one tcgen05 (TF32, fp32 accumulate) implicit GEMM whose epilogue writes the per-anchor teacher
 * cache, the threshold sums and the stash directly, so the teacher logits are never re-read (and need not exist).
 * cls_feat[l] / reg_feat[l]: outputs of the teacher's cls / reg tower at level l, (N, H_l, W_l, 256) fp32 -- NHWC, i.e.
 * torch channels_last storage of the (N, 256, H_l, W_l) tensor; 16 B aligned.
 * head->w_cls / w_reg: the conv weights re-laid-out once by erd_teacher_head_pack (the teacher is frozen);
 * b_cls (ori,), b_reg (4*(reg_max+1),) biases; scale[l] the Scale parameters.
 * t_cls_out / t_box_out: NULL, or per-level NCHW fp32 tensors that receive the logits exactly as GFLHead.forward
 * would emit them (write-only; they are what erd_loss_fwd_bwd reads for an ERS anchor that missed the stash).
 * Zeroes cls_count / box_count like erd_ers_select's scan.  Follow with erd_step_prepare(..., ERD_PREPARE_TEACHER_CACHED)
 * on the same stream and workspace. */
typedef struct ErdTeacherHead {
  const float* w_cls;                 /* erd_teacher_head_pack(gfl_cls.weight)            */
  const float* w_reg;                 /* erd_teacher_head_pack(gfl_reg.weight)            */
  const float* b_cls;
  const float* b_reg;
  float scale[ERD_MAX_LEVELS];
} ErdTeacherHead;
/* floats of the packed image of an (out_channels, 256, 3, 3) weight */
size_t erd_teacher_head_packed_floats(int32_t out_channels);
int erd_teacher_head_pack(const float* w_oihw, int32_t out_channels, float* packed, void* stream);
int erd_teacher_head_fused(const ErdShape* shape, const ErdTeacherHead* head, const float* const* cls_feat,
                           const float* const* reg_feat, float* const* t_cls_out, float* const* t_box_out,
                           int32_t* cls_count, int32_t* box_count, void* ws, void* stream);

/* erd_ers_select without its streaming scan, for a workspace whose teacher cache and threshold sums
 * erd_teacher_head_fused has just written (same stream): thresholds, flags, counts and the ordered lists
 * (GFLIncrementERD.sel_pos, gfl_increment_erd.py:143-200).  Outputs as erd_ers_select. */
int erd_ers_select_cached(const ErdShape* shape, int32_t* cls_inds, int32_t* cls_count, int32_t* box_inds,
                          int32_t* box_count, float* thr, uint8_t* sel_flags, void* ws, void* stream);

/* --- inference post-process (next row of the scope table) ---------------------------
 * Replaces GFLHead._predict_by_feat_single (mmdet/models/dense_heads/gfl_head.py:408-502) with
 * filter_scores_and_topk (mmdet/models/utils/misc.py:308-354) and BaseDenseHead._bbox_post_process
 * (mmdet/models/dense_heads/base_dense_head.py:424-486), with_nms=True, for a whole batch.
 * cls_scores[l] (N, num_classes, H_l, W_l) logits and bbox_preds[l] (N, 4*(reg_max+1), H_l, W_l), fp32 NCHW as the
 * head emits them; img_hw (N,2) int32 = img_meta['img_shape'][:2] (clamp limits); inv_scale (N,2) fp32 =
 * 1 / scale_factor (w, h) for rescale=True, or NULL.  Outputs: dets (N, max_per_img, 5) = x1, y1, x2, y2, score in
 * descending score order, labels (N, max_per_img) int32 (-1 beyond the count), num_dets (N,) int32.
 * ErdShape: num_imgs, num_levels, num_classes, reg_max, level_h/w, stride are read.  nms_pre <= 3276. */
typedef struct ErdPredictConfig {
  int32_t nms_pre;        /* test_cfg.nms_pre: candidates kept per level (1000)   */
  int32_t max_per_img;    /* test_cfg.max_per_img (100)                           */
  float score_thr;        /* test_cfg.score_thr (0.05), strict >                  */
  float iou_threshold;    /* test_cfg.nms.iou_threshold (0.6), class-aware NMS    */
  float min_bbox_size;    /* test_cfg.min_bbox_size (0); < 0: no size filter      */
} ErdPredictConfig;
int erd_predict_workspace_bytes(const ErdShape* shape, const ErdPredictConfig* cfg, size_t* bytes);
int erd_predict(const ErdShape* shape, const ErdPredictConfig* cfg, const float* const* cls_scores,
                const float* const* bbox_preds, const int32_t* img_hw, const float* inv_scale, float* dets,
                int32_t* labels, int32_t* num_dets, void* workspace, void* stream);

/* Introspection for tests and diagnostics: device address and size of a named workspace array
 * ("t_slot": uint16 [N][A] stash row + 1 of each anchor; "pthr_state": uint32 [4] provisional thresholds as
 * ~ordered bits, 0 = none;
 * "t_m", "t_u": float [N][A] the teacher cache the thresholds are taken over; "t_arg": int32 [N][A] argmax class,
 * "t_dist": float [N][A][4] softmax-integral distances).  ERD_ERR_BAD_SHAPE for an
 * unknown name.  No reference counterpart. */
int erd_workspace_field(const ErdShape* shape, void* workspace, const char* name, void** ptr, size_t* bytes);

/* Launch accounting and optional per-kernel timing (CUDA events on the launching stream).
 * No reference counterpart; bench.py reports roofline numbers from it. */
int erd_profile_enable(unsigned int kernel_mask);   /* bit k set: time kernel id k; 0 = off */
unsigned long long erd_launch_count(void);   /* kernels launched by this library so far */
int erd_profile_num_kernels(void);
const char* erd_profile_kernel_name(int id);
int erd_profile_collect(float* total_ms, int* count);
int erd_profile_mark(void* stream);                          /* reference point for the timeline */
int erd_profile_timeline(float* start_ms, float* end_ms);    /* last launch of each kernel vs the mark */

#ifdef __cplusplus
}
#endif
#endif /* ERD_B200_H_ */
