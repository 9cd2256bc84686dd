"""Torch-facing wrappers over the C ABI: device memory, streams and the all-reduce.

PyTorch is plumbing here (allocation, current stream, ``torch.distributed``); all
arithmetic happens in liberd_b200.so.  There is no CPU path: tensors must be CUDA.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _native as N
from .dist_utils import peer_exchange, reduce_mean_

STRIDES = (8, 16, 32, 64, 128)


def _ptrs(ts: Sequence[torch.Tensor]):
    return N.PtrArray(*[t.data_ptr() for t in ts])


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _check_level_tensors(name: str, ts: Sequence[torch.Tensor], n: int, ch: int, shapes):
    if len(ts) != len(shapes):
        raise AssertionError(f'{name}: expected {len(shapes)} levels, got {len(ts)}')
    for t, (h, w) in zip(ts, shapes):
        if not t.is_cuda:
            raise RuntimeError(f'{name} must be CUDA tensors (erd_b200 has no CPU path)')
        if t.dtype != torch.float32 or tuple(t.shape) != (n, ch, h, w):
            raise AssertionError(f'{name}: expected float32 {(n, ch, h, w)}, got {t.dtype} {tuple(t.shape)}')


class Plan:
    """Buffers for one static batch geometry on one device (reused across steps)."""

    def __init__(self, key, device):
        n, shapes, num_classes, ori, reg_max, strides, scale, weights, max_gt = key
        self.max_gt = max_gt
        self.key, self.device = key, device
        self.n, self.shapes, self.C, self.ori, self.reg_max = n, shapes, num_classes, ori, reg_max
        self.shape = N.ErdShape()
        self.shape.num_imgs, self.shape.num_levels = n, len(shapes)
        self.shape.num_classes, self.shape.ori_classes, self.shape.reg_max = num_classes, ori, reg_max
        for l in range(min(len(shapes), N.MAX_LEVELS)):
            self.shape.level_h[l], self.shape.level_w[l] = shapes[l]
            self.shape.stride[l] = strides[l]
        self.shape.anchor_scale = scale
        (self.shape.loss_weight_cls, self.shape.loss_weight_bbox, self.shape.loss_weight_dfl,
         self.shape.loss_weight_ld, self.shape.kd_temperature) = weights
        self.shape.max_gt_per_img = max_gt
        sizes = N.ErdSizes()
        N.check(N.load().erd_sizes(C.byref(self.shape), C.byref(sizes)), 'erd_sizes')
        self.A, self.sel_cap = int(sizes.anchors_per_img), int(sizes.sel_cap)
        self.num_losses = int(sizes.num_losses)
        i32 = dict(dtype=torch.int32, device=device)
        self.ws = torch.empty(int(sizes.workspace_bytes), dtype=torch.uint8, device=device)
        with torch.cuda.device(device):   # zero once; the kernels keep it clean afterwards
            N.check(N.load().erd_workspace_init(C.byref(self.shape), self.ws.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream), 'erd_workspace_init')
        self.cls_inds = torch.empty(n, self.sel_cap, **i32)
        self.box_inds = torch.empty(n, self.sel_cap, **i32)
        self.keep = torch.empty(n, self.sel_cap, **i32)
        self.cls_count = torch.zeros(n, **i32)
        self.box_count = torch.zeros(n, **i32)
        self.keep_count = torch.zeros(n, **i32)
        self.num_pos = torch.zeros(n, **i32)
        self.gt_inds = torch.empty(n, self.A, **i32)
        self.thr = torch.empty(n, 2, dtype=torch.float32, device=device)
        self.sel_flags = torch.zeros(n, self.A, dtype=torch.uint8, device=device)
        self.avg = torch.zeros(2, dtype=torch.float32, device=device)
        self.meta = torch.zeros(3 * n + 1, **i32)            # gt_offsets (n+1) | pad_hw (2n)
        # pinned staging for `meta`, rotated: a host that runs ahead of the GPU must not rewrite a buffer
        # whose asynchronous copy has not run yet (each buffer is re-used only after its copy's event)
        self._meta_hosts = [torch.zeros(3 * n + 1, dtype=torch.int32).pin_memory() for _ in range(3)]
        self._meta_events = [None, None, None]
        self._meta_turn = 0
        # fixed-capacity GT buffers (n * max_gt rows): addresses and launch geometry never change, so a CUDA graph
        # captured over the step stays valid when the next batch's GT is loaded
        self.gt_boxes = torch.zeros(n * max_gt, 4, dtype=torch.float32, device=device)
        self.gt_labels = torch.zeros(n * max_gt, dtype=torch.int64, device=device)
        self.shape.total_gt = n * max_gt
        self.ers_generation = 0      # bumped every time the ERS buffers are rewritten
        self.bufs = N.ErdStepBuffers(
            self.cls_inds.data_ptr(), self.cls_count.data_ptr(), self.box_inds.data_ptr(),
            self.box_count.data_ptr(), self.thr.data_ptr(), self.sel_flags.data_ptr(), self.gt_inds.data_ptr(),
            self.num_pos.data_ptr(),
            self.keep.data_ptr(), self.keep_count.data_ptr(), self.avg.data_ptr())

    @property
    def gt_offsets(self):
        return self.meta[:self.n + 1]

    @property
    def pad_hw(self):
        return self.meta[self.n + 1:]

    def workspace_field(self, name: str, dtype: torch.dtype) -> torch.Tensor:
        """A named workspace array as a tensor view (tests / diagnostics; include/erd_b200.h)."""
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        N.check(N.load().erd_workspace_field(C.byref(self.shape), self.ws.data_ptr(), name.encode(), C.byref(ptr),
                                             C.byref(nbytes)), 'erd_workspace_field')
        off = ptr.value - self.ws.data_ptr()
        return self.ws[off:off + nbytes.value].view(dtype)

    def set_targets(self, gt_bboxes: Sequence[torch.Tensor], gt_labels: Sequence[torch.Tensor],
                    pad_shapes: Sequence[Tuple[int, int]]):
        """Pack per-image GT into CSR form and upload offsets + pad shapes in one copy."""
        n = self.n
        if not (len(gt_bboxes) == len(gt_labels) == len(pad_shapes) == n):
            raise AssertionError('gt / meta list lengths must equal the batch size')   # gfl_head.py:518
        turn = self._meta_turn
        self._meta_turn = (turn + 1) % len(self._meta_hosts)
        if self._meta_events[turn] is not None:
            self._meta_events[turn].synchronize()
        meta_host = self._meta_hosts[turn]
        counts = [int(b.shape[0]) for b in gt_bboxes]
        vals, off = [], 0
        for c in counts:                       # CSR offsets
            vals.append(off)
            off += c
        vals.append(off)
        for i in range(n):
            if counts[i] > self.max_gt:
                raise ValueError(f'image {i} has {counts[i]} GT boxes, plan capacity is {self.max_gt}')
            ph, pw = int(pad_shapes[i][0]), int(pad_shapes[i][1])
            if ph < 1 or pw < 1:
                # reference: ValueError when an image has no valid anchor (gfl_head.py:613-617)
                raise ValueError('There is no valid anchor inside the image boundary.')
            vals += [ph, pw]
        meta_host.copy_(torch.tensor(vals, dtype=torch.int32))   # one host-side copy instead of 3n element writes
        self.meta.copy_(meta_host, non_blocking=True)
        ev = self._meta_events[turn] or torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._meta_events[turn] = ev
        if off > 0:
            self.gt_boxes[:off].copy_(torch.cat([b.reshape(-1, 4) for b in gt_bboxes]), non_blocking=True)
            self.gt_labels[:off].copy_(torch.cat([l.reshape(-1) for l in gt_labels]), non_blocking=True)


    def load_selection(self, cls_inds: Sequence[torch.Tensor], box_inds: Sequence[torch.Tensor]):
        """Adopt caller-provided ERS index lists (the reference's ``loss_by_feat`` takes them as
        arguments): pad them into the plan buffers and rebuild the per-anchor flags."""
        if len(cls_inds) != self.n or len(box_inds) != self.n:
            raise AssertionError('one index tensor per image expected')
        self.sel_flags.zero_()
        for lst, inds, cnt, bit in ((cls_inds, self.cls_inds, self.cls_count, 1),
                                    (box_inds, self.box_inds, self.box_count, 2)):
            counts = []
            for i, x in enumerate(lst):
                x = x.reshape(-1).to(device=self.device)
                if x.numel() > self.sel_cap:
                    raise ValueError(f'index list of image {i} exceeds capacity {self.sel_cap}')
                inds[i, :x.numel()] = x.to(torch.int32)
                row = self.sel_flags[i]
                row[x.long()] = row[x.long()] | bit
                counts.append(x.numel())
            cnt.copy_(torch.tensor(counts, dtype=torch.int32), non_blocking=False)


class TeacherHead:
    """The frozen teacher's last head convolutions in the layout ``erd_teacher_head_fused`` streams
    (gfl_cls / gfl_reg weights re-laid-out once, biases, per-level Scale values; gfl_head.py:228-230)."""

    def __init__(self, gfl_cls_weight: torch.Tensor, gfl_cls_bias: torch.Tensor, gfl_reg_weight: torch.Tensor,
                 gfl_reg_bias: torch.Tensor, scales: Sequence[float]):
        lib = N.load()
        if not gfl_cls_weight.is_cuda:
            raise RuntimeError('erd_b200 runs on CUDA tensors only; there is no CPU fallback')
        for w in (gfl_cls_weight, gfl_reg_weight):
            if tuple(w.shape[1:]) != (256, 3, 3):
                raise ValueError(f'teacher head weight must be (O, 256, 3, 3), got {tuple(w.shape)}')
        if len(scales) != N.MAX_LEVELS:
            raise ValueError('one Scale value per level')
        self.ori = int(gfl_cls_weight.shape[0])
        self.packed = []
        for w in (gfl_cls_weight, gfl_reg_weight):
            w = w.detach().float().contiguous()
            out = torch.empty(lib.erd_teacher_head_packed_floats(int(w.shape[0])), dtype=torch.float32, device=w.device)
            N.check(lib.erd_teacher_head_pack(w.data_ptr(), int(w.shape[0]), out.data_ptr(), _stream()),
                    'erd_teacher_head_pack')
            self.packed.append(out)
        self.b_cls = gfl_cls_bias.detach().float().contiguous()
        self.b_reg = gfl_reg_bias.detach().float().contiguous()
        self.c = N.ErdTeacherHead(self.packed[0].data_ptr(), self.packed[1].data_ptr(), self.b_cls.data_ptr(),
                                  self.b_reg.data_ptr(), (C.c_float * N.MAX_LEVELS)(*[float(v) for v in scales]))


class ErdPath:
    """The hot path behind the reference's ``sel_pos`` / ``loss_by_feat`` (one per process)."""

    def __init__(self, strides: Sequence[int] = STRIDES, anchor_scale: float = 8.0, nms_iou_thr: float = 0.005,
                 loss_weights: Tuple[float, float, float, float] = (1.0, 2.0, 0.25, 0.25), kd_T: float = 10.0):
        """loss_weights = (QFL, GIoU, DFL, LD) loss_weight values of the head config."""
        self.lib = N.load()
        self.weights = tuple(float(w) for w in loss_weights) + (float(kd_T),)
        self.strides, self.anchor_scale, self.nms_iou_thr = tuple(strides), float(anchor_scale), float(nms_iou_thr)
        self._plans: Dict[tuple, Plan] = {}
        self.max_plans = 8
        self._cap: Dict[tuple, int] = {}
        self._ctx: Dict[int, C.c_void_p] = {}
        self._fused_exchange: Dict[int, bool] = {}   # device -> the context posts / waits inside the step's kernels
        self._posted = False                         # the last prepare() posted this rank's factors

    def _context(self, device) -> C.c_void_p:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx not in self._ctx:
            ctx = C.c_void_p()
            with torch.cuda.device(idx):
                N.check(self.lib.erd_create(C.byref(ctx)), 'erd_create')
            self._ctx[idx] = ctx
        return self._ctx[idx]

    def plan(self, s_or_t_cls: Sequence[torch.Tensor], num_classes: int, ori: int, reg_max: int = 16,
             max_gt: int = 0) -> Plan:
        """``max_gt``: most GT boxes any image of the batch holds (rounded up to a power of two,
        at least 128); the plan with the largest capacity seen so far is reused."""
        t0 = s_or_t_cls[0]
        if not t0.is_cuda:
            raise RuntimeError('erd_b200 runs on CUDA tensors only; there is no CPU fallback')
        shapes = tuple((int(t.shape[2]), int(t.shape[3])) for t in s_or_t_cls)
        cap = 128
        while cap < max_gt:
            cap *= 2
        base = (int(t0.shape[0]), shapes, int(num_classes), int(ori), int(reg_max), self.strides, self.anchor_scale,
                self.weights)
        cap = max(cap, self._cap.get(base + (t0.device,), 128))
        self._cap[base + (t0.device,)] = cap
        key = base + (cap,)
        full = key + (t0.device,)
        plan = self._plans.pop(full, None)
        if plan is None:
            plan = Plan(key, t0.device)
            # Resize(keep_ratio) + pad_size_divisor yield many padded batch geometries and capacity doublings
            # supersede plans: keep the most recently used few (a plan is ~140 MB of workspace at 16 images)
            while len(self._plans) >= self.max_plans:
                self._plans.pop(next(iter(self._plans)))
        self._plans[full] = plan   # most recently used last
        return plan

    # ---- individual stages (used by the parity tests and by sel_pos) -----------------
    def ers_select(self, p: Plan, t_cls, t_box):
        _check_level_tensors('teacher cls_scores', t_cls, p.n, p.ori, p.shapes)
        _check_level_tensors('teacher bbox_preds', t_box, p.n, 4 * (p.reg_max + 1), p.shapes)
        N.check(self.lib.erd_ers_select(C.byref(p.shape), _ptrs(t_cls), _ptrs(t_box), p.cls_inds.data_ptr(),
                                        p.cls_count.data_ptr(), p.box_inds.data_ptr(), p.box_count.data_ptr(),
                                        p.thr.data_ptr(), p.sel_flags.data_ptr(), p.ws.data_ptr(), _stream()),
                'erd_ers_select')
        p.ers_generation += 1

    def teacher_head_fused(self, p: Plan, head: TeacherHead, cls_feats, reg_feats, t_cls=None, t_box=None):
        """The teacher head's last convolutions fused with the teacher pass (``erd_teacher_head_fused``).
        ``cls_feats`` / ``reg_feats``: tower outputs (N, 256, H, W) in channels_last storage.  ``t_cls`` / ``t_box``:
        optional NCHW tensors that receive the logits.  Follow with ``prepare(..., teacher_cached=True)``."""
        if head.ori != p.ori:
            raise ValueError('teacher head classes != plan ori classes')
        for name, fs in (('cls_feat', cls_feats), ('reg_feat', reg_feats)):
            if len(fs) != len(p.shapes):
                raise ValueError(f'{name}: one tensor per level')
            for f, (h, w) in zip(fs, p.shapes):
                if tuple(f.shape) != (p.n, 256, h, w) or f.dtype != torch.float32:
                    raise ValueError(f'{name}: expected fp32 {(p.n, 256, h, w)}, got {tuple(f.shape)} {f.dtype}')
                if not f.is_contiguous(memory_format=torch.channels_last):
                    raise ValueError(f'{name}: channels_last (NHWC) storage required')
        emit = t_cls is not None
        if emit:
            _check_level_tensors('teacher cls_scores', t_cls, p.n, p.ori, p.shapes)
            _check_level_tensors('teacher bbox_preds', t_box, p.n, 4 * (p.reg_max + 1), p.shapes)
        oc, ob = (_ptrs(t_cls), _ptrs(t_box)) if emit else (None, None)
        N.check(self.lib.erd_teacher_head_fused(C.byref(p.shape), C.byref(head.c), _ptrs(cls_feats), _ptrs(reg_feats),
                                                C.byref(oc) if emit else None, C.byref(ob) if emit else None,
                                                p.cls_count.data_ptr(), p.box_count.data_ptr(), p.ws.data_ptr(),
                                                _stream()), 'erd_teacher_head_fused')

    def ers_select_cached(self, p: Plan):
        """Thresholds, flags and ordered lists from the cache ``teacher_head_fused`` has just written."""
        N.check(self.lib.erd_ers_select_cached(C.byref(p.shape), p.cls_inds.data_ptr(), p.cls_count.data_ptr(),
                                               p.box_inds.data_ptr(), p.box_count.data_ptr(), p.thr.data_ptr(),
                                               p.sel_flags.data_ptr(), p.ws.data_ptr(), _stream()),
                'erd_ers_select_cached')
        p.ers_generation += 1

    def atss_assign(self, p: Plan):
        N.check(self.lib.erd_atss_assign(C.byref(p.shape), p.gt_boxes.data_ptr(), p.gt_labels.data_ptr(),
                                         p.gt_offsets.data_ptr(), p.pad_hw.data_ptr(), p.gt_inds.data_ptr(),
                                         p.num_pos.data_ptr(), p.ws.data_ptr(), _stream()), 'erd_atss_assign')

    def avg_factors(self, p: Plan, s_cls, s_box):
        self._posted = False   # fresh local factors: a following reduce_avg has to reduce them
        N.check(self.lib.erd_avg_factors(C.byref(p.shape), _ptrs(s_cls), _ptrs(s_box), p.gt_boxes.data_ptr(),
                                         p.gt_labels.data_ptr(),
                                         p.gt_offsets.data_ptr(), p.gt_inds.data_ptr(), p.num_pos.data_ptr(),
                                         p.avg.data_ptr(), p.ws.data_ptr(), _stream()), 'erd_avg_factors')

    def teacher_nms(self, p: Plan):
        N.check(self.lib.erd_teacher_nms(C.byref(p.shape), p.box_inds.data_ptr(), p.box_count.data_ptr(),
                                         p.pad_hw.data_ptr(), self.nms_iou_thr, p.keep.data_ptr(),
                                         p.keep_count.data_ptr(), p.sel_flags.data_ptr(), p.ws.data_ptr(),
                                         _stream()), 'erd_teacher_nms')

    def _ensure_exchange(self, device) -> bool:
        """Give the device's context the peer buffers of the avg-factor exchange (first call: a collective
        rendezvous, outside any graph capture).  True when the exchange is fused into the step's kernels."""
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx not in self._fused_exchange:
            ex = peer_exchange(self.lib, device)
            if ex is not None:
                N.check(self.lib.erd_context_set_exchange(self._context(device), ex.peers, ex.rank, ex.world_size),
                        'erd_context_set_exchange')
            self._fused_exchange[idx] = ex is not None
        return self._fused_exchange[idx]

    def reduce_avg(self, p: Plan):
        """reduce_mean of both normalisers (dist_utils.py:59-65) in one 8-byte exchange.  After a fused
        ``prepare`` on one node nothing is left to do here: the assignment kernel has posted this rank's factors
        to the peers and the student pass averages them in its prologue.  Otherwise: a single kernel over NVLink
        peer memory when the ranks share a node, else one NCCL all-reduce."""
        if self._posted:
            return
        ex = peer_exchange(self.lib, p.device)
        if ex is not None:
            ex.reduce_mean_(p.avg)
        else:
            reduce_mean_(p.avg)

    def loss_fwd_bwd(self, p: Plan, t_cls, t_box, s_cls, s_box, g_cls, g_box, losses, dist_loss_weight: float,
                     upstream: Optional[torch.Tensor] = None, skip_if_unit: bool = False):
        N.check(self.lib.erd_loss_fwd_bwd(
            self._context(p.device), C.byref(p.shape), _ptrs(s_cls), _ptrs(s_box), _ptrs(t_cls), _ptrs(t_box), p.gt_boxes.data_ptr(),
            p.gt_labels.data_ptr(), p.gt_offsets.data_ptr(), p.pad_hw.data_ptr(), p.gt_inds.data_ptr(),
            p.num_pos.data_ptr(), p.cls_inds.data_ptr(), p.cls_count.data_ptr(), p.sel_flags.data_ptr(), p.box_inds.data_ptr(), p.box_count.data_ptr(),
            p.keep.data_ptr(),
            p.keep_count.data_ptr(), p.avg.data_ptr(), float(dist_loss_weight),
            upstream.data_ptr() if upstream is not None else None, 1 if skip_if_unit else 0,
            losses.data_ptr(), _ptrs(g_cls), _ptrs(g_box), p.ws.data_ptr(), _stream()), 'erd_loss_fwd_bwd')

    # ---- fused step -------------------------------------------------------------------
    def prepare(self, p: Plan, t_cls, t_box, s_cls, s_box, ers_done: bool = False, teacher_cached: bool = False):
        self._posted = self._ensure_exchange(p.device)
        N.check(self.lib.erd_step_prepare(
            self._context(p.device), C.byref(p.shape), _ptrs(t_cls), _ptrs(t_box), _ptrs(s_cls), _ptrs(s_box),
            p.gt_boxes.data_ptr(), p.gt_labels.data_ptr(), p.gt_offsets.data_ptr(), p.pad_hw.data_ptr(),
            self.nms_iou_thr, C.byref(p.bufs), p.ws.data_ptr(), _stream(),
            (1 if ers_done else 0) | (4 if teacher_cached else 0)),
            'erd_step_prepare')
        if not ers_done:
            p.ers_generation += 1

    def step(self, t_cls, t_box, s_cls, s_box, gt_bboxes, gt_labels, pad_shapes, num_classes: int, ori: int,
             reg_max: int = 16, dist_loss_weight: float = 1.0, upstream: Optional[torch.Tensor] = None,
             g_cls: Optional[List[torch.Tensor]] = None, g_box: Optional[List[torch.Tensor]] = None,
             targets_set: bool = False, ers_done: bool = False, teacher_cached: bool = False):
        """ERS + assignment + NMS + fused loss forward/backward.
        Returns (plan, losses (3L+2N,), g_cls[5], g_box[5])."""
        p = self.plan(s_cls, num_classes, ori, reg_max,
                      max((int(b.shape[0]) for b in gt_bboxes), default=0) if gt_bboxes is not None else 0)
        _check_level_tensors('cls_scores', s_cls, p.n, p.C, p.shapes)
        _check_level_tensors('bbox_preds', s_box, p.n, 4 * (p.reg_max + 1), p.shapes)
        _check_level_tensors('teacher cls_scores', t_cls, p.n, p.ori, p.shapes)
        _check_level_tensors('teacher bbox_preds', t_box, p.n, 4 * (p.reg_max + 1), p.shapes)
        if not targets_set:
            p.set_targets(gt_bboxes, gt_labels, pad_shapes)
        if g_cls is None:
            g_cls = [torch.empty_like(t) for t in s_cls]
        if g_box is None:
            g_box = [torch.empty_like(t) for t in s_box]
        losses = torch.empty(p.num_losses, dtype=torch.float32, device=p.device)
        self.prepare(p, t_cls, t_box, s_cls, s_box, ers_done, teacher_cached)
        self.reduce_avg(p)
        self.loss_fwd_bwd(p, t_cls, t_box, s_cls, s_box, g_cls, g_box, losses, dist_loss_weight, upstream)
        return p, losses, g_cls, g_box


_default_path: Optional[ErdPath] = None


def default_path() -> ErdPath:
    global _default_path
    if _default_path is None:
        _default_path = ErdPath()
    return _default_path
