"""Teacher head-output producer fusion: the teacher's last head convolutions
(gfl_head.py:228-230) as a tcgen05 implicit GEMM whose epilogue is the teacher pass.

Three claims, each checked through the C ABI:
  1. the logits it emits are the convolution's: against an fp64 convolution on the CPU, inside the worst-case bound
     of TF32 operand truncation (2 * 2**-10 * sum |x||w|, the tolerance of the precision cuDNN runs the fp32
     reference convolution in by default);
  2. its epilogue IS the teacher pass: cache (max sigmoid, argmax, max box logit, integral distances) bit-identical
     to erd_ers_select run on the emitted logits, thresholds and selections identical;
  3. a whole step behind it (ERD_PREPARE_TEACHER_CACHED) gives the same bits as the standard step on those logits.
"""
import pytest
import torch
import torch.nn.functional as F

from erd_b200.ops import ErdPath, TeacherHead
from erd_b200.synth import make_batch

pytestmark = pytest.mark.gpu


def _towers(n, shapes, seed):
    g = torch.Generator().manual_seed(seed)
    mk = lambda h, w: torch.randn(n, 256, h, w, generator=g).relu_()   # tower outputs are post-ReLU
    return [mk(h, w) for h, w in shapes], [mk(h, w) for h, w in shapes]


def _head_params(ori, seed):
    g = torch.Generator().manual_seed(seed + 1)
    w_cls = torch.randn(ori, 256, 3, 3, generator=g) * 0.03
    w_reg = torch.randn(68, 256, 3, 3, generator=g) * 0.03
    b_cls = torch.full((ori,), -4.59511985013459) + 0.3 * torch.randn(ori, generator=g)
    b_reg = 0.2 * torch.randn(68, generator=g)
    scales = [1.0, 0.9, 1.1, 1.25, 0.8]
    return w_cls, b_cls, w_reg, b_reg, scales


@pytest.mark.parametrize('ori,img_hw,n', [(40, (200, 264), 2), (70, (136, 200), 3), (40, (333, 190), 1)])
def test_fused_teacher_head(ori, img_hw, n):
    batch = make_batch(n, img_hw, ori=ori, seed=31 + ori, num_gt=(2, 5))
    shapes = batch.shapes
    cls_f, reg_f = _towers(n, shapes, seed=ori)
    w_cls, b_cls, w_reg, b_reg, scales = _head_params(ori, seed=ori)
    dev = 'cuda'
    head = TeacherHead(w_cls.to(dev), b_cls.to(dev), w_reg.to(dev), b_reg.to(dev), scales)
    cl = lambda xs: [x.to(dev).contiguous(memory_format=torch.channels_last) for x in xs]
    cls_d, reg_d = cl(cls_f), cl(reg_f)
    b = batch.to(dev)

    fused, plain = ErdPath(), ErdPath()
    pf = fused.plan(b.s_cls, batch.num_classes, ori, batch.reg_max)
    pp = plain.plan(b.s_cls, batch.num_classes, ori, batch.reg_max)
    assert pf is not pp
    t_cls = [torch.full((n, ori, h, w), float('nan'), device=dev) for h, w in shapes]
    t_box = [torch.full((n, 68, h, w), float('nan'), device=dev) for h, w in shapes]

    for rep in range(2):   # second round: the stash is live (provisional thresholds of round one)
        # ---- 1. the logits
        fused.teacher_head_fused(pf, head, cls_d, reg_d, t_cls, t_box)
        torch.cuda.synchronize()
        for l in range(5):
            x_c, x_r = cls_f[l].double(), reg_f[l].double()
            ref_c = F.conv2d(x_c, w_cls.double(), b_cls.double(), padding=1)
            ref_r = F.conv2d(x_r, w_reg.double(), b_reg.double(), padding=1) * scales[l]
            bound_c = 2.0e-3 * F.conv2d(x_c.abs(), w_cls.double().abs(), padding=1) + 1e-5
            bound_r = (2.0e-3 * F.conv2d(x_r.abs(), w_reg.double().abs(), padding=1) + 1e-5) * scales[l]
            err_c = (t_cls[l].cpu().double() - ref_c).abs()
            err_r = (t_box[l].cpu().double() - ref_r).abs()
            assert torch.isfinite(t_cls[l]).all() and torch.isfinite(t_box[l]).all(), f'level {l}: unwritten logits'
            assert (err_c <= bound_c).all(), f'level {l} cls: max err {err_c.max():.3e}, bound {bound_c.max():.3e}'
            assert (err_r <= bound_r).all(), f'level {l} box: max err {err_r.max():.3e}, bound {bound_r.max():.3e}'
            # and far inside it on average (truncation errors do not all line up)
            assert err_c.mean() < 0.1 * bound_c.mean() and err_r.mean() < 0.1 * bound_r.mean()

        # ---- 2. the epilogue is the teacher pass
        plain.ers_select(pp, t_cls, t_box)
        torch.cuda.synchronize()
        for name, dt in (('t_m', torch.int32), ('t_u', torch.int32), ('t_arg', torch.int32), ('t_dist', torch.int32)):
            a, c = pf.workspace_field(name, dt), pp.workspace_field(name, dt)
            assert torch.equal(a, c), f'{name}: {(a != c).sum().item()} of {a.numel()} words differ'

        # ---- 3. the step behind it
        outs = []
        for path, p, cached in ((fused, pf, True), (plain, pp, False)):
            p.set_targets(b.gt_bboxes, b.gt_labels, b.pad_shapes)
            if cached:   # (the cache was written above; the standard path scans the logits itself)
                pass
            _, losses, g_cls, g_box = path.step(t_cls, t_box, b.s_cls, b.s_box, None, None, None, batch.num_classes,
                                                ori, batch.reg_max, targets_set=True, teacher_cached=cached)
            torch.cuda.synchronize()
            outs.append((losses.clone(), [g.clone() for g in g_cls], [g.clone() for g in g_box], p.thr.clone(),
                         p.cls_count.clone(), p.box_count.clone(), p.sel_flags.clone(), p.keep_count.clone()))
        (lf, gcf, gbf, thrf, ccf, bcf, sff, kcf), (lp, gcp, gbp, thrp, ccp, bcp, sfp, kcp) = outs
        assert torch.equal(thrf.view(torch.int32), thrp.view(torch.int32)), (thrf, thrp)
        assert torch.equal(ccf, ccp) and torch.equal(bcf, bcp) and torch.equal(kcf, kcp)
        assert int(ccp.sum()) > 0 and int(bcp.sum()) > 0, 'nothing selected: the test would be vacuous'
        assert torch.equal(sff, sfp)
        assert torch.equal(lf.view(torch.int32), lp.view(torch.int32)), (lf, lp)
        for a, c in zip(gcf + gbf, gcp + gbp):
            assert torch.equal(a.view(torch.int32), c.view(torch.int32))
        if rep == 1:   # the fused epilogue stashed: selected anchors found their rows
            slots = pf.workspace_field('t_slot', torch.int16).view(n, -1)
            sel = (sff.view(n, -1) & 3) != 0
            assert (slots[sel] != 0).float().mean() > 0.5, 'the fused epilogue did not stash the selected columns'
        # the next round of the fused path starts from its own provisional thresholds


def test_detector_loss_with_fused_teacher_head():
    """GFLIncrementERD.loss (gfl_increment_erd.py:202-220) with ``fuse_teacher_head=True``: the teacher's towers in
    PyTorch, its last convolutions inside the teacher pass.  The loss dict and the student's gradients must be the
    bits of the standard route (teacher forward -> sel_pos -> loss) run on the logits the fused route emitted, and
    those logits must be the teacher's convolutions."""
    import torch.nn as nn
    from erd_b200.detector import GFLIncrementERD
    from erd_b200.head import GFLHeadIncrementERD, parse_losses

    torch.manual_seed(7)
    dev = 'cuda'
    tc = dict(assigner=dict(type='ATSSAssigner', topk=9), allowed_border=-1, pos_weight=-1)

    class Body(nn.Module):   # stand-in for backbone + FPN: five 256-channel levels
        def __init__(self):
            super().__init__()
            self.stem = nn.Conv2d(3, 256, 3, stride=8, padding=1)

        def forward(self, x):
            f = torch.relu(self.stem(x))
            out = [f]
            for _ in range(4):
                out.append(torch.nn.functional.avg_pool2d(out[-1], 2, 2, ceil_mode=True))
            return out

    class Teacher(nn.Module):
        def __init__(self):
            super().__init__()
            self.body = Body()
            self.bbox_head = GFLHeadIncrementERD(40, 256, stacked_convs=2, reg_max=16, train_cfg=tc)

        def extract_feat(self, x):
            return self.body(x)

        def forward(self, x):
            return self.bbox_head(self.extract_feat(x))

    teacher = Teacher().to(dev)
    with torch.no_grad():   # a trained-looking teacher: some confident responses
        teacher.bbox_head.gfl_cls.weight.mul_(6.0)
        teacher.bbox_head.gfl_reg.weight.mul_(6.0)
        teacher.bbox_head.scales.copy_(torch.tensor([1.0, 0.9, 1.1, 1.2, 0.8]))
    student_body = Body().to(dev)
    head = GFLHeadIncrementERD(80, 256, stacked_convs=2, reg_max=16, train_cfg=tc).to(dev)
    n, H, W = 2, 256, 320
    x = torch.randn(n, 3, H, W, device=dev)
    gts = [type('GT', (), dict(bboxes=torch.tensor([[20., 30., 120., 140.], [150., 60., 300., 220.]], device=dev),
                               labels=torch.tensor([3, 17], device=dev)))() for _ in range(n)]
    samples = [type('S', (), dict(gt_instances=g, metainfo=dict(img_shape=(H, W), pad_shape=(H, W))))() for g in gts]

    def run(fuse, ori_outs_override=None):
        det = GFLIncrementERD(head, 40, ori_model=teacher if ori_outs_override is None else None,
                              extract_feat=student_body, fuse_teacher_head=fuse)
        if ori_outs_override is not None:
            det.ori_model = lambda _x: ori_outs_override
        head.zero_grad(set_to_none=True)
        student_body.zero_grad(set_to_none=True)
        losses = det.loss(x, samples)
        parse_losses(losses).backward()
        torch.cuda.synchronize()
        flat = torch.stack(losses['loss_cls'] + losses['loss_bbox'] + losses['loss_dfl'] + losses['loss_dist_cls'] +
                           losses['loss_dist_bbox']).detach().clone()
        grads = [p.grad.detach().clone() for p in list(head.parameters()) + list(student_body.parameters())
                 if p.grad is not None]
        return flat, grads

    flat_f, grads_f = run(True)
    # the logits the fused route emitted (re-run the fused head to fetch them) against the teacher's own forward
    from erd_b200.head import fused_teacher_head
    with torch.no_grad():
        (t_cls, t_box), (cls_sel, box_sel), plan = fused_teacher_head(head.path, teacher.bbox_head, teacher.extract_feat(x), 80, 16)
        ref_cls, ref_box = teacher(x)
    for a, r in zip(t_cls + t_box, ref_cls + ref_box):
        assert (a - r).abs().max() <= 2e-2 * max(1.0, float(r.abs().max())), 'emitted logits are not the head convolutions'
    assert sum(int(c) for c in plan.cls_count) > 0 and sum(int(c) for c in plan.box_count) > 0
    flat_s, grads_s = run(False, ori_outs_override=(t_cls, t_box))
    assert torch.isfinite(flat_f).all()
    assert torch.equal(flat_f.view(torch.int32), flat_s.view(torch.int32)), (flat_f, flat_s)
    assert len(grads_f) == len(grads_s) and len(grads_f) > 0
    for a, c in zip(grads_f, grads_s):
        # (the head-output gradients are the same bits; cuDNN's TF32 backward of the student convs in between is
        # not run-to-run exact)
        assert float((a - c).abs().max()) <= 1e-3 * max(float(c.abs().max()), 1e-12)


@pytest.mark.parametrize('seed', [0, 1])
def test_fused_teacher_head_random_geometries(seed):
    """Ragged geometries (widths and heights that are no multiples of the 16-pixel patch or of 8, single-row levels,
    one image): the emitted logits against cuDNN's convolution of the same features (both TF32: loose tolerance), and
    the epilogue's cache + thresholds + selections bit-identical to the teacher pass on the emitted logits."""
    import random
    rng = random.Random(1000 + seed)
    dev = 'cuda'
    for _ in range(4):
        n = rng.choice([1, 2, 5])
        ori = rng.choice([40, 70, 24])
        img_hw = (rng.randrange(64, 420), rng.randrange(64, 520))
        batch = make_batch(n, img_hw, ori=ori, num_classes=max(80, ori + 10), seed=rng.randrange(1 << 20), num_gt=2)
        shapes = batch.shapes
        cls_f, reg_f = _towers(n, shapes, seed=rng.randrange(1 << 20))
        w_cls, b_cls, w_reg, b_reg, scales = _head_params(ori, seed=rng.randrange(1 << 20))
        head = TeacherHead(w_cls.to(dev), b_cls.to(dev), w_reg.to(dev), b_reg.to(dev), scales)
        cls_d = [x.to(dev).contiguous(memory_format=torch.channels_last) for x in cls_f]
        reg_d = [x.to(dev).contiguous(memory_format=torch.channels_last) for x in reg_f]
        b = batch.to(dev)
        fused, plain = ErdPath(), ErdPath()
        pf = fused.plan(b.s_cls, batch.num_classes, ori, batch.reg_max)
        pp = plain.plan(b.s_cls, batch.num_classes, ori, batch.reg_max)
        t_cls = [torch.full((n, ori, h, w), float('nan'), device=dev) for h, w in shapes]
        t_box = [torch.full((n, 68, h, w), float('nan'), device=dev) for h, w in shapes]
        fused.teacher_head_fused(pf, head, cls_d, reg_d, t_cls, t_box)
        fused.ers_select_cached(pf)
        plain.ers_select(pp, t_cls, t_box)
        torch.cuda.synchronize()
        for l in range(5):
            ref_c = F.conv2d(cls_d[l], w_cls.to(dev), b_cls.to(dev), padding=1)
            ref_r = F.conv2d(reg_d[l], w_reg.to(dev), b_reg.to(dev), padding=1) * scales[l]
            assert torch.isfinite(t_cls[l]).all() and torch.isfinite(t_box[l]).all(), (img_hw, l)
            assert (t_cls[l] - ref_c).abs().max() <= 3e-2 and (t_box[l] - ref_r).abs().max() <= 3e-2, (img_hw, ori, l)
        for name in ('t_m', 't_u', 't_arg', 't_dist'):
            assert torch.equal(pf.workspace_field(name, torch.int32), pp.workspace_field(name, torch.int32)), (img_hw, ori, name)
        assert torch.equal(pf.thr.view(torch.int32), pp.thr.view(torch.int32)), (img_hw, ori)
        assert torch.equal(pf.cls_count, pp.cls_count) and torch.equal(pf.box_count, pp.box_count)
        assert torch.equal(pf.sel_flags, pp.sel_flags)
        for i in range(n):
            assert torch.equal(pf.cls_inds[i, :int(pf.cls_count[i])], pp.cls_inds[i, :int(pp.cls_count[i])])
            assert torch.equal(pf.box_inds[i, :int(pf.box_count[i])], pp.box_inds[i, :int(pp.box_count[i])])
