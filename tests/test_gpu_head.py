"""The reference-facing plugin API on the GPU: GFLIncrementERD.sel_pos +
GFLHeadIncrementERD.loss_by_feat + autograd backward, against the oracle."""
import pytest
import torch

from erd_b200.synth import make_batch
from util import rel_err, run_oracle

pytestmark = pytest.mark.gpu


def _api_run(batch, dist_loss_weight=1.0, weights=None, foreign=False):
    from erd_b200.detector import GFLIncrementERD
    from erd_b200.head import GFLHeadIncrementERD, parse_losses
    head = GFLHeadIncrementERD(batch.num_classes, 256, reg_max=batch.reg_max, build_convs=False,
                               train_cfg=dict(assigner=dict(type='ATSSAssigner', topk=9), allowed_border=-1,
                                              pos_weight=-1))
    det = GFLIncrementERD(head, batch.ori, dist_loss_weight=dist_loss_weight)
    b = batch.to('cuda')
    s_cls = [t.requires_grad_() for t in b.s_cls]
    s_box = [t.requires_grad_() for t in b.s_box]
    sel = det.sel_pos(b.t_cls, b.t_box)
    if foreign:   # plain index tensors, as a caller of the reference API could pass
        sel = ([x.clone() for x in sel[0]], None, [x.clone() for x in sel[2]], None)
    gts = [type('GT', (), dict(bboxes=x, labels=y))() for x, y in zip(b.gt_bboxes, b.gt_labels)]
    metas = [dict(img_shape=i, pad_shape=p) for i, p in zip(batch.img_shapes, batch.pad_shapes)]
    out = head.loss_by_feat((b.t_cls, b.t_box), (s_cls, s_box), sel[0], sel[1], sel[2], sel[3], batch.ori,
                            dist_loss_weight, det, gts, metas)
    if weights is None:
        parse_losses(out).backward()
    else:
        flat = out['loss_cls'] + out['loss_bbox'] + out['loss_dfl'] + out['loss_dist_cls'] + out['loss_dist_bbox']
        sum(w * x for w, x in zip(weights, flat)).backward()
    torch.cuda.synchronize()
    return out, sel, [t.grad.cpu() for t in s_cls], [t.grad.cpu() for t in s_box]


def test_loss_by_feat_dict_contract_and_grads():
    batch = make_batch(2, (480, 640), ori=40, seed=61, num_gt=5, mode='trained', gt_size_pow=2.0)
    out, sel, g_cls, g_box = _api_run(batch)
    o = run_oracle(batch)
    assert set(out) == {'loss_cls', 'loss_bbox', 'loss_dfl', 'loss_dist_cls', 'loss_dist_bbox'}
    assert [len(out[k]) for k in ('loss_cls', 'loss_bbox', 'loss_dfl', 'loss_dist_cls', 'loss_dist_bbox')] == [5, 5, 5, 2, 2]
    for k, v in o['losses'].items():
        for x, y in zip(out[k], v):
            assert x.dim() == 0 and abs(float(x) - y) <= 1e-5 * max(abs(y), 1e-7)
    for l in range(5):
        assert rel_err(g_cls[l], o['g_cls'][l]) <= 1e-5 and rel_err(g_box[l], o['g_box'][l]) <= 1e-5
    # sel_pos returns the reference's four lists; indices materialise to int64 tensors
    for i in range(2):
        assert torch.equal(sel[0][i].cpu(), o['cls_inds'][i]) and torch.equal(sel[2][i].cpu(), o['box_inds'][i])
        assert sel[1][i].shape == (o['cls_inds'][i].numel(), 40) and sel[3][i].shape == (o['box_inds'][i].numel(), 68)


def test_weighted_terms_take_the_regrad_path():
    from oracle import erd_oracle as O
    batch = make_batch(2, (320, 480), ori=40, seed=62, num_gt=3, mode='trained', gt_size_pow=2.0)
    w = [0.5 + 0.1 * i for i in range(15 + 4)]
    out, _, g_cls, g_box = _api_run(batch, weights=w)
    s_cls = [t.clone().requires_grad_() for t in batch.s_cls]
    s_box = [t.clone().requires_grad_() for t in batch.s_box]
    ci, bi = O.sel_pos(batch.t_cls, batch.t_box)
    lo = O.loss_by_feat(batch.t_cls, batch.t_box, s_cls, s_box, ci, bi, batch.ori, 1.0, batch.gt_bboxes,
                        batch.gt_labels, batch.pad_shapes)
    flat = lo['loss_cls'] + lo['loss_bbox'] + lo['loss_dfl'] + lo['loss_dist_cls'] + lo['loss_dist_bbox']
    sum(a * x for a, x in zip(w, flat)).backward()
    for l in range(5):
        assert rel_err(g_cls[l], s_cls[l].grad) <= 1e-5 and rel_err(g_box[l], s_box[l].grad) <= 1e-5


def test_foreign_index_lists_are_honoured():
    batch = make_batch(2, (320, 480), ori=40, seed=63, num_gt=3, mode='trained', gt_size_pow=2.0)
    out_a, _, gca, gba = _api_run(batch)
    out_b, _, gcb, gbb = _api_run(batch, foreign=True)
    for k in out_a:
        assert [float(x) for x in out_a[k]] == [float(x) for x in out_b[k]]
    for l in range(5):
        assert torch.equal(gca[l], gcb[l]) and torch.equal(gba[l], gbb[l])


def test_head_loss_with_data_samples():
    from erd_b200.head import GFLHeadIncrementERD
    batch = make_batch(2, (256, 320), ori=40, seed=64, num_gt=2)
    head = GFLHeadIncrementERD(80, 256, build_convs=False)
    b = batch.to('cuda')

    class DS:
        def __init__(self, i):
            self.gt_instances = type('GT', (), dict(bboxes=b.gt_bboxes[i], labels=b.gt_labels[i]))()
            self.metainfo = dict(img_shape=batch.img_shapes[i], pad_shape=batch.pad_shapes[i])
    from erd_b200.detector import GFLIncrementERD
    det = GFLIncrementERD(head, 40)
    sel = det.sel_pos(b.t_cls, b.t_box)
    out = head.loss((b.t_cls, b.t_box), (b.s_cls, b.s_box), [DS(0), DS(1)], *sel, 40, 1, det)
    o = run_oracle(batch)
    for k, v in o['losses'].items():
        for x, y in zip(out[k], v):
            assert abs(float(x) - y) <= 1e-5 * max(abs(y), 1e-7)
