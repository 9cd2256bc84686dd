"""Parity of the CUDA path (through the C ABI) against the fixtures generated from the real
reference and against the CPU oracle on identical seeded inputs.

Bar (BASELINE.json north_star): ATSS assignment, ERS index sets and NMS keep lists bit-exact;
losses and input gradients within 1e-5 relative (fp32).  Gradients are compared relative to the
largest reference magnitude of the same tensor (elementwise ratios are meaningless for the
values that cancel to ~0) and, more strictly, elementwise with a floor of 1e-3 of that scale.
"""
import pytest
import torch

from erd_b200.synth import make_batch
from oracle.golden_cases import CASES, case_batch, load_golden
from util import elem_rel_err, rel_err, run_cuda, run_oracle

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _assert_ints_equal(c, o, n):
    for i in range(n):
        assert torch.equal(c['gt_inds'][i], o['gt_inds'][i]), f'image {i}: ATSS assigned_gt_inds differ'
        assert torch.equal(c['cls_inds'][i], o['cls_inds'][i]), f'image {i}: ERS cls_inds differ'
        assert torch.equal(c['box_inds'][i], o['box_inds'][i]), f'image {i}: ERS bbox_inds differ'
        assert torch.equal(c['keep'][i], o['keep'][i]), f'image {i}: NMS keep list differs'


def _assert_losses_close(c_losses, o_losses):
    for k, v in o_losses.items():
        for j, (x, y) in enumerate(zip(c_losses[k], v)):
            assert abs(x - y) <= RTOL * max(abs(y), 1e-7), f'{k}[{j}]: {x} vs {y}'


def _assert_grads_close(c, o):
    for l in range(5):
        for key in ('g_cls', 'g_box'):
            a, b = c[key][l], o[key][l]
            assert rel_err(a, b) <= RTOL, f'{key}[{l}] rel-to-scale {rel_err(a, b):.2e}'
            scale = float(b.abs().max())
            if scale > 0:   # elementwise, for every element above 10 % of the tensor's scale
                assert elem_rel_err(a, b, 0.1 * scale) <= 2 * RTOL, f'{key}[{l}] elementwise'
            assert torch.equal(a != 0, b != 0) or float((a - b).abs().max()) <= RTOL * scale, f'{key}[{l}] support'


@pytest.mark.parametrize('name', list(CASES))
def test_cuda_matches_reference_fixture(name):
    """CUDA path vs outputs of the REAL reference (committed fixtures)."""
    gold = load_golden(name)
    batch = case_batch(name)
    c = run_cuda(batch)
    n = batch.num_imgs
    for i in range(n):
        assert torch.equal(c['cls_inds'][i], gold['cls_inds'][i])
        assert torch.equal(c['box_inds'][i], gold['box_inds'][i])
        assert torch.equal(c['keep'][i], gold['keep'][i])
        gi = c['gt_inds'][i]
        assert torch.equal((gi > 0).nonzero().squeeze(1), gold['pos'][i])
        assert torch.equal(gi[gi > 0], gold['pos_gt'][i])
        assert int((gi < 0).sum()) == gold['num_invalid'][i]
    _assert_losses_close(c['losses'], gold['losses'])
    for l in range(5):
        for key in ('g_cls', 'g_box'):
            g, dg = c[key][l].reshape(-1), gold[key][l]
            vals = torch.cat([dg['val'].reshape(-1), dg.get('nz_val', dg['val']).reshape(-1)])
            scale = max(float(vals.abs().max()) if vals.numel() else 0.0, 1e-30)
            assert float((g[dg['idx']] - dg['val']).abs().max()) <= RTOL * scale
            assert abs(float(g.double().sum()) - dg['sum']) <= RTOL * max(dg['abssum'], 1e-30)
            assert abs(float((g.double() ** 2).sum()) - dg['sqsum']) <= 10 * RTOL * max(dg['sqsum'], 1e-30)
            if 'nz_idx' in dg:   # sparse gradient: same support, same values
                assert torch.equal((g != 0).nonzero().squeeze(1), dg['nz_idx'])
                if dg['nz_idx'].numel():
                    assert float((g[dg['nz_idx']] - dg['nz_val']).abs().max()) <= RTOL * scale
    if 'g_cls_full' in gold:
        full = dict(g_cls=gold['g_cls_full'], g_box=gold['g_box_full'])
        _assert_grads_close(c, full)


@pytest.mark.parametrize('kw', [
    dict(num_imgs=2, img_hw=(800, 1333), ori=40, seed=101),                                        # configs[0]
    dict(num_imgs=16, img_hw=(800, 1333), ori=40, seed=102),                                       # configs[1]
    dict(num_imgs=4, img_hw=(800, 1333), ori=70, seed=103, mode='trained', gt_size_pow=2.0),       # configs[3]
    dict(num_imgs=2, img_hw=(1600, 1600), ori=40, seed=104, num_gt=100, mode='trained', gt_size_pow=3.0),  # configs[4]
    dict(num_imgs=3, img_hw=(480, 640), ori=40, seed=105, num_gt=[0, 0, 0]),                       # no GT at all
    dict(num_imgs=3, img_hw=(512, 512), ori=40, seed=106, num_gt=[5, 1, 9], mode='trained', gt_size_pow=2.0,
         pad_shapes=[(512, 512), (300, 500), (512, 260)]),                                         # ragged pads
    dict(num_imgs=1, img_hw=(96, 96), ori=40, seed=107, num_gt=2, gt_size_pow=1.2),                # tiny, levels 1x1
    dict(num_imgs=2, img_hw=(333, 500), ori=10, num_classes=20, seed=108, num_gt=6, gt_size_pow=2.0),  # odd sizes
], ids=['cfg1', 'cfg2_n16', 'cfg4_70_10', 'cfg5_dense', 'no_gt', 'ragged_pad', 'tiny', 'odd_20cls'])
def test_cuda_matches_oracle(kw):
    batch = make_batch(**kw)
    o, c = run_oracle(batch), run_cuda(batch)
    _assert_ints_equal(c, o, batch.num_imgs)
    # the oracle reports avg2 after clamp(min=1) (gfl_head_increment_erd.py:407); the device buffer holds the raw sum
    assert c['avg'][0] == o['avg'][0] and abs(max(c['avg'][1], 1.0) - o['avg'][1]) <= 1e-6 * o['avg'][1]
    _assert_losses_close(c['losses'], o['losses'])
    _assert_grads_close(c, o)


def test_dist_loss_weight_and_upstream_weights():
    """Per-term upstream gradients (what autograd hands back when the caller weights the loss
    terms) and dist_loss_weight != 1 against autograd on the oracle."""
    from oracle import erd_oracle as O
    batch = make_batch(2, (320, 480), ori=40, seed=21, num_gt=4, mode='trained', gt_size_pow=2.0)
    n = batch.num_imgs
    g = torch.Generator().manual_seed(0)
    up = torch.rand(15 + 2 * n, generator=g) + 0.5
    s_cls = [t.clone().requires_grad_() for t in batch.s_cls]
    s_box = [t.clone().requires_grad_() for t in batch.s_box]
    ci, bi = O.sel_pos(batch.t_cls, batch.t_box)
    losses = O.loss_by_feat(batch.t_cls, batch.t_box, s_cls, s_box, ci, bi, batch.ori, 2.5, batch.gt_bboxes,
                            batch.gt_labels, batch.pad_shapes)
    flat = losses['loss_cls'] + losses['loss_bbox'] + losses['loss_dfl'] + losses['loss_dist_cls'] + losses['loss_dist_bbox']
    sum(u * x for u, x in zip(up, flat)).backward()
    c = run_cuda(batch, dist_loss_weight=2.5, upstream=up.cuda())
    for j, x in enumerate(flat):
        assert abs(float(c['loss_vec'][j]) - float(x)) <= RTOL * max(abs(float(x)), 1e-7)
    o = dict(g_cls=[t.grad for t in s_cls], g_box=[t.grad for t in s_box])
    _assert_grads_close(c, o)


def test_run_to_run_determinism_and_idempotence():
    """Same inputs, same plan, twice: identical integers, identical losses and gradients bit-for-bit
    (the fp64 accumulators make the atomics order-independent at fp32)."""
    batch = make_batch(4, (800, 1333), ori=40, seed=33, mode='trained', gt_size_pow=2.0)
    from erd_b200.ops import ErdPath
    path = ErdPath()
    a = run_cuda(batch, path=path)
    b = run_cuda(batch, path=path)
    _assert_ints_equal(a, b, batch.num_imgs)
    assert torch.equal(a['loss_vec'], b['loss_vec'])
    for l in range(5):
        assert torch.equal(a['g_cls'][l], b['g_cls'][l]) and torch.equal(a['g_box'][l], b['g_box'][l])


def test_full_size_properties_n16():
    """Size-independent properties at BASELINE.json's full size (16 x 22 400 anchors):
    gradient support and linearity in the upstream gradient."""
    batch = make_batch(16, (800, 1333), ori=40, seed=44, mode='trained', gt_size_pow=2.0)
    n = batch.num_imgs
    from erd_b200.ops import ErdPath
    path = ErdPath()
    c1 = run_cuda(batch, path=path)
    up = torch.full((15 + 2 * n,), 3.0, device='cuda')
    c3 = run_cuda(batch, path=path, upstream=up)
    A = batch.anchors_per_image
    starts = [0]
    for h, w in batch.shapes:
        starts.append(starts[-1] + h * w)
    for l in range(5):
        for key in ('g_cls', 'g_box'):
            ref = 3.0 * c1[key][l]
            assert float((c3[key][l] - ref).abs().max()) <= 1e-6 * float(ref.abs().max())
        hw = batch.shapes[l][0] * batch.shapes[l][1]
        for i in range(n):
            # old-class cls gradient is non-zero exactly on the ERS cls rows of the image
            rows = (c1['g_cls'][l][i, :batch.ori].reshape(batch.ori, hw) != 0).any(0).nonzero().squeeze(1) + starts[l]
            sel = c1['cls_inds'][i]
            sel = sel[(sel >= starts[l]) & (sel < starts[l + 1])]
            assert torch.equal(rows, sel)
            # box gradient is non-zero only on positives and NMS survivors
            rows = (c1['g_box'][l][i].reshape(68, hw) != 0).any(0).nonzero().squeeze(1) + starts[l]
            allowed = torch.cat([(c1['gt_inds'][i] > 0).nonzero().squeeze(1), c1['box_inds'][i][c1['keep'][i]]])
            assert set(rows.tolist()) <= set(allowed.tolist())
    assert all(len(set(k.tolist())) == len(k) for k in c1['keep'])
    assert all(torch.equal(torch.sort(x)[0], x) for x in c1['cls_inds'] + c1['box_inds'])
    assert torch.equal(c1['loss_vec'], c3['loss_vec'])   # upstream does not change loss values


def test_accuracy_against_fp64_truth():
    """How far the fp32 CUDA result is from the same algorithm evaluated in fp64, next to how far
    the reference's own fp32 CPU result is: the CUDA path must be as accurate as the reference."""
    from oracle import erd_oracle as O
    batch = make_batch(2, (800, 1333), ori=40, seed=101)
    o, c = run_oracle(batch), run_cuda(batch)
    s_cls = [t.double().requires_grad_() for t in batch.s_cls]
    s_box = [t.double().requires_grad_() for t in batch.s_box]
    lo = O.loss_by_feat([t.double() for t in batch.t_cls], [t.double() for t in batch.t_box], s_cls, s_box,
                        o['cls_inds'], o['box_inds'], batch.ori, 1.0, batch.gt_bboxes, batch.gt_labels,
                        batch.pad_shapes)
    O.total_loss(lo).backward()
    for l in range(5):
        for key, truth in (('g_cls', s_cls[l].grad), ('g_box', s_box[l].grad)):
            scale = float(truth.abs().max())
            if scale == 0:
                continue
            e_cuda = float((c[key][l].double() - truth).abs().max()) / scale
            e_ref = float((o[key][l].double() - truth).abs().max()) / scale
            assert e_cuda <= max(3 * e_ref, 2e-6), f'{key}[{l}]: cuda {e_cuda:.2e} reference-fp32 {e_ref:.2e}'
    for k, v in lo.items():
        for x, y in zip(c['losses'][k], v):
            assert abs(x - float(y)) <= 2e-6 * max(abs(float(y)), 1e-7), k
