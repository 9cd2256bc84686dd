"""Stage-level parity through the individual C-ABI entry points: ATSS on adversarial GT
boxes, NMS on dense random boxes, ERS edge cases."""
import pytest
import torch

from erd_b200.ops import ErdPath
from erd_b200.synth import make_batch
from oracle import erd_oracle as O

pytestmark = pytest.mark.gpu


def _atss_both(batch, path):
    b = batch.to('cuda')
    p = path.plan(b.s_cls, batch.num_classes, batch.ori, batch.reg_max)
    p.set_targets(b.gt_bboxes, b.gt_labels, b.pad_shapes)
    path.atss_assign(p)
    torch.cuda.synchronize()
    sizes = batch.shapes
    anchors = torch.cat([O.level_anchors(h, w, s) for (h, w), s in zip(sizes, O.STRIDES)])
    n_level = [h * w for h, w in sizes]
    out = []
    for i in range(batch.num_imgs):
        valid = torch.cat([O.level_valid_flags(h, w, s, *batch.pad_shapes[i]) for (h, w), s in zip(sizes, O.STRIDES)])
        rep = {}
        t = O.image_targets(anchors, valid, n_level, batch.gt_bboxes[i], batch.gt_labels[i], batch.num_classes, rep)
        out.append((p.gt_inds[i].cpu().long(), t['gt_inds'], int(p.num_pos[i]), t['num_pos'], rep))
    return out


def test_atss_adversarial_boxes():
    """Tiny boxes, boxes at the image border, huge boxes, boxes whose centre lies outside the
    valid lattice of a ragged image, boxes sharing anchors (conflict resolution)."""
    path = ErdPath()
    batch = make_batch(4, (512, 640), ori=40, seed=9, num_gt=1)
    W, H = 640.0, 512.0
    batch.gt_bboxes = [
        torch.tensor([[3.3, 4.1, 9.7, 8.2], [630.2, 500.1, 639.9, 511.7], [0.3, 0.2, 639.6, 511.1],
                      [100.3, 100.7, 140.1, 131.9], [101.9, 99.2, 143.4, 135.5]]),
        torch.tensor([[0.4, 200.3, 6.1, 260.9], [300.1, 0.2, 360.7, 5.9], [250.7, 250.3, 251.9, 251.4],
                      [10.1, 10.3, 600.7, 500.9], [12.7, 14.9, 598.2, 497.3], [15.2, 9.1, 602.3, 503.4]]),
        torch.tensor([[420.5, 300.2, 639.1, 511.3], [200.3, 150.9, 440.2, 350.1], [205.1, 148.2, 433.9, 352.6]]),
        torch.tensor([[17.77, 33.31, 48.13, 71.19], [47.1, 70.3, 90.9, 130.2], [500.5, 400.5, 520.25, 430.75]]),
    ]
    batch.gt_labels = [torch.arange(b.size(0)) % 40 for b in batch.gt_bboxes]
    batch.pad_shapes = [(512, 640), (512, 640), (384, 420), (512, 640)]
    for i, (g_cuda, g_or, np_c, np_o, rep) in enumerate(_atss_both(batch, path)):
        assert not rep.get('topk_boundary_ties'), f'image {i}: test input has a distance tie at the k-th boundary'
        assert torch.equal(g_cuda, g_or), f'image {i}'
        assert np_c == np_o


@pytest.mark.parametrize('seed', [1, 2, 3])
def test_atss_random_many_gt(seed):
    path = ErdPath()
    batch = make_batch(3, (800, 1333), ori=40, seed=200 + seed, num_gt=(20, 60), gt_size_pow=2.5)
    for i, (g_cuda, g_or, np_c, np_o, rep) in enumerate(_atss_both(batch, path)):
        if torch.equal(g_cuda, g_or):
            assert np_c == np_o
            continue
        # never skipped: a difference is acceptable only where the reference itself is arbitrary -- an exact
        # distance tie at the k-th candidate (torch.topk order) -- and only at anchors of the tied GTs
        ties = rep.get('topk_boundary_ties')
        assert ties, f'image {i}: assignment differs without a top-k tie (IoU-threshold margin {rep.get("min_thr_margin")})'
        diff = (g_cuda != g_or).nonzero().squeeze(1)
        print(f'image {i}: {diff.numel()} anchors differ; top-k boundary ties in the oracle: {ties[:4]}')
        assert diff.numel() <= 2 * len(ties)


def test_nms_dense_boxes_against_oracle_and_torchvision():
    """erd_teacher_nms on a planted teacher: every anchor of a 40x40 patch selected, heavy overlap,
    several classes -- checks sort order, class offsets and the bit-matrix scan."""
    tv = pytest.importorskip('torchvision')
    path = ErdPath()
    batch = make_batch(2, (320, 320), ori=40, seed=17, mode='trained')
    g = torch.Generator().manual_seed(3)
    for lv in range(2):   # a 16x16 (8x8) patch of confident, heavily overlapping teacher boxes in 4 classes
        n, _, h, w = batch.t_box[lv].shape
        side = 16 >> lv
        tb = batch.t_box[lv].view(n, 4, 17, h, w)
        patch = torch.randn(n, 4, 17, side, side, generator=g)
        bins = torch.randint(2, 17, (n, 4, 1, side, side), generator=g)
        tb[:, :, :, 4:4 + side, 4:4 + side] = patch + torch.zeros_like(patch).scatter_(2, bins, 7.0)
        batch.t_cls[lv][:, :4, 4:4 + side, 4:4 + side] += 4.0 * torch.rand(n, 4, side, side, generator=g)
    b = batch.to('cuda')
    p = path.plan(b.s_cls, batch.num_classes, batch.ori, batch.reg_max)
    p.set_targets(b.gt_bboxes, b.gt_labels, b.pad_shapes)
    path.ers_select(p, b.t_cls, b.t_box)
    path.teacher_nms(p)
    torch.cuda.synchronize()
    tc, tbx = O.flatten_levels(batch.t_cls), O.flatten_levels(batch.t_box)
    anchors = torch.cat([O.level_anchors(h, w, s) for (h, w), s in zip(batch.shapes, O.STRIDES)])
    ctr = torch.stack([(anchors[:, 0] + anchors[:, 2]) / 2, (anchors[:, 1] + anchors[:, 3]) / 2], -1)
    for i in range(batch.num_imgs):
        ci, bi = O.ers_select_single(tc[i], tbx[i])
        kb = int(p.box_count[i])
        assert torch.equal(p.box_inds[i, :kb].cpu().long(), bi)
        assert kb > 50
        boxes = O.points_to_box(ctr, O.integral(tbx[i]))
        conf, ids = tc[i].sigmoid().max(-1)
        _, keep = O.batched_nms(boxes[bi], conf[bi], ids[bi], dict(iou_threshold=0.005))
        got = p.keep[i, :int(p.keep_count[i])].cpu().long()
        assert torch.equal(got, keep)
        assert keep.numel() < bi.numel()          # NMS really suppressed something
        off = ids[bi].float() * (boxes[bi].max() + 1)
        assert torch.equal(keep, tv.ops.nms(boxes[bi] + off[:, None], conf[bi], 0.005))


def test_ers_constant_teacher_selects_nothing_and_distill_cls_is_nan():
    """Zero selected rows: torch.mean of an empty tensor is NaN in the reference
    (gfl_head_increment_erd.py:329-331); the CUDA path reproduces that, everything else stays finite."""
    from util import run_cuda, run_oracle
    batch = make_batch(1, (96, 128), ori=40, seed=5, num_gt=2, gt_size_pow=1.2)
    for t in batch.t_cls:
        t.fill_(-3.0)
    o, c = run_oracle(batch), run_cuda(batch)
    assert c['cls_inds'][0].numel() == 0 and o['cls_inds'][0].numel() == 0
    assert c['losses']['loss_dist_cls'][0] != c['losses']['loss_dist_cls'][0]   # NaN
    assert o['losses']['loss_dist_cls'][0] != o['losses']['loss_dist_cls'][0]
    assert all(x == x for k in ('loss_cls', 'loss_bbox', 'loss_dfl', 'loss_dist_bbox') for x in c['losses'][k])
    assert all(bool(torch.isfinite(g).all()) for g in c['g_cls'] + c['g_box'])


def _staged(path, b, ctx_none=False, preclear=False, poison=False):
    """The path through the individual stage entry points (all on the caller's stream), then the
    loss call -- with the context (helper streams) or with ctx = NULL (everything serial)."""
    import ctypes as C
    from erd_b200 import _native as N
    from erd_b200.ops import _ptrs, _stream
    p = path.plan(b.s_cls, b.num_classes, b.ori, b.reg_max)
    p.set_targets(b.gt_bboxes, b.gt_labels, b.pad_shapes)
    fill = float('nan') if poison else 0.0
    g_cls = [torch.full_like(t, fill) for t in b.s_cls]
    g_box = [torch.full_like(t, fill) for t in b.s_box]
    losses = torch.empty(p.num_losses, device='cuda')
    if preclear:
        path.prepare(p, b.t_cls, b.t_box, b.s_cls, b.s_box)
    else:
        path.ers_select(p, b.t_cls, b.t_box)
        path.atss_assign(p)
        path.avg_factors(p, b.s_cls, b.s_box)
        path.teacher_nms(p)
    path.reduce_avg(p)
    if ctx_none:
        N.check(path.lib.erd_loss_fwd_bwd(
            None, C.byref(p.shape), _ptrs(b.s_cls), _ptrs(b.s_box), _ptrs(b.t_cls), _ptrs(b.t_box),
            p.gt_boxes.data_ptr(), p.gt_labels.data_ptr(), p.gt_offsets.data_ptr(), p.pad_hw.data_ptr(),
            p.gt_inds.data_ptr(), p.num_pos.data_ptr(), p.cls_inds.data_ptr(), p.cls_count.data_ptr(),
            p.sel_flags.data_ptr(), p.box_inds.data_ptr(), p.box_count.data_ptr(), p.keep.data_ptr(),
            p.keep_count.data_ptr(), p.avg.data_ptr(), 1.0, None, 0, losses.data_ptr(), _ptrs(g_cls), _ptrs(g_box),
            p.ws.data_ptr(), _stream()), 'erd_loss_fwd_bwd')
    else:
        path.loss_fwd_bwd(p, b.t_cls, b.t_box, b.s_cls, b.s_box, g_cls, g_box, losses, 1.0)
    torch.cuda.synchronize()
    return p, losses, g_cls, g_box


@pytest.mark.parametrize('mode', ['staged_ctx', 'staged_null_ctx', 'prepare'])
def test_every_call_sequence_gives_the_same_bits(mode):
    """The fused step (erd_step_prepare + loss), the stage-by-stage sequence
    and the NULL-context single-stream sequence run the same kernels in different orders on
    different streams: identical index lists, losses and gradients, and every gradient element
    is written (buffers are poisoned with NaN first)."""
    path = ErdPath()
    batch = make_batch(3, (480, 640), ori=40, seed=77, num_gt=[4, 0, 7], mode='trained', gt_size_pow=2.0,
                       pad_shapes=[(480, 640), (300, 500), (480, 400)])
    b = batch.to('cuda')
    ref_p, ref_l, ref_gc, ref_gb = path.step(b.t_cls, b.t_box, b.s_cls, b.s_box, b.gt_bboxes, b.gt_labels,
                                             b.pad_shapes, b.num_classes, b.ori, b.reg_max)[:4]
    torch.cuda.synchronize()
    ref_l, ref_gc, ref_gb = ref_l.clone(), [t.clone() for t in ref_gc], [t.clone() for t in ref_gb]
    ref_keep = ref_p.keep.clone(), ref_p.keep_count.clone()
    p, l, gc, gb = _staged(path, b, ctx_none=(mode == 'staged_null_ctx'), preclear=(mode == 'prepare'),
                           poison=True)
    assert torch.equal(p.keep_count, ref_keep[1])
    for i in range(batch.num_imgs):
        k = int(p.keep_count[i])
        assert torch.equal(p.keep[i, :k], ref_keep[0][i, :k])
    assert torch.equal(l.isnan(), ref_l.isnan())
    assert float((l - ref_l).nan_to_num().abs().max()) <= 2e-6 * float(ref_l.nan_to_num().abs().max())
    for a, r in zip(gc + gb, ref_gc + ref_gb):
        assert not a.isnan().any()              # every element written
        assert torch.equal(a, r)                # gradients do not depend on the launch order


def test_cuda_graph_of_the_step_survives_new_targets():
    """A CUDA graph captured over prepare + loss with one batch's GT must stay valid when the next batch's GT
    (different boxes, different count) is loaded into the plan: GT buffers and launch geometry are fixed-capacity."""
    path = ErdPath()
    a = make_batch(3, (512, 640), ori=40, seed=61, num_gt=[2, 9, 4]).to('cuda')
    b2 = make_batch(3, (512, 640), ori=40, seed=62, num_gt=[7, 0, 11]).to('cuda')
    p = path.plan(a.s_cls, a.num_classes, a.ori, a.reg_max, 16)
    g_cls = [torch.empty_like(t) for t in a.s_cls]
    g_box = [torch.empty_like(t) for t in a.s_box]
    losses = torch.empty(p.num_losses, device='cuda')

    def step():
        path.prepare(p, a.t_cls, a.t_box, a.s_cls, a.s_box)
        path.loss_fwd_bwd(p, a.t_cls, a.t_box, a.s_cls, a.s_box, g_cls, g_box, losses, 1.0)
    p.set_targets(a.gt_bboxes, a.gt_labels, a.pad_shapes)
    step()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    p.set_targets(b2.gt_bboxes, b2.gt_labels, b2.pad_shapes)   # new GT, same tensors otherwise
    graph.replay()
    torch.cuda.synchronize()
    got = (losses.clone(), p.gt_inds.clone(), [t.clone() for t in g_cls + g_box])
    step()                                                     # the same thing launched eagerly
    torch.cuda.synchronize()
    assert torch.equal(got[1], p.gt_inds) and int((p.gt_inds > 0).sum()) > 0
    assert torch.equal(got[0], losses)
    for x, y in zip(got[2], g_cls + g_box):
        assert torch.equal(x, y)


def test_stash_hits_and_misses_give_the_same_bits():
    """The teacher pass stashes by provisional thresholds left by the PREVIOUS call; whatever they are -- none (first
    call), well matched (same distribution), too low (flooding the stash) or too high (nothing stashed, every ERS
    column gathered from the tensors) -- the step's results must be bit-identical to a fresh plan's."""
    import torch
    iid = make_batch(2, (480, 640), ori=40, seed=71, mode='gaussian', gt_size_pow=2.0).to('cuda')
    planted = make_batch(2, (480, 640), ori=40, seed=72, mode='trained', gt_size_pow=2.0).to('cuda')

    def run(path, b):
        p, losses, g_cls, g_box = path.step(b.t_cls, b.t_box, b.s_cls, b.s_box, b.gt_bboxes, b.gt_labels, b.pad_shapes,
                                            b.num_classes, b.ori, b.reg_max)
        torch.cuda.synchronize()
        slot = p.workspace_field('t_slot', torch.int16).view(b.num_imgs, -1).int() & 0xffff
        sel = (p.sel_flags & 3) != 0
        hit = float((sel & (slot != 0)).sum()) / max(int(sel.sum()), 1)
        return (losses.clone(), [t.clone() for t in g_cls + g_box], p.sel_flags.clone() & 3, p.keep_count.clone()), hit

    fresh = {name: run(ErdPath(), b)[0] for name, b in (('iid', iid), ('planted', planted))}
    path = ErdPath()
    hits = []
    for name, b in (('iid', iid), ('iid', iid), ('planted', planted), ('planted', planted), ('iid', iid), ('iid', iid)):
        got, hit = run(path, b)
        hits.append(round(hit, 3))
        want = fresh[name]
        assert torch.equal(got[0], want[0]), (name, hits)
        assert all(torch.equal(x, y) for x, y in zip(got[1], want[1])), (name, hits)
        assert torch.equal(got[2], want[2]) and torch.equal(got[3], want[3])
    print('stash hit rates over the sequence:', hits)
    assert hits[0] == 0.0 and hits[1] == 1.0      # nothing to go by on the first call, a perfect estimate on the second
    assert hits[4] < 0.5 and hits[5] == 1.0       # planted -> iid: the estimate is too high once, then it has adapted
