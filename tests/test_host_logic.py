"""Host-side logic that needs no GPU: geometry, sharding, reduce_mean over gloo (world 2),
config validation and error behaviour of the reference-facing classes."""
import os
import subprocess
import sys

import pytest
import torch

from erd_b200.dist_utils import shard_images
from erd_b200.synth import level_shapes, make_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_level_shapes_match_survey():
    assert level_shapes(800, 1344) == [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
    assert level_shapes(1600, 1600) == [(200, 200), (100, 100), (50, 50), (25, 25), (13, 13)]
    b = make_batch(1, (800, 1333))
    assert b.anchors_per_image == 22400 and b.canvas == (800, 1344)


def test_synth_is_deterministic_and_labels_in_new_class_range():
    a, b = make_batch(2, (96, 128), seed=3, ori=70), make_batch(2, (96, 128), seed=3, ori=70)
    assert all(torch.equal(x, y) for x, y in zip(a.s_cls + a.t_box, b.s_cls + b.t_box))
    assert all(int(l.max()) < 10 and int(l.min()) >= 0 for l in a.gt_labels if l.numel())
    assert a.t_cls[0].shape[1] == 70 and a.s_cls[0].shape[1] == 80


def test_shard_images_partitions_the_batch():
    for world in (1, 2, 3, 8):
        parts = [shard_images(16, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(16))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_head_config_validation_mirrors_reference_contract():
    from erd_b200.head import GFLHeadIncrementERD
    ok = GFLHeadIncrementERD(80, 256, build_convs=False)
    assert ok.strides == (8, 16, 32, 64, 128) and ok.reg_max == 16
    with pytest.raises(ValueError):
        GFLHeadIncrementERD(80, 256, build_convs=False, loss_cls=dict(type='FocalLoss'))
    with pytest.raises(ValueError):
        GFLHeadIncrementERD(80, 256, build_convs=False, loss_cls=dict(type='QualityFocalLoss', beta=1.0))
    with pytest.raises(ValueError):
        GFLHeadIncrementERD(80, 256, build_convs=False, train_cfg=dict(assigner=dict(type='ATSSAssigner', topk=5)))
    with pytest.raises(AssertionError):   # reference asserts square strides, gfl_head_increment_erd.py:256
        GFLHeadIncrementERD(80, 256, build_convs=False, anchor_generator=dict(
            type='AnchorGenerator', ratios=[1.0], octave_base_scale=8, scales_per_octave=1,
            strides=[(8, 16), 16, 32, 64, 128]))


def test_product_path_refuses_cpu_tensors():
    """No CPU fallback: the public API raises instead of computing on the host."""
    from erd_b200.head import GFLHeadIncrementERD
    head = GFLHeadIncrementERD(80, 256, build_convs=False)
    b = make_batch(1, (96, 128))
    gts = [type('G', (), dict(bboxes=x, labels=y))() for x, y in zip(b.gt_bboxes, b.gt_labels)]
    with pytest.raises(RuntimeError, match='CUDA'):
        head.loss_by_feat((b.t_cls, b.t_box), (b.s_cls, b.s_box), None, None, None, None, 40, 1.0, None, gts,
                          [dict(pad_shape=p) for p in b.pad_shapes])


def test_product_does_not_import_the_oracle():
    """Nothing under erd_b200/ may import oracle/ (the oracle is test infrastructure)."""
    pkg = os.path.join(ROOT, 'erd_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert 'import oracle' not in src and 'from oracle' not in src, fn


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from erd_b200.dist_utils import reduce_mean_, shard_images, world
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:' + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, ws = world()
# each rank owns its shard of 5 images; avg = [sum max(num_pos,1), sum w] of its shard
num_pos = [3, 0, 7, 1, 0]
wsum = [0.5, 0.0, 1.25, 0.125, 0.0]
mine = shard_images(5, rank, ws)
avg = torch.tensor([float(sum(max(num_pos[i], 1) for i in mine)), float(sum(wsum[i] for i in mine))])
reduce_mean_(avg)
exp0 = sum(max(p, 1) for p in num_pos) / 2.0
exp1 = sum(wsum) / 2.0
assert abs(float(avg[0]) - exp0) < 1e-6 and abs(float(avg[1]) - exp1) < 1e-6, (avg, exp0, exp1)
# the NVLink peer-memory exchange is a CUDA + NCCL mechanism: on this group it must decline
# (the caller then keeps the all-reduce above) instead of trying to map device memory
from erd_b200.dist_utils import PeerAvgExchange
assert PeerAvgExchange.create(None, torch.device('cpu')) is None
dist.destroy_process_group()
print('ok', rank)
'''


def test_reduce_mean_world_size_two_gloo(tmp_path):
    """The N>1 host path: one 2-float all-reduce, mean over ranks (dist_utils.py:59-65)."""
    script = tmp_path / 'worker.py'
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 1000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all('ok' in o for o in outs)


def test_peer_exchange_declines_without_a_process_group():
    from erd_b200.dist_utils import PeerAvgExchange
    import torch
    assert PeerAvgExchange.create(None, torch.device('cpu')) is None


def test_checkpoint_surgery_expands_the_cls_head():
    """gfl_increment_erd.py:67-93: the teacher checkpoint (40-class cls conv) seeds the 80-class student; the 40 new
    rows keep the student's own initialisation; a 'module.' prefix (DDP checkpoint) is stripped."""
    import torch
    from collections import OrderedDict
    from erd_b200.detector import GFLIncrementERD
    from erd_b200.head import GFLHeadIncrementERD
    torch.manual_seed(0)
    teacher_head = GFLHeadIncrementERD(40, 16, stacked_convs=1, feat_channels=16,
                                       norm_cfg=dict(type='GN', num_groups=4, requires_grad=True))
    student_head = GFLHeadIncrementERD(80, 16, stacked_convs=1, feat_channels=16,
                                       norm_cfg=dict(type='GN', num_groups=4, requires_grad=True))
    det = GFLIncrementERD(student_head, 40)
    own_w, own_b = student_head.gfl_cls.weight.detach().clone(), student_head.gfl_cls.bias.detach().clone()
    ckpt = dict(state_dict=OrderedDict(('module.bbox_head.' + k, v.clone()) for k, v in teacher_head.state_dict().items()))
    missing, unexpected = det.load_checkpoint_for_new_model(ckpt, strict=True)
    assert not missing and not unexpected
    assert torch.equal(student_head.gfl_cls.weight[:40], teacher_head.gfl_cls.weight)
    assert torch.equal(student_head.gfl_cls.bias[:40], teacher_head.gfl_cls.bias)
    assert torch.equal(student_head.gfl_cls.weight[40:], own_w[40:]) and torch.equal(student_head.gfl_cls.bias[40:], own_b[40:])
    assert torch.equal(student_head.gfl_reg.weight, teacher_head.gfl_reg.weight)
    with pytest.raises(RuntimeError):
        det.load_checkpoint_for_new_model(dict(weights=1))


def test_fused_teacher_head_host_checks():
    """The fused teacher head without a GPU: argument checking of the C ABI entry points happens on
    the host before any launch, the packed-weight size is host arithmetic, and the Python wrapper has no CPU path."""
    import ctypes as C
    from erd_b200 import _native as N
    from erd_b200.ops import TeacherHead
    lib = N.load()
    assert lib.erd_teacher_head_packed_floats(40) == 48 * 256 * 9      # 40 classes padded to a legal UMMA N
    assert lib.erd_teacher_head_packed_floats(68) == 80 * 256 * 9
    assert lib.erd_teacher_head_packed_floats(0) == 0
    assert lib.erd_teacher_head_pack(None, 40, None, None) == -2         # ERD_ERR_NULL
    shape = N.ErdShape(num_imgs=2, num_levels=5, num_classes=80, ori_classes=40, reg_max=16, total_gt=0, anchor_scale=8.0,
                       loss_weight_cls=1.0, loss_weight_bbox=2.0, loss_weight_dfl=0.25, loss_weight_ld=0.25,
                       kd_temperature=10.0, max_gt_per_img=0)
    for l, (h, w) in enumerate(level_shapes(256, 320)):
        shape.level_h[l], shape.level_w[l], shape.stride[l] = h, w, 8 << l
    assert lib.erd_teacher_head_fused(C.byref(shape), None, N.PtrArray(), N.PtrArray(), None, None, None, None, None,
                                      None) == -2
    assert b'NULL' in lib.erd_last_error()
    assert lib.erd_ers_select_cached(C.byref(shape), None, None, None, None, None, None, None, None) == -2
    w = torch.zeros(40, 256, 3, 3)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        TeacherHead(w, torch.zeros(40), torch.zeros(68, 256, 3, 3), torch.zeros(68), [1.0] * 5)
