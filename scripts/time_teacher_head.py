"""Timing of the teacher head-output producer fusion (erd_teacher_head_fused) against what it replaces:
cuDNN's two last head convolutions (TF32, as torch runs fp32 convolutions by default) + the streaming teacher pass.
BASELINE.json configs[1] geometry: 16 images, 800x1333 (padded 800x1344), 40 old classes.

    python scripts/time_teacher_head.py [--imgs 16] [--ori 40] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erd_b200 import _native as N   # noqa: E402
from erd_b200.ops import ErdPath, TeacherHead   # noqa: E402
from erd_b200.synth import level_shapes   # noqa: E402


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--imgs', type=int, default=16)
    ap.add_argument('--ori', type=int, default=40)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--hw', type=int, nargs=2, default=(800, 1344))
    ap.add_argument('--kernel-only', action='store_true', help='time the fused kernel only')
    args = ap.parse_args()
    dev = 'cuda'
    n, ori = args.imgs, args.ori
    shapes = level_shapes(*args.hw)
    A = sum(h * w for h, w in shapes)
    g = torch.Generator(device=dev).manual_seed(5)
    feat = lambda h, w: torch.randn(n, 256, h, w, device=dev, generator=g).relu_().contiguous(memory_format=torch.channels_last)
    cls_f = [feat(h, w) for h, w in shapes]
    reg_f = [feat(h, w) for h, w in shapes]
    w_cls = torch.randn(ori, 256, 3, 3, device=dev, generator=g) * 0.03
    w_reg = torch.randn(68, 256, 3, 3, device=dev, generator=g) * 0.03
    b_cls = torch.full((ori,), -4.6, device=dev)
    b_reg = torch.zeros(68, device=dev)
    scales = [1.0] * 5
    head = TeacherHead(w_cls, b_cls, w_reg, b_reg, scales)
    path = ErdPath()
    s_cls = [torch.empty(n, 80, h, w, device=dev) for h, w in shapes]
    p = path.plan(s_cls, 80, ori, 16)
    t_cls = [torch.empty(n, ori, h, w, device=dev) for h, w in shapes]
    t_box = [torch.empty(n, 68, h, w, device=dev) for h, w in shapes]
    out = {'imgs': n, 'ori': ori, 'anchors': n * A, 'tf32_conv_allowed': torch.backends.cudnn.allow_tf32}

    if args.kernel_only:
        out['fused_no_emit_ms'] = timed(lambda: path.teacher_head_fused(p, head, cls_f, reg_f), args.iters)
        print(json.dumps(out))
        return

    def cudnn_cl():
        for l in range(5):
            F.conv2d(cls_f[l], w_cls, b_cls, padding=1)
            F.conv2d(reg_f[l], w_reg, b_reg, padding=1).mul_(scales[l])
    out['cudnn_convs_channels_last_ms'] = timed(cudnn_cl, args.iters)
    cls_n = [f.contiguous() for f in cls_f]
    reg_n = [f.contiguous() for f in reg_f]

    def cudnn_nchw():
        for l in range(5):
            F.conv2d(cls_n[l], w_cls, b_cls, padding=1)
            F.conv2d(reg_n[l], w_reg, b_reg, padding=1).mul_(scales[l])
    out['cudnn_convs_nchw_ms'] = timed(cudnn_nchw, args.iters)
    del cls_n, reg_n
    out['teacher_pass_ers_select_ms'] = timed(lambda: path.ers_select(p, t_cls, t_box), args.iters)
    lib = N.load()
    names = [lib.erd_profile_kernel_name(i).decode() for i in range(lib.erd_profile_num_kernels())]
    out['fused_emit_ms'] = timed(lambda: path.teacher_head_fused(p, head, cls_f, reg_f, t_cls, t_box), args.iters)
    out['fused_no_emit_ms'] = timed(lambda: path.teacher_head_fused(p, head, cls_f, reg_f), args.iters)
    flops = 2.0 * n * A * 2304 * (ori + 68)
    out['fused_no_emit_useful_tflops'] = flops / out['fused_no_emit_ms'] / 1e9
    out['cudnn_cl_useful_tflops'] = flops / out['cudnn_convs_channels_last_ms'] / 1e9
    # the scan alone inside erd_ers_select
    lib.erd_profile_enable(1 << names.index('ers_scan'))
    for _ in range(5):
        path.ers_select(p, t_cls, t_box)
    torch.cuda.synchronize()
    import ctypes as C
    tot = (C.c_float * len(names))()
    cnt = (C.c_int * len(names))()
    lib.erd_profile_collect(tot, cnt)
    lib.erd_profile_enable(0)
    i = names.index('ers_scan')
    out['teacher_pass_scan_ms'] = tot[i] / max(cnt[i], 1)
    # ---- the whole teacher side + loss step, replayed from CUDA graphs (16 images of synthetic student outputs / GT)
    from erd_b200.synth import make_batch
    batch = make_batch(n, (args.hw[0], args.hw[1] - 11 if args.hw[1] == 1344 else args.hw[1]), ori=ori, seed=1234)
    assert batch.shapes == shapes, (batch.shapes, shapes)
    b = batch.to(dev)
    p.set_targets(b.gt_bboxes, b.gt_labels, b.pad_shapes)
    g_cls = [torch.empty_like(t) for t in b.s_cls]
    g_box = [torch.empty_like(t) for t in b.s_box]
    losses = torch.empty(p.num_losses, device=dev)

    def loss_step(cached):
        path.prepare(p, t_cls, t_box, b.s_cls, b.s_box, teacher_cached=cached)
        path.reduce_avg(p)
        path.loss_fwd_bwd(p, t_cls, t_box, b.s_cls, b.s_box, g_cls, g_box, losses, 1.0)

    def graphed(fn):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
        return timed(gr.replay, args.iters)

    path.teacher_head_fused(p, head, cls_f, reg_f, t_cls, t_box)   # real logits for the standard step
    out['loss_step_standard_ms'] = graphed(lambda: loss_step(False))
    l_std = losses.clone()

    def fused_then_step():
        path.teacher_head_fused(p, head, cls_f, reg_f, t_cls, t_box)
        loss_step(True)
    out['fused_emit_plus_loss_step_ms'] = graphed(fused_then_step)
    out['losses_identical_to_standard_step'] = bool(torch.equal(losses, l_std))

    def fused_noemit_then_step():
        path.teacher_head_fused(p, head, cls_f, reg_f)
        loss_step(True)
    out['fused_no_emit_plus_loss_step_ms'] = graphed(fused_noemit_then_step)

    def cudnn_then_step():
        for l in range(5):
            F.conv2d(cls_f[l], w_cls, b_cls, padding=1)
            F.conv2d(reg_f[l], w_reg, b_reg, padding=1).mul_(scales[l])
        loss_step(False)
    out['cudnn_convs_plus_loss_step_ms'] = graphed(cudnn_then_step)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
