#!/usr/bin/env python
"""Benchmark of the GFL+ERD loss path (BASELINE.json metric: anchors/s, loss fwd+bwd).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port)

One "step" = ERS selection + ATSS assignment + avg-factor all-reduce + teacher NMS + fused
QFL/GIoU/DFL/distillation forward and backward over one batch of synthetic head outputs
(BASELINE.json configs[1]: 40+40 split, 16 images/GPU, 800x1333 -> 22 400 anchors/image).
Prints ONE JSON line on rank 0.  Under torchrun every rank owns its own 16 images (weak
scaling); the only collective is the 8-byte avg-factor all-reduce (NCCL).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'anchors_per_sec_gfl_erd_loss_fwd_bwd'
UNIT = 'anchors/s'
NUM_CLASSES, REG_MAX = 80, 16
# --config: the BASELINE.json configurations that fit one GPU (configs[1] is the one the metric is quoted on
# and the default; the others are reported in DESIGN.md section 5)
CONFIGS = {
    'cfg1': dict(hw=(800, 1333), ori=40, imgs=16, num_gt=None, mode='gaussian',
                 name='GFL R50-FPN 40+40 ERD loss fwd+bwd, {n} img/GPU 800x1333, 80-class head, reg_max=16 '
                      '(BASELINE.json configs[1])'),
    'cfg4_70_10': dict(hw=(800, 1333), ori=70, imgs=16, num_gt=None, mode='gaussian',
                       name='70+10 split ERD loss fwd+bwd, {n} img/GPU 800x1333, 80-class head, 70 old classes distilled '
                            '(BASELINE.json configs[3])'),
    'cfg5_dense': dict(hw=(1600, 1600), ori=40, imgs=8, num_gt=100, mode='trained',
                       name='dense scene 40+40 ERD loss fwd+bwd, {n} img/GPU 1600x1600, 100 GT boxes/image, planted-object '
                            'teacher (BASELINE.json configs[4])'),
    'trained': dict(hw=(800, 1333), ori=40, imgs=16, num_gt=None, mode='trained',
                    name='40+40 ERD loss fwd+bwd, {n} img/GPU 800x1333, planted-object teacher (NMS-heavy mode of '
                         'SURVEY.md 8(d))'),
}
CFG = CONFIGS['cfg1']


def bytes_per_anchor(ori):
    """compulsory fp32 traffic per anchor, fwd+bwd.  teacher pass: reads the
    teacher's old-class logits and box distributions once; student pass: reads every student logit once and
    writes every gradient element once."""
    teacher, student = 4 * (ori + 68), 8 * (NUM_CLASSES + 68)
    return {'path': teacher + student, 'ers_scan': teacher, 'student_pass': student}


# Per-launch figures of the two dense kernels run one at a time under ncu at the default workload (cfg1, 16
# images, each on the 124 SMs it gets inside the step): duration and DRAM bytes (dram__bytes_read.sum +
# dram__bytes_write.sum) from the --set full captures in profiles/r2_ncu_full_student_teacher.csv; the launch list
# of a whole bench run is profiles/r2_launches.csv.  A kernel
# that writes shows less than its algorithmic bytes: dirty lines still sit in the 126 MB L2 when it ends.
NCU_ALONE_US = {'ers_scan': 47.8, 'student_pass': 93.2}
NCU_TRAFFIC_BYTES = {'ers_scan': 161.8e6, 'student_pass': 368.5e6}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='cfg1', choices=sorted(CONFIGS))
    ap.add_argument('--imgs', type=int, default=0, help='images per GPU (default: the config\'s)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-teacher-head', action='store_true', help='skip the fused teacher-head side measurement')
    return ap.parse_args()


def make_inputs(n_imgs, seed):
    from erd_b200.synth import make_batch
    return make_batch(n_imgs, CFG['hw'], ori=CFG['ori'], num_classes=NUM_CLASSES, reg_max=REG_MAX, seed=seed,
                      num_gt=CFG['num_gt'], mode=CFG['mode'], gt_size_pow=2.0 if CFG['num_gt'] else 1.0)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / throttle-reason samples taken through NVML while the timed region runs
    (2 ms period, so even a 20 ms region gets samples; nvidia-smi -lms cannot go that fast)."""

    def __init__(self, index):
        self.index, self.sm, self.mx, self.reasons, self.err = index, [], None, set(), None
        self._stop = threading.Event()
        self._thr = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {'hw_slowdown': 0x8, 'sw_power_cap': 0x4, 'hw_thermal_slowdown': 0x40,
                     'sw_thermal_slowdown': 0x20, 'hw_power_brake_slowdown': 0x80}

            def loop():
                while not self._stop.is_set():
                    try:
                        self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for nm, bit in names.items():
                            if r & bit:
                                self.reasons.add(nm)
                    except Exception as e:  # keep sampling
                        self.err = repr(e)
                    time.sleep(0.002)
            self._thr = threading.Thread(target=loop, daemon=True)
            self._thr.start()
        except Exception as e:
            self.err = repr(e)

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=1.0)
        out = {'sm_mhz': statistics.median(self.sm) if self.sm else None, 'sm_max_mhz': self.mx,
               'samples': len(self.sm), 'reasons': sorted(self.reasons)}
        if self.err:
            out['error'] = self.err
        return out


# ----------------------------------------------------------------------------- CPU reference arm
def time_cpu_oracle(n_imgs, iters, warmup, seed=1234, budget_s=None):
    """The oracle port (torch CPU, all host threads) on n_imgs images of the workload per step.
    ``budget_s``: stop timing early (after at least one step) when the run would exceed it."""
    from oracle import erd_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    b = make_inputs(n_imgs, seed)
    times = []
    t_start = time.perf_counter()
    for it in range(warmup + iters):
        s_cls = [t.clone().requires_grad_() for t in b.s_cls]
        s_box = [t.clone().requires_grad_() for t in b.s_box]
        t0 = time.perf_counter()
        O.erd_step(b.t_cls, b.t_box, s_cls, s_box, b.gt_bboxes, b.gt_labels, b.pad_shapes, b.ori,
                   1.0, b.num_classes, b.reg_max)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        if budget_s is not None and times and time.perf_counter() - t_start + dt > budget_s:
            break
    anchors = n_imgs * b.anchors_per_image
    return anchors, times, torch.get_num_threads()


def run_reference(args):
    """The reference's CPU implementation of the path (the oracle port: the reference itself is Python on
    mmcv/mmengine and cannot travel to the GPU box) on the SAME workload as our arm: all images of the config
    per step, the steps and warm-up the driver asked for, every host core.  Rank 0 only."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    n = args.imgs
    anchors, times, cores = time_cpu_oracle(n, args.steps, args.warmup, budget_s=420.0)
    steps = len(times)
    ms = 1e3 * sum(times) / steps
    val = anchors / (sum(times) / steps)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': CFG['name'].format(n=n), 'images_per_step': n,
                   'note': 'one CPU process on rank 0 (the whole host), not one per GPU'
                           + ('' if steps == args.steps else f'; stopped after {steps} of {args.steps} steps (time budget)')},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': f'{n} images x {anchors // n} anchors per step, {steps} steps after {args.warmup} '
                                   f'warm-up, torch-CPU oracle port on {cores} threads'},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm
def time_teacher_head(path, plan, b, g_cls, g_box, losses, n, A, ori, steps):
    """The teacher side + loss step with the last head convolutions (a) in cuDNN followed by the standard step and
    (b) inside erd_teacher_head_fused followed by the step with ERD_PREPARE_TEACHER_CACHED.  Synthetic tower
    features (post-ReLU N(0,1), NHWC) and head weights; CUDA graph replay, CUDA events."""
    import torch.nn.functional as F
    from erd_b200.ops import TeacherHead
    dev = b.s_cls[0].device
    gen = torch.Generator(device=dev).manual_seed(77)
    shapes = [tuple(t.shape[2:]) for t in b.s_cls]
    feat = lambda h, w: torch.randn(n, 256, h, w, device=dev, generator=gen).relu_().contiguous(memory_format=torch.channels_last)
    cls_f, reg_f = [feat(h, w) for h, w in shapes], [feat(h, w) for h, w in shapes]
    w_cls = torch.randn(ori, 256, 3, 3, device=dev, generator=gen) * 0.03
    w_reg = torch.randn(68, 256, 3, 3, device=dev, generator=gen) * 0.03
    b_cls, b_reg = torch.full((ori,), -4.6, device=dev), torch.zeros(68, device=dev)
    head = TeacherHead(w_cls, b_cls, w_reg, b_reg, [1.0] * 5)
    t_cls = [torch.empty(n, ori, h, w, device=dev) for h, w in shapes]
    t_box = [torch.empty(n, 68, h, w, device=dev) for h, w in shapes]

    def loss_step(cached):
        path.prepare(plan, t_cls, t_box, b.s_cls, b.s_box, teacher_cached=cached)
        path.reduce_avg(plan)
        path.loss_fwd_bwd(plan, t_cls, t_box, b.s_cls, b.s_box, g_cls, g_box, losses, 1.0)

    def cudnn_convs():
        for l in range(5):
            t_cls[l] = F.conv2d(cls_f[l], w_cls, b_cls, padding=1)
            t_box[l] = F.conv2d(reg_f[l], w_reg, b_reg, padding=1).mul_(1.0)

    def timed(fn, iters):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    def graphed(fn, iters):
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
        return timed(gr.replay, iters)

    iters = max(10, min(steps, 50))
    out = {'what': 'teacher head last convs (256 -> ori / 68, 3x3) + teacher pass + loss step, one GPU, '
                   f'{n} images x {A} anchors, tower features NHWC fp32 synthetic; CUDA graph replay',
           'math': 'tcgen05 kind::tf32, fp32 accumulate (cuDNN arm: torch default allow_tf32 for convolutions)'}
    out['fused_kernel_ms'] = timed(lambda: path.teacher_head_fused(plan, head, cls_f, reg_f, t_cls, t_box), iters)
    out['fused_kernel_no_logits_ms'] = timed(lambda: path.teacher_head_fused(plan, head, cls_f, reg_f), iters)
    out['cudnn_convs_ms'] = timed(cudnn_convs, iters)
    flops = 2.0 * n * A * 2304 * (ori + 68)
    out['fused_kernel_useful_tflops'] = flops / out['fused_kernel_no_logits_ms'] / 1e9
    out['cudnn_convs_useful_tflops'] = flops / out['cudnn_convs_ms'] / 1e9
    try:
        peak_bf16 = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('bf16_tflops', 1632.0))
    except Exception:
        peak_bf16 = 1632.0
    npad = ((ori + 15) // 16) * 16 + 80          # columns the MMAs compute (N padded to legal UMMA shapes)
    tiles = sum(-(-h // 16) * -(-w // 16) for h, w in shapes) * n   # 16 x 16 pixel patches incl. the ragged edges
    executed = 2.0 * tiles * 256 * 2304 * npad
    out['roofline'] = {'bound': 'tensor', 'unit': 'TFLOP/s', 'achieved': flops / out['fused_kernel_no_logits_ms'] / 1e9,
                       'executed_incl_padding': executed / out['fused_kernel_no_logits_ms'] / 1e9,
                       'peak': peak_bf16 / 2, 'peak_source': 'MEASURED_PEAKS.json bf16_tflops / 2 (kind::tf32 issues at half the '
                       'bf16 rate: K = 8 instead of 16 per MMA)',
                       'frac': flops / out['fused_kernel_no_logits_ms'] / 1e9 / (peak_bf16 / 2),
                       'note': f'N = {npad - 80} / 80 per MMA: each MMA fetches a 4 KB A tile from shared memory for '
                               f'{(npad - 80) // 2} / 40 cycles of math, so shared-memory operand bandwidth, not the tensor pipe, '
                               'bounds it (DESIGN 4b)'}
    t_cls = [torch.empty(n, ori, h, w, device=dev) for h, w in shapes]
    t_box = [torch.empty(n, 68, h, w, device=dev) for h, w in shapes]
    path.teacher_head_fused(plan, head, cls_f, reg_f, t_cls, t_box)

    def std_nocopy():
        for l in range(5):
            F.conv2d(cls_f[l], w_cls, b_cls, padding=1)
            F.conv2d(reg_f[l], w_reg, b_reg, padding=1).mul_(1.0)
        loss_step(False)

    def fused():
        path.teacher_head_fused(plan, head, cls_f, reg_f, t_cls, t_box)
        loss_step(True)
    out['cudnn_convs_plus_step_ms'] = graphed(std_nocopy, iters)
    loss_step(False)
    torch.cuda.synchronize()
    l_std = losses.clone()
    out['fused_plus_step_ms'] = graphed(fused, iters)
    torch.cuda.synchronize()
    out['losses_bit_identical_to_standard_step_on_the_emitted_logits'] = bool(torch.equal(losses, l_std))
    out['speedup_teacher_side_plus_step'] = out['cudnn_convs_plus_step_ms'] / out['fused_plus_step_ms']
    return out


def run_ours(args):
    if os.environ.get('ERD_BENCH_DEBUG'):
        import faulthandler
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        _fh = open(os.path.join(ROOT, 'gpurun_out', f"hang_rank{os.environ.get('RANK', 0)}.txt"), 'w')
        faulthandler.dump_traceback_later(int(os.environ['ERD_BENCH_DEBUG']), exit=True, file=_fh)
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    import ctypes as C
    from erd_b200 import _native as N
    from erd_b200.head import GFLHeadIncrementERD, parse_losses
    from erd_b200.detector import GFLIncrementERD

    lib = N.load()
    head = GFLHeadIncrementERD(NUM_CLASSES, 256, reg_max=REG_MAX, build_convs=False,
                               train_cfg=dict(assigner=dict(type='ATSSAssigner', topk=9), allowed_border=-1,
                                              pos_weight=-1))
    path = head.path
    n = args.imgs
    host = make_inputs(n, 1234 + rank)
    A = host.anchors_per_image
    b = host.to(dev)
    ORI = CFG['ori']
    BYTES_PER_ANCHOR = bytes_per_anchor(ORI)
    plan = path.plan(b.s_cls, NUM_CLASSES, ORI, REG_MAX, max((int(x.shape[0]) for x in b.gt_bboxes), default=0))
    plan.set_targets(b.gt_bboxes, b.gt_labels, b.pad_shapes)
    g_cls = [torch.empty_like(t) for t in b.s_cls]
    g_box = [torch.empty_like(t) for t in b.s_box]
    losses = torch.empty(plan.num_losses, device=dev)

    def step():
        # the two C-ABI calls around the 8-byte all-reduce; grads written into fixed buffers
        path.prepare(plan, b.t_cls, b.t_box, b.s_cls, b.s_box)
        path.reduce_avg(plan)
        path.loss_fwd_bwd(plan, b.t_cls, b.t_box, b.s_cls, b.s_box, g_cls, g_box, losses, 1.0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    nk = lib.erd_profile_num_kernels()
    names = [lib.erd_profile_kernel_name(i).decode() for i in range(nk)]
    dense = [k for k in ('ers_scan', 'student_pass') if k in names]

    def collect():
        tot, cnt = (C.c_float * nk)(), (C.c_int * nk)()
        lib.erd_profile_collect(tot, cnt)
        return {names[i]: (tot[i] / cnt[i] if cnt[i] else 0.0) for i in range(nk)}

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    barrier()
    # (1) eager launches, K steps, CUDA events around the three dense kernels only (roofline line)
    lib.erd_profile_enable(sum(1 << names.index(k) for k in dense))
    launches0 = lib.erd_launch_count()
    ms_eager = timed(step, args.steps)
    launches = lib.erd_launch_count() - launches0
    lib.erd_profile_enable(0)
    dense_ms = {k: v for k, v in collect().items() if k in dense}
    dom = max(dense_ms, key=dense_ms.get)
    dom_ms = dense_ms[dom]
    # (2) the same K steps replayed from one CUDA graph (no per-launch CPU cost): the headline value
    sampler = ClockSampler(local)
    ms_graph, graph_err = None, None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        for _ in range(warm):
            graph.replay()
        if rank == 0:
            sampler.start()
        ms_graph = timed(graph.replay, args.steps)
    except Exception as e:  # fall back to the eager number, say why
        graph_err = repr(e)[:200]
        if rank == 0:
            sampler.start()
        ms_eager = timed(step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ms_graph if ms_graph is not None else ms_eager
    ms_step = ms_total / args.steps
    value = world * n * A / (ms_step * 1e-3)
    # (3) per-kernel breakdown (all kernels bracketed by events; informational, not the timed run)
    lib.erd_profile_enable((1 << nk) - 1)
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    lib.erd_profile_enable(0)
    kern = collect()

    # ---- the same step through the plugin API with DEVICE-resident inputs: sel_pos + loss_by_feat + backward
    # (what a training loop pays per iteration on top of the conv stacks; eager launches, autograd included)
    api = None
    if not args.no_e2e:
        det_a = GFLIncrementERD(head, ORI)
        gts_a = [type('GT', (), dict(bboxes=x, labels=y))() for x, y in zip(b.gt_bboxes, b.gt_labels)]
        metas_a = [dict(img_shape=i, pad_shape=p) for i, p in zip(host.img_shapes, host.pad_shapes)]
        sc = [t.clone().requires_grad_() for t in b.s_cls]
        sb = [t.clone().requires_grad_() for t in b.s_box]

        def api_step():
            for t in sc + sb:
                t.grad = None
            sel = det_a.sel_pos(b.t_cls, b.t_box)
            out = head.loss_by_feat((b.t_cls, b.t_box), (sc, sb), sel[0], sel[1], sel[2], sel[3], ORI, 1.0, None,
                                    gts_a, metas_a)
            parse_losses(out).backward()
        for _ in range(3):
            api_step()
        ms_api = timed(api_step, args.steps) / args.steps
        api = {'ms_per_step': ms_api, 'value': world * n * A / (ms_api * 1e-3), 'unit': UNIT,
               'api': 'GFLIncrementERD.sel_pos + GFLHeadIncrementERD.loss_by_feat + backward, inputs resident in HBM, '
                      'eager launches (no CUDA graph), CUDA events'}

    # ---- multi-rank check (N > 1): the reduced avg factors every rank holds must be the mean of the ranks'
    # local factors as reduce_mean computes it (t / W summed in rank order), bit for bit on every rank
    multi = None
    if world > 1:
        path.prepare(plan, b.t_cls, b.t_box, b.s_cls, b.s_box)
        local = plan.avg.clone()
        path.reduce_avg(plan)
        path.loss_fwd_bwd(plan, b.t_cls, b.t_box, b.s_cls, b.s_box, g_cls, g_box, losses, 1.0)
        torch.cuda.synchronize()
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        want = torch.zeros_like(local)
        for t in gathered:
            want += t / world
        ok = bool(torch.equal(plan.avg, want)) and bool(torch.isfinite(losses).all())
        okt = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        multi = {'avg_factors_equal_mean_of_ranks_bitwise': bool(int(okt.item())),
                 'local_avg_rank0': [float(x) for x in local.tolist()], 'reduced_avg': [float(x) for x in plan.avg.tolist()]}
        if not int(okt.item()):
            raise SystemExit(f'rank {rank}: reduced avg factors {plan.avg.tolist()} != mean of ranks {want.tolist()}')

    # ---- e2e: the reference-facing plugin API with HOST buffers (pinned), H2D of every head
    # output and D2H of the loss vector inside the timed region, autograd backward included.
    e2e = None
    if not args.no_e2e:
        det = GFLIncrementERD(head, ORI)
        pin = lambda ts: [t.pin_memory() for t in ts]
        h = dict(t_cls=pin(host.t_cls), t_box=pin(host.t_box), s_cls=pin(host.s_cls), s_box=pin(host.s_box))
        d = {k: [torch.empty_like(t, device=dev) for t in v] for k, v in h.items()}
        gts = [type('GT', (), dict(bboxes=x, labels=y))() for x, y in zip(b.gt_bboxes, b.gt_labels)]
        metas = [dict(img_shape=i, pad_shape=p) for i, p in zip(host.img_shapes, host.pad_shapes)]
        h2d = sum(t.numel() * 4 for v in h.values() for t in v)
        loss_host = torch.empty(plan.num_losses).pin_memory()

        def e2e_step():
            with torch.no_grad():
                for k in h:
                    for src, dst in zip(h[k], d[k]):
                        dst.copy_(src, non_blocking=True)
            s_cls = [t.requires_grad_() for t in d['s_cls']]
            s_box = [t.requires_grad_() for t in d['s_box']]
            for t in s_cls + s_box:
                t.grad = None
            sel = det.sel_pos(d['t_cls'], d['t_box'])
            out = head.loss_by_feat((d['t_cls'], d['t_box']), (s_cls, s_box), sel[0], sel[1], sel[2], sel[3],
                                    ORI, 1.0, None, gts, metas)
            parse_losses(out).backward()
            vec = torch.stack(out['loss_cls'] + out['loss_bbox'] + out['loss_dfl'] + out['loss_dist_cls']
                              + out['loss_dist_bbox']).detach()
            loss_host.copy_(vec, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        e_steps = max(3, min(args.steps, 10))
        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {'value': world * n * A * e_steps / float(dt.item()), 'unit': UNIT, 'h2d_bytes_per_step': h2d,
               'd2h_bytes_per_step': plan.num_losses * 4, 'steps': e_steps,
               # the copies are what this number measures: per-rank host->device rate over the whole step (N ranks share
               # the host's memory and PCIe root complexes, which is why e2e does not scale like `value`)
               'h2d_gbs_per_rank': h2d * e_steps / float(dt.item()) / 1e9,
               'api': 'GFLIncrementERD.sel_pos + GFLHeadIncrementERD.loss_by_feat + backward, pinned host tensors'}

    # ---- SURVEY 8(f) rank 1, measured beside the headline (not part of `value`): the teacher head's last convolutions
    # fused with the teacher pass (erd_teacher_head_fused, tcgen05 TF32) against cuDNN's convolutions + the standard
    # step, teacher side + loss step replayed from one CUDA graph each; same student tensors / GT as the headline.
    fused_head = None
    if world == 1 and not args.no_teacher_head:
        try:
            fused_head = time_teacher_head(path, plan, b, g_cls, g_box, losses, n, A, ORI, args.steps)
        except Exception as e:   # never lose the headline over the side measurement
            fused_head = {'error': repr(e)[:300]}

    if world > 1:
        from erd_b200.dist_utils import peer_exchange
        collective = ('8-byte avg-factor mean: one kernel over NVLink peer memory (erd_avg_exchange)'
                      if peer_exchange(lib, dev) is not None else '8-byte avg-factor mean: ncclAllReduce')
    else:
        collective = 'none (1 rank)'
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        achieved = (n * A * BYTES_PER_ANCHOR[dom]) / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else None
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': warm, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': CFG['name'].format(n=n), 'name': args.config, 'anchors_per_image': A,
                       'images_per_gpu': n, 'parallelism': f'dp{world} over images', 'collective': collective,
                       'l2': f'inputs+grads {n * A * BYTES_PER_ANCHOR["path"] / 1e6:.0f} MB per step > 126 MB L2, no flush needed'},
            'roofline': {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': (achieved / peak) if achieved else None,
                         'traffic': NCU_TRAFFIC_BYTES.get(dom) if (args.config == 'cfg1' and n == CONFIGS['cfg1']['imgs']) else None,
                         'traffic_source': 'profiles/r2_ncu_full_student_teacher.csv (ncu --set full, per launch, bytes)',
                         'kernel_alone': ({'us': NCU_ALONE_US[dom],
                                           'achieved_gbs': round(n * A * BYTES_PER_ANCHOR[dom] / (NCU_ALONE_US[dom] * 1e-6) / 1e9, 1),
                                           'frac': round(n * A * BYTES_PER_ANCHOR[dom] / (NCU_ALONE_US[dom] * 1e-6) / 1e9 / peak, 3),
                                           'source': 'profiles/r2_ncu_full_student_teacher.csv (ncu, kernel run alone, cold cache)'}
                                          if (args.config == 'cfg1' and n == CONFIGS['cfg1']['imgs']) and dom in NCU_ALONE_US else None),
                         'peak_source': 'measured' if peaks else 'fallback',
                         'algorithmic_bytes_per_anchor': BYTES_PER_ANCHOR[dom],
                         'path_bytes_per_anchor': BYTES_PER_ANCHOR['path'],
                         'path_achieved_gbs': n * A * BYTES_PER_ANCHOR['path'] / (ms_step * 1e-3) / 1e9,
                         'kernel_ms_dominant_in_timed_eager_run': round(dom_ms, 5),
                         'dense_kernels_in_timed_eager_run': {
                             k: {'ms': round(v, 5), 'algorithmic_bytes_per_anchor': BYTES_PER_ANCHOR[k],
                                 'achieved_gbs': round(n * A * BYTES_PER_ANCHOR[k] / (v * 1e-3) / 1e9, 1) if v else None}
                             for k, v in dense_ms.items()},
                         'kernel_ms_breakdown_pass': {k: round(v, 5) for k, v in kern.items() if v}},
            'launch_mode': {'value_from': 'cuda_graph_replay' if ms_graph is not None else 'eager',
                            'ms_per_step_eager': ms_eager / args.steps, 'graph_error': graph_err},
            'gpu_launches': int(launches), 'clocks': clocks, 'e2e': e2e, 'api_device_resident': api,
            'multi_rank_check': multi, 'teacher_head_fused': fused_head,
        }
        if not args.no_cpu_baseline and world == 1:
            anchors, times, cores = time_cpu_oracle(n, 10, 2, budget_s=25.0)
            line['cpu_baseline'] = {'value': anchors / (sum(times) / len(times)), 'unit': UNIT, 'cores': cores,
                                    'kind': 'port', 'sample': f'{n} images x {A} anchors per step (the whole workload), '
                                                              f'{len(times)} steps after 2 warm-up, torch-CPU oracle port '
                                                              f'of the reference on {cores} threads'}
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL teardown with a captured graph alive can block; everything is printed, so leave
        # through a barrier and a hard exit instead of destroy_process_group()
        sys.stdout.flush()
        try:
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os._exit(0)


if __name__ == '__main__':
    a = parse()
    CFG = CONFIGS[a.config]
    if a.imgs <= 0:
        a.imgs = CFG['imgs']
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
