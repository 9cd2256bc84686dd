"""One data-parallel training step of the incremental GFL detector with the plugin on the loss path
(BASELINE.json configs[2]: "40+40 ERD end-to-end training step data-parallel ..., NCCL reduce_mean + grad
allreduce"; reference sequence: GFLIncrementERD.loss, mmdet/models/detectors/gfl_increment_erd.py:202-220).

R50 + FPN + GFL towers in plain PyTorch/cuDNN (random init; north_star: the conv stacks are not rewritten), a
frozen 40-class teacher, an 80-class student wrapped in DistributedDataParallel (bucketed gradient all-reduce
overlapping the conv backward), the loss path through `GFLIncrementERD.sel_pos` + `GFLHeadIncrementERD.loss`
(C ABI; its 8-byte avg-factor exchange is the path's only own collective), SGD step.

    python scripts/train_step_ddp.py [imgs_per_gpu] [bf16] [fuse] [channels_last]           # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        --master-port 29511 scripts/train_step_ddp.py [imgs_per_gpu] [bf16]                 # N GPUs

Rank 0 prints one JSON line: images/s over all ranks (CUDA events, max over ranks) and the phase split."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
from erd_b200.detector import GFLIncrementERD           # noqa: E402
from erd_b200.head import GFLHeadIncrementERD, fused_teacher_head, parse_losses   # noqa: E402
from erd_b200.synth import make_gt                       # noqa: E402
from time_train_step import R50FPN, Teacher              # noqa: E402


class Student(nn.Module):
    def __init__(self):
        super().__init__()
        self.body = R50FPN()
        self.head = GFLHeadIncrementERD(80, 256)

    def forward(self, x):
        return self.head(self.body(x))


def main():
    args = [a for a in sys.argv[1:]]
    n = int(args[0]) if args and args[0].isdigit() else 16
    amp = 'bf16' in args
    fuse = 'fuse' in args   # the teacher's last head convolutions inside the teacher pass
    rank, local, world = int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(0)
    student = Student().to(dev).train()
    model = nn.parallel.DistributedDataParallel(student, device_ids=[local], bucket_cap_mb=25) if world > 1 else student
    det = GFLIncrementERD(student.head, 40, ori_model=Teacher().to(dev).eval(), extract_feat=None)
    cl = 'channels_last' in args   # run the frozen teacher NHWC end to end: its tower outputs are then what the fused head streams
    if cl:
        det.ori_model.to(memory_format=torch.channels_last)
    params = [p for p in student.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=1e-4, momentum=0.9, weight_decay=1e-4)
    x = torch.randn(n, 3, 800, 1344, device=dev)
    x_t = x.contiguous(memory_format=torch.channels_last) if cl else x   # the teacher's view of the batch
    rng = np.random.RandomState(rank)

    class DS:
        def __init__(self):
            b, l = make_gt(rng, int(rng.randint(1, 10)), 800, 1333, 40)
            self.gt_instances = type('GT', (), dict(bboxes=b.to(dev), labels=l.to(dev)))()
            self.metainfo = dict(img_shape=(800, 1333), pad_shape=(800, 1344))
    samples = [DS() for _ in range(n)]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    times = []
    warm, steps = 3, 6
    for it in range(warm + steps):
        opt.zero_grad(set_to_none=True)
        ev[0].record()
        if fuse:
            with torch.no_grad():
                with torch.autocast('cuda', torch.bfloat16, enabled=amp):
                    feats = det.ori_model.body(x_t)
                ori_outs, (cls_sel, box_sel), _ = fused_teacher_head(student.head.path, det.ori_model.head, feats, 80, 16)
            ev[1].record()
            sel = (cls_sel, None, box_sel, None)
        else:
            with torch.no_grad(), torch.autocast('cuda', torch.bfloat16, enabled=amp):
                ori_outs = det.ori_model(x_t)
            ori_outs = ([t.float().contiguous() for t in ori_outs[0]], [t.float().contiguous() for t in ori_outs[1]])
            ev[1].record()
            sel = det.sel_pos(*ori_outs)                                    # gfl_increment_erd.py:207-209
        with torch.autocast('cuda', torch.bfloat16, enabled=amp):
            new_outs = model(x)                                         # :210-211 (through DDP)
        new_outs = ([t.float() for t in new_outs[0]], [t.float() for t in new_outs[1]])
        ev[2].record()
        losses = student.head.loss(ori_outs, new_outs, samples, *sel, 40, 1, det)   # :213-219
        total = parse_losses(losses)
        ev[3].record()
        total.backward()                                                # conv backward + bucketed all-reduce
        ev[4].record()
        opt.step()
        ev[5].record()
        torch.cuda.synchronize()
        if it >= warm:
            times.append([ev[i].elapsed_time(ev[i + 1]) for i in range(5)])
    t = torch.tensor(np.median(np.array(times), axis=0), device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t = t.cpu().numpy()
    step_ms = float(t.sum())
    if rank == 0:
        print(json.dumps({
            'metric': 'images_per_sec_train_step', 'value': world * n / (step_ms * 1e-3), 'n_gpus': world,
            'images_per_gpu': n, 'ms_per_step': round(step_ms, 3),
            'conv_precision': 'bf16 autocast' if amp else 'fp32 (TF32 convs)', 'teacher_head_fused': fuse, 'teacher_channels_last': cl,
            'ms': {'teacher_forward': round(float(t[0]), 3), 'sel_pos_plus_student_forward': round(float(t[1]), 3),
                   'erd_loss_path_fwd_bwd': round(float(t[2]), 3), 'conv_backward_plus_grad_allreduce': round(float(t[3]), 3),
                   'sgd_step': round(float(t[4]), 3)},
            'loss_value': float(total.detach()), 'data': 'synthetic, random-init R50-FPN (torchvision trunk)'}), flush=True)
    if world > 1:
        sys.stdout.flush()
        try:
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os._exit(0)


if __name__ == '__main__':
    main()
