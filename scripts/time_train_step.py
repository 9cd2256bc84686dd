"""Context for the bench line: what one training step of the incremental GFL detector costs
around the loss path (BASELINE.json north_star: "the R50 backbone, FPN and head convs stay on
PyTorch/cuDNN and are timed separately rather than rewritten").

R50 + FPN (P3..P7, 256 ch) + the GFL conv towers in plain PyTorch, random init, a frozen
40-class teacher and an 80-class student; the loss path is the product (`GFLIncrementERD.loss`
-> C ABI).  Phases are timed with CUDA events on one B200, 16 images 800x1344:

    teacher forward (no_grad) | student backbone+FPN+head forward | ERS + loss fwd/bwd (ours)
    | backward through the conv stacks

Prints one JSON line.  Not part of the tests or of bench.py; needs torchvision.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from erd_b200.detector import GFLIncrementERD           # noqa: E402
from erd_b200.head import GFLHeadIncrementERD, parse_losses   # noqa: E402
from erd_b200.synth import make_gt                       # noqa: E402


class R50FPN(nn.Module):
    """torchvision ResNet-50 trunk (BN frozen, as norm_eval=True in the configs) + the FPN of
    configs/gfl_increment (start_level=1, add_extra_convs='on_output', num_outs=5)."""

    def __init__(self):
        super().__init__()
        import torchvision
        r = torchvision.models.resnet50(weights=None)
        self.stem = nn.Sequential(r.conv1, r.bn1, r.relu, r.maxpool)
        self.layers = nn.ModuleList([r.layer1, r.layer2, r.layer3, r.layer4])
        self.lateral = nn.ModuleList([nn.Conv2d(c, 256, 1) for c in (512, 1024, 2048)])
        self.out = nn.ModuleList([nn.Conv2d(256, 256, 3, padding=1) for _ in range(3)])
        self.extra = nn.ModuleList([nn.Conv2d(256, 256, 3, stride=2, padding=1) for _ in range(2)])
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()
                for p in m.parameters():
                    p.requires_grad = False

    def train(self, mode=True):
        super().train(mode)
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()
        return self

    def forward(self, x):
        x = self.stem(x)
        feats = []
        for i, layer in enumerate(self.layers):
            x = layer(x)
            if i >= 1:
                feats.append(x)
        lat = [l(f) for l, f in zip(self.lateral, feats)]
        for i in (2, 1):
            lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[-2:], mode='nearest')
        outs = [o(l) for o, l in zip(self.out, lat)]
        for e in self.extra:
            outs.append(e(outs[-1]))
        return outs


class Teacher(nn.Module):
    def __init__(self):
        super().__init__()
        self.body = R50FPN()
        self.head = GFLHeadIncrementERD(40, 256)

    def forward(self, x):
        return self.head(self.body(x))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    amp = len(sys.argv) > 2 and sys.argv[2] == 'bf16'
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    student_body = R50FPN().to(dev).train()
    head = GFLHeadIncrementERD(80, 256).to(dev)
    det = GFLIncrementERD(head, 40, ori_model=Teacher().to(dev).eval(), extract_feat=student_body)
    params = [p for p in list(student_body.parameters()) + list(head.parameters()) if p.requires_grad]
    x = torch.randn(n, 3, 800, 1344, device=dev)
    rng = np.random.RandomState(0)

    class DS:
        def __init__(self):
            b, l = make_gt(rng, int(rng.randint(1, 10)), 800, 1333, 40)
            self.gt_instances = type('GT', (), dict(bboxes=b.to(dev), labels=l.to(dev)))()
            self.metainfo = dict(img_shape=(800, 1333), pad_shape=(800, 1344))
    samples = [DS() for _ in range(n)]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    times = []
    for it in range(6):
        for p in params:
            p.grad = None
        ev[0].record()
        with torch.no_grad(), torch.autocast('cuda', torch.bfloat16, enabled=amp):
            ori_outs = det.ori_model(x)
        ori_outs = ([t.float() for t in ori_outs[0]], [t.float() for t in ori_outs[1]])
        ev[1].record()
        with torch.autocast('cuda', torch.bfloat16, enabled=amp):
            new_outs = head(student_body(x))
        new_outs = ([t.float() for t in new_outs[0]], new_outs[1])
        ev[2].record()
        # the product path: ERS selection + fused loss forward/backward through the C ABI; the
        # gradients of the head outputs exist when this returns, autograd only hands them over
        sel = det.sel_pos(*ori_outs)
        losses = head.loss(ori_outs, new_outs, samples, *sel, 40, 1, det)
        total = parse_losses(losses)
        ev[3].record()
        total.backward()
        ev[4].record()
        torch.cuda.synchronize()
        if it >= 2:
            times.append([ev[i].elapsed_time(ev[i + 1]) for i in range(4)])
    t = np.median(np.array(times), axis=0)
    anchors = n * 22400
    print(json.dumps({
        'images': n, 'anchors': anchors, 'conv_precision': 'bf16 autocast' if amp else 'fp32 (TF32 convs)',
        'ms': {'teacher_forward': round(float(t[0]), 3), 'student_forward': round(float(t[1]), 3),
               'erd_loss_path_fwd_bwd': round(float(t[2]), 3), 'conv_backward': round(float(t[3]), 3)},
        'loss_path_share_of_step': round(float(t[2] / t.sum()), 5),
        'loss_value': float(total.detach()),
        'note': 'loss path timed eagerly through the Python plugin API (includes its host-side launch cost)'}))


if __name__ == '__main__':
    main()
