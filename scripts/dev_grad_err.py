"""Diagnostic: distribution of gradient errors (CUDA vs fp32 oracle vs fp64 oracle)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from erd_b200.synth import make_batch
from util import run_cuda, run_oracle
from oracle import erd_oracle as O

b = make_batch(2, (800, 1333), ori=40, seed=101)
o = run_oracle(b)
c = run_cuda(b)
# fp64 "truth": same index sets, double arithmetic
b64 = make_batch(2, (800, 1333), ori=40, seed=101)
s_cls = [t.double().requires_grad_() for t in b64.s_cls]; s_box = [t.double().requires_grad_() for t in b64.s_box]
t_cls = [t.double() for t in b64.t_cls]; t_box = [t.double() for t in b64.t_box]
try:
    lo = O.loss_by_feat(t_cls, t_box, s_cls, s_box, o['cls_inds'], o['box_inds'], b.ori, 1.0,
                        b.gt_bboxes, b.gt_labels, b.pad_shapes)
    O.total_loss(lo).backward()
    have64 = True
except Exception as e:
    print('fp64 oracle failed', repr(e)); have64 = False
for l in range(5):
    for key, ref64 in (('g_cls', s_cls), ('g_box', s_box)):
        a, r = c[key][l].double(), o[key][l].double()
        scale = r.abs().max().item()
        if scale == 0: continue
        line = f'{key}[{l}] scale {scale:.3e} max|cuda-o32|/scale {((a-r).abs().max()/scale):.2e}'
        for fl in (1e-3, 1e-2, 1e-1):
            e = ((a - r).abs() / r.abs().clamp(min=fl * scale)).max().item()
            line += f' el@{fl:g} {e:.2e}'
        if have64:
            t = ref64[l].grad
            line += f' | vs64: cuda {((a-t).abs().max()/scale):.2e} o32 {((r-t).abs().max()/scale):.2e}'
            for fl in (1e-3, 1e-2):
                line += f' el64@{fl:g} cuda {((a - t).abs() / t.abs().clamp(min=fl * scale)).max().item():.2e} o32 {((r - t).abs() / t.abs().clamp(min=fl * scale)).max().item():.2e}'
        print(line)
# linearity check
from erd_b200.ops import ErdPath
p = ErdPath()
c1 = run_cuda(b, path=p)
c3 = run_cuda(b, path=p, upstream=torch.full((15 + 4,), 3.0, device='cuda'))
for l in range(5):
    for key in ('g_cls', 'g_box'):
        x, y = c3[key][l], 3.0 * c1[key][l]
        d = (x - y).abs()
        i = d.argmax()
        print(key, l, 'max abs diff', d.max().item(), 'at vals', x.reshape(-1)[i].item(), y.reshape(-1)[i].item(),
              'max rel', (d / y.abs().clamp(min=1e-30)).max().item())
