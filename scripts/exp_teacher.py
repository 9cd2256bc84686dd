"""Dev experiment: the teacher chain (erd_ers_select = teacher pass + flags + lists) timed back to back (L2 holds its
own lines), after a 0.85 GB device copy (L2 full of someone else's dirty lines, as after the student pass), and the
student-side chain alone."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erd_b200.ops import ErdPath
from erd_b200.synth import make_batch
b = make_batch(16, (800, 1333), ori=40, seed=1234).to('cuda')
path = ErdPath(); p = path.plan(b.s_cls, 80, 40, 16); p.set_targets(b.gt_bboxes, b.gt_labels, b.pad_shapes)
big_a = torch.empty(106 * 1024 * 1024, dtype=torch.float32, device='cuda'); big_b = torch.empty_like(big_a)
def timed(fn, pre=None, n=20):
    tot = 0.0
    for _ in range(n):
        if pre: pre()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n * 1e3
sel = lambda: path.ers_select(p, b.t_cls, b.t_box)
for _ in range(3): sel()
print('teacher chain, back to back       : %.1f us' % timed(sel))
print('teacher chain, after a 424 MB copy: %.1f us' % timed(sel, pre=lambda: big_b.copy_(big_a)))
att = lambda: (path.atss_assign(p), path.avg_factors(p, b.s_cls, b.s_box))
for _ in range(3): att()
print('atss + avg chain alone            : %.1f us' % timed(att))
