"""Dev: how many ERS anchors the teacher pass stashed.  The provisional thresholds come from the PREVIOUS call, so
the script runs a sequence of different batches through one plan and reports the hit rate of each."""
import os, struct, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erd_b200.ops import ErdPath
from erd_b200.synth import make_batch
n = int(os.environ.get('IMGS', 16))
hw = tuple(int(x) for x in os.environ.get('HW', '800x1333').split('x'))
seq = os.environ.get('SEQ', 'gaussian:1,gaussian:2,gaussian:2,trained:3,trained:4,gaussian:5').split(',')
path = ErdPath()
def decode(v):
    if v == 0: return None
    k = (~v) & 0xffffffff
    u = (k & 0x7fffffff) if (k & 0x80000000) else ((~k) & 0xffffffff)
    return struct.unpack('f', struct.pack('I', u))[0]
for item in seq:
    mode, seed = item.split(':')
    b = make_batch(n, hw, ori=40, seed=int(seed), mode=mode, gt_size_pow=2.0).to('cuda')
    p = path.plan(b.s_cls, 80, 40, 16)
    before = [decode(int(v) & 0xffffffff) for v in p.workspace_field('pthr_state', torch.int32).cpu().tolist()[:2]]
    p, losses, gc, gb = path.step(b.t_cls, b.t_box, b.s_cls, b.s_box, b.gt_bboxes, b.gt_labels, b.pad_shapes, 80, 40, 16)
    torch.cuda.synchronize()
    slot = p.workspace_field('t_slot', torch.int16).view(n, -1).int() & 0xffff
    sel = (p.sel_flags & 3) != 0
    print(f'{mode}:{seed}  provisional (from the previous call) {before}  real thr (min over images) '
          f'{[round(float(x), 4) for x in p.thr.min(0)[0].tolist()]}  selected {int(sel.sum())} stashed {int((slot != 0).sum())} '
          f'hit rate {float((sel & (slot != 0)).sum()) / max(int(sel.sum()), 1):.3f}')
