"""Dev: how many ERS anchors the teacher pass stashed (provisional thresholds vs the real ones)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erd_b200.ops import ErdPath
from erd_b200.synth import make_batch
mode = os.environ.get('MODE', 'gaussian'); n = int(os.environ.get('IMGS', 16))
hw = tuple(int(x) for x in os.environ.get('HW', '800x1333').split('x'))
b = make_batch(n, hw, ori=40, seed=1234, mode=mode, gt_size_pow=2.0).to('cuda')
path = ErdPath()
p, losses, gc, gb = path.step(b.t_cls, b.t_box, b.s_cls, b.s_box, b.gt_bboxes, b.gt_labels, b.pad_shapes, 80, 40, 16)
torch.cuda.synchronize()
slot = p.workspace_field('t_slot', torch.int16).view(n, -1).int() & 0xffff
pthr = p.workspace_field('pthr', torch.float32).view(n, 2)
sel = (p.sel_flags & 3) != 0
print('mode', mode, 'thr', p.thr[:3].tolist(), 'pthr', pthr[:3].tolist())
print('selected', int(sel.sum()), 'stashed', int((slot != 0).sum()), 'selected&stashed', int((sel & (slot != 0)).sum()),
      'hit rate', float((sel & (slot != 0)).sum()) / max(int(sel.sum()), 1), 'max rows/img', int(slot.max()))
