"""Dev: per-tile timestamps of one warp of CTA 0 of the teacher pass."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erd_b200 import _native as N
from erd_b200.ops import ErdPath
from erd_b200.synth import make_batch
hw = tuple(int(x) for x in os.environ.get('HW', '800x1333').split('x'))
b = make_batch(16, hw, ori=40, seed=1234).to('cuda')
path = ErdPath(); lib = N.load()
p = path.plan(b.s_cls, 80, 40, 16); p.set_targets(b.gt_bboxes, b.gt_labels, b.pad_shapes)
for _ in range(3): path.ers_select(p, b.t_cls, b.t_box)
torch.cuda.synchronize()
lib.erd_teacher_trace.argtypes = [C.c_void_p, C.c_int]
names = ['loop_top', 'full', 'scanned', 'fenced', 'counted', 'flag', 'extracted', 'requested']
for w in (0, 7):
    tr = torch.zeros(32 * 8, dtype=torch.int64, device='cuda')
    lib.erd_teacher_trace(tr.data_ptr(), w)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); path.ers_select(p, b.t_cls, b.t_box); e1.record(); torch.cuda.synchronize()
    lib.erd_teacher_trace(None, 0)
    t = tr.view(32, 8).cpu(); t0 = int(t[0, 0])
    print('warp', w, 'ers_select total ms', e0.elapsed_time(e1))
    print('tile ' + ' '.join(f'{n:>9s}' for n in names))
    for k in range(32):
        if int(t[k, 0]) == 0: break
        print(f'{k:4d} ' + ' '.join(f'{(int(x) - t0) / 1e3:9.2f}' for x in t[k]))
