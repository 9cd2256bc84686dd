"""Dev: per-tile timestamps of CTA 0 of the student pass."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erd_b200 import _native as N
from erd_b200.ops import ErdPath
from erd_b200.synth import make_batch
hw = tuple(int(x) for x in os.environ.get('HW', '800x1333').split('x'))
b = make_batch(16, hw, ori=40, seed=1234).to('cuda')
path = ErdPath(); lib = N.load()
p = path.plan(b.s_cls, 80, 40, 16); p.set_targets(b.gt_bboxes, b.gt_labels, b.pad_shapes)
g_cls = [torch.empty_like(t) for t in b.s_cls]; g_box = [torch.empty_like(t) for t in b.s_box]
losses = torch.empty(p.num_losses, device='cuda')
def step():
    path.prepare(p, b.t_cls, b.t_box, b.s_cls, b.s_box); path.reduce_avg(p)
    path.loss_fwd_bwd(p, b.t_cls, b.t_box, b.s_cls, b.s_box, g_cls, g_box, losses, 1.0)
lib.erd_student_dev(int(os.environ.get('DEV', 0)), int(os.environ.get('STAGES', 5)))
for _ in range(3): step()
torch.cuda.synchronize()
tr = torch.zeros(64 * 16, dtype=torch.int64, device='cuda')
lib.erd_student_trace.argtypes = [C.c_void_p]
lib.erd_student_trace(tr.data_ptr())
step(); torch.cuda.synchronize()
lib.erd_student_trace(None)
t = tr.view(64, 16).cpu()
t0 = int(t[0, 0])
names = ['ld_empty', 'ld_tma', 'ld_roles', 'ld_hdr', 'ld_stage', 'ld_arrive', 'c0_full', 'c0_dense', 'c0_items', 'c0_done', 'c15_full', 'c15_done', 'st_done', 'st_read']
order = [0, 8, 9, 10, 11, 1, 2, 12, 13, 3, 6, 7, 4, 5]
print('tile ' + ' '.join(f'{n:>9s}' for n in names))
for k in range(64):
    if int(t[k, 0]) == 0: break
    print(f'{k:4d} ' + ' '.join(f'{(int(t[k, j]) - t0) / 1e3:9.2f}' for j in order))
