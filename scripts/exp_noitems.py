"""Dev experiment: student pass time with no special columns at all (constant teacher -> empty ERS sets, no GT)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erd_b200 import _native as N
from erd_b200.ops import ErdPath
from erd_b200.synth import make_batch
b = make_batch(16, (800, 1333), ori=40, seed=1234, num_gt=int(os.environ.get('NGT', 0))).to('cuda')
if os.environ.get('CONST_TEACHER', '1') == '1':
    for t in b.t_cls + b.t_box: t.fill_(0.25)
path = ErdPath(); lib = N.load()
p = path.plan(b.s_cls, 80, 40, 16); p.set_targets(b.gt_bboxes, b.gt_labels, b.pad_shapes)
g_cls = [torch.empty_like(t) for t in b.s_cls]; g_box = [torch.empty_like(t) for t in b.s_box]
losses = torch.empty(p.num_losses, device='cuda')
def step():
    path.prepare(p, b.t_cls, b.t_box, b.s_cls, b.s_box); path.reduce_avg(p)
    path.loss_fwd_bwd(p, b.t_cls, b.t_box, b.s_cls, b.s_box, g_cls, g_box, losses, 1.0)
for _ in range(5): step()
torch.cuda.synchronize()
nk = lib.erd_profile_num_kernels(); names = [lib.erd_profile_kernel_name(i).decode() for i in range(nk)]
lib.erd_profile_enable((1 << nk) - 1)
for _ in range(10): step()
torch.cuda.synchronize(); lib.erd_profile_enable(0)
tot, cnt = (C.c_float * nk)(), (C.c_int * nk)(); lib.erd_profile_collect(tot, cnt)
print('selected', int(p.cls_count.sum()), int(p.box_count.sum()), 'positives', int(p.num_pos.sum()),
      {names[i]: round(1e3 * tot[i] / cnt[i], 1) for i in range(nk) if cnt[i] and names[i] in ('ers_scan', 'student_pass')})
