"""Developer diagnostic: per-kernel start/end times of one eager step relative to its start
(CUDA events on the launching streams) -- shows what overlaps and what is the critical path."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from erd_b200 import _native as N
from erd_b200.ops import ErdPath
from erd_b200.synth import make_batch
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 16
b = make_batch(n_img, (800, 1333), ori=40, seed=1234).to('cuda')
path = ErdPath(); lib = path.lib
p = path.plan(b.s_cls, 80, 40, 16); p.set_targets(b.gt_bboxes, b.gt_labels, b.pad_shapes)
g_cls = [torch.empty_like(t) for t in b.s_cls]; g_box = [torch.empty_like(t) for t in b.s_box]
losses = torch.empty(p.num_losses, device='cuda')
def step():
    path.prepare(p, b.t_cls, b.t_box, b.s_cls, b.s_box); path.reduce_avg(p)
    path.loss_fwd_bwd(p, b.t_cls, b.t_box, b.s_cls, b.s_box, g_cls, g_box, losses, 1.0)
for _ in range(5): step()
torch.cuda.synchronize()
nk = lib.erd_profile_num_kernels()
names = [lib.erd_profile_kernel_name(i).decode() for i in range(nk)]
lib.erd_profile_enable((1 << nk) - 1)
rows = {}
for it in range(5):
    torch.cuda.synchronize()
    lib.erd_profile_mark(torch.cuda.current_stream().cuda_stream)
    step()
    torch.cuda.synchronize()
    s, e = (C.c_float * nk)(), (C.c_float * nk)()
    lib.erd_profile_timeline(s, e)
    for i in range(nk):
        if s[i] >= 0: rows.setdefault(names[i], []).append((s[i] * 1e3, e[i] * 1e3))
    lib.erd_profile_collect(None, None)
lib.erd_profile_enable(0)
print(f'{"kernel":16s} {"start us":>9s} {"end us":>9s} {"dur":>7s}   (median of 5 eager steps, events add overhead)')
for k, v in sorted(rows.items(), key=lambda kv: sorted(x[0] for x in kv[1])[len(kv[1]) // 2]):
    st = sorted(x[0] for x in v)[len(v) // 2]; en = sorted(x[1] for x in v)[len(v) // 2]
    print(f'{k:16s} {st:9.1f} {en:9.1f} {en - st:7.1f}  ' + ' ' * int(st / 4) + '#' * max(1, int((en - st) / 4)))

print()
print('--- same step replayed from a CUDA graph (external event nodes), median of 5 replays')
lib.erd_profile_collect(None, None)
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side): step()
torch.cuda.current_stream().wait_stream(side)
gmask = int(os.environ.get('ERD_TL_MASK', '0'), 0) or (1 << nk) - 1
lib.erd_profile_enable(gmask)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    lib.erd_profile_mark(torch.cuda.current_stream().cuda_stream)
    step()
lib.erd_profile_enable(0)
rows = {}
for it in range(6):
    graph.replay(); torch.cuda.synchronize()
    s, e = (C.c_float * nk)(), (C.c_float * nk)()
    lib.erd_profile_timeline(s, e)
    if it == 0: continue
    for i in range(nk):
        if s[i] >= 0: rows.setdefault(names[i], []).append((s[i] * 1e3, e[i] * 1e3))
for k, v in sorted(rows.items(), key=lambda kv: sorted(x[0] for x in kv[1])[len(kv[1]) // 2]):
    st = sorted(x[0] for x in v)[len(v) // 2]; en = sorted(x[1] for x in v)[len(v) // 2]
    print(f'{k:16s} {st:9.1f} {en:9.1f} {en - st:7.1f}  ' + ' ' * int(st / 4) + '#' * max(1, int((en - st) / 4)))
