"""Dev: time the step and its kernels for several tuning variants in ONE process."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erd_b200 import _native as N
from erd_b200.ops import ErdPath
from erd_b200.synth import make_batch
n = int(os.environ.get('IMGS', 16))
hw = tuple(int(x) for x in os.environ.get('HW', '800x1333').split('x'))
b = make_batch(n, hw, ori=int(os.environ.get('ORI', 40)), seed=1234, mode=os.environ.get('MODE', 'gaussian')).to('cuda')
print('image', hw, 'levels', b.shapes, 'anchors', b.anchors_per_image)
path = ErdPath(); lib = N.load()
p = path.plan(b.s_cls, 80, b.ori, 16); p.set_targets(b.gt_bboxes, b.gt_labels, b.pad_shapes)
g_cls = [torch.empty_like(t) for t in b.s_cls]; g_box = [torch.empty_like(t) for t in b.s_box]
losses = torch.empty(p.num_losses, device='cuda')
def step():
    path.prepare(p, b.t_cls, b.t_box, b.s_cls, b.s_box); path.reduce_avg(p)
    path.loss_fwd_bwd(p, b.t_cls, b.t_box, b.s_cls, b.s_box, g_cls, g_box, losses, 1.0)
nk = lib.erd_profile_num_kernels(); names = [lib.erd_profile_kernel_name(i).decode() for i in range(nk)]
for rep in range(int(os.environ.get('REPS', 1))):
    for _ in range(5): step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): step()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g): step()
    for _ in range(3): g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): g.replay()
    e1.record(); torch.cuda.synchronize()
    graph_ms = e0.elapsed_time(e1) / 20
    lib.erd_profile_enable((1 << nk) - 1)
    for _ in range(10): step()
    torch.cuda.synchronize()
    lib.erd_profile_enable(0)
    tot, cnt = (C.c_float * nk)(), (C.c_int * nk)(); lib.erd_profile_collect(tot, cnt)
    print(f'graph_step_ms={graph_ms:.4f}', {names[i]: round(1e3 * tot[i] / cnt[i], 1) for i in range(nk) if cnt[i]}, flush=True)
