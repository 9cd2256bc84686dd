"""Developer diagnostic: uncontended per-kernel times, each stage launched alone in a loop."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from erd_b200.ops import ErdPath
from erd_b200.synth import make_batch
mode = sys.argv[1] if len(sys.argv) > 1 else 'gaussian'
b = make_batch(16, (800, 1333), ori=40, seed=1234, mode=mode).to('cuda')
path = ErdPath(); lib = path.lib
p = path.plan(b.s_cls, 80, 40, 16); p.set_targets(b.gt_bboxes, b.gt_labels, b.pad_shapes)
g_cls = [torch.empty_like(t) for t in b.s_cls]; g_box = [torch.empty_like(t) for t in b.s_box]
losses = torch.empty(p.num_losses, device='cuda')
nk = lib.erd_profile_num_kernels(); names = [lib.erd_profile_kernel_name(i).decode() for i in range(nk)]
stages = [('ers_select', lambda: path.ers_select(p, b.t_cls, b.t_box)), ('atss', lambda: path.atss_assign(p)),
          ('avg', lambda: path.avg_factors(p, b.s_cls, b.s_box)), ('nms', lambda: path.teacher_nms(p)),
          ]
for _, fn in stages: fn()
torch.cuda.synchronize()
print('counts', p.box_count.tolist()[:4], p.keep_count.tolist()[:4], p.num_pos.tolist()[:4])
for name, fn in stages:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    lib.erd_profile_enable((1 << nk) - 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    lib.erd_profile_enable(0)
    tot, cnt = (C.c_float * nk)(), (C.c_int * nk)()
    lib.erd_profile_collect(tot, cnt)
    ks = ', '.join(f'{names[i]} {1e3 * tot[i] / cnt[i]:.1f}' for i in range(nk) if cnt[i])
    print(f'{name:10s} wall {1e3 * e0.elapsed_time(e1) / 20:7.1f} us/iter | kernels (us): {ks}')

print('--- CUDA-graph replay of each stage alone (no CPU launch cost), us per replay')
for name, fn in stages:
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f'{name:10s} {1e3 * e0.elapsed_time(e1) / 50:7.1f}')
