"""Dev: where the host time of one plugin-API step goes (sel_pos + loss_by_feat + backward, device-resident inputs)."""
import cProfile, os, pstats, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erd_b200.head import GFLHeadIncrementERD, parse_losses
from erd_b200.detector import GFLIncrementERD
from erd_b200.synth import make_batch
host = make_batch(16, (800, 1333), ori=40, seed=1234)
b = host.to('cuda')
head = GFLHeadIncrementERD(80, 256, reg_max=16, build_convs=False)
det = GFLIncrementERD(head, 40)
gts = [type('GT', (), dict(bboxes=x, labels=y))() for x, y in zip(b.gt_bboxes, b.gt_labels)]
metas = [dict(img_shape=i, pad_shape=p) for i, p in zip(host.img_shapes, host.pad_shapes)]
sc = [t.clone().requires_grad_() for t in b.s_cls]; sb = [t.clone().requires_grad_() for t in b.s_box]
def step():
    for t in sc + sb: t.grad = None
    sel = det.sel_pos(b.t_cls, b.t_box)
    out = head.loss_by_feat((b.t_cls, b.t_box), (sc, sb), sel[0], sel[1], sel[2], sel[3], 40, 1.0, None, gts, metas)
    parse_losses(out).backward()
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): step()
host_ms = (time.perf_counter() - t0) / 50 * 1e3
torch.cuda.synchronize()
wall_ms = (time.perf_counter() - t0) / 50 * 1e3
print(f'host launch time {host_ms:.3f} ms/step, wall incl. GPU drain {wall_ms:.3f} ms/step')
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
