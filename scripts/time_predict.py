"""Dev: time the inference post-process (erd_predict) next to the CPU oracle on the same inputs."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erd_b200.predict import ErdPredictor
from erd_b200.synth import make_batch
from oracle import predict_oracle as P
n = int(os.environ.get('IMGS', 16)); shift = float(os.environ.get('SHIFT', 3.0))
b = make_batch(n, (800, 1333), ori=40, seed=5, mode='trained')
s_cls = [t + shift for t in b.s_cls]
dc, db = [t.cuda() for t in s_cls], [t.cuda() for t in b.s_box]
pred = ErdPredictor()
for _ in range(3): out = pred.predict_by_feat(dc, db, b.img_shapes)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): out = pred.predict_by_feat(dc, db, b.img_shapes)
e1.record(); torch.cuda.synchronize()
gpu_ms = e0.elapsed_time(e1) / 20
t0 = time.perf_counter(); ref = P.predict_by_feat(s_cls, b.s_box, b.img_shapes); cpu_ms = (time.perf_counter() - t0) * 1e3
cand = sum(int((t.sigmoid() > 0.05).sum()) for t in s_cls)
print(f'{n} images 800x1333, {cand} scores above 0.05, detections/img {[int(o["labels"].numel()) for o in out][:4]}: '
      f'erd_predict {gpu_ms:.3f} ms per batch (incl. the count read-back), CPU oracle {cpu_ms:.0f} ms ({torch.get_num_threads()} threads)')
