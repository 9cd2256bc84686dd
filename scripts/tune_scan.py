"""Tuning: time erd_ers_select (scan + select) for each ERD_SCAN_TILE in a fresh process."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from erd_b200.ops import ErdPath
from erd_b200.synth import make_batch
b = make_batch(16, (800, 1333), ori=40, seed=1).to('cuda')
path = ErdPath(); p = path.plan(b.s_cls, 80, 40, 16)
for _ in range(5): path.ers_select(p, b.t_cls, b.t_box)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): path.ers_select(p, b.t_cls, b.t_box)
e1.record(); torch.cuda.synchronize()
print(os.environ.get('ERD_SCAN_TILE'), 'scan+select us/iter', 1000 * e0.elapsed_time(e1) / 50)
''' % ROOT
for v in sys.argv[1:] or ['64', '128', '256']:
    env = dict(os.environ, ERD_SCAN_TILE=v)
    print(subprocess.run([sys.executable, '-c', CHILD], env=env, capture_output=True, text=True).stdout.strip(), flush=True)
