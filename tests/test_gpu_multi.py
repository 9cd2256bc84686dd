"""Multi-GPU parity of the path's one collective -- stand-alone and fused into the step, the latter against the oracle with
world-averaged factors (skipped on a single-GPU box; bench.py --gpus N asserts the factors there)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs')
def test_peer_memory_avg_exchange_matches_nccl():
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}',
           '--master-addr', '127.0.0.1', '--master-port', '29533',
           os.path.join(ROOT, 'scripts', 'check_peer_exchange.py')]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert 'peer exchange ok' in res.stdout
