"""Multi-GPU check of erd_avg_exchange (run under torchrun, one rank per GPU): the peer-memory
reduce_mean must equal the NCCL all-reduce, give identical bits on every rank, and survive
CUDA-graph replay with changing inputs.  tests/test_gpu_multi.py launches it when >= 2 GPUs exist."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erd_b200 import _native as N                      # noqa: E402
from erd_b200.dist_utils import peer_exchange, reduce_mean_   # noqa: E402


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    lib = N.load()
    ex = peer_exchange(lib, dev)
    assert ex is not None, 'peer exchange could not be set up'
    g = torch.Generator(device='cpu').manual_seed(1234 + rank)
    worst = 0.0
    for step in range(20):
        mine = (torch.rand(2, generator=g) * 1000).to(dev)
        want = reduce_mean_(mine.clone())
        got = ex.reduce_mean_(mine.clone())
        torch.cuda.synchronize()
        worst = max(worst, float(((got - want).abs() / want.abs()).max()))
        every = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(every, got)
        assert all(torch.equal(every[0], e) for e in every), 'ranks disagree'
    assert worst <= 1e-6, worst
    # graph replay: the epoch lives in device memory, inputs change between replays
    buf = torch.zeros(2, device=dev)
    src = torch.zeros(2, device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        buf.copy_(src)
        ex.reduce_mean_(buf)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        buf.copy_(src)
        ex.reduce_mean_(buf)
    for step in range(50):
        src.copy_((torch.rand(2, generator=g) * 10).to(dev))
        want = reduce_mean_(src.clone())
        graph.replay()
        torch.cuda.synchronize()
        assert float(((buf - want).abs() / want.abs()).max()) <= 1e-6
    assert not ex.timed_out()
    dist.barrier()
    if rank == 0:
        print(f'peer exchange ok: world={world}, worst rel err vs NCCL {worst:.2e}', flush=True)
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == '__main__':
    main()
