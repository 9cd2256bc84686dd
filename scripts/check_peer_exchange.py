"""Multi-GPU check of the avg-factor exchange (run under torchrun, one rank per GPU).
(1) erd_avg_exchange, the stand-alone kernel: equals the NCCL all-reduce, identical bits on every rank, survives
    CUDA-graph replay with changing inputs.
(2) The exchange fused into the step (the assignment prepass posts, the student pass waits): every rank runs the whole
    path on its OWN batch; the factors it ends up with must be the mean of the ranks' local factors bit for bit, and its
    losses and gradients must match the CPU oracle fed those world-averaged factors through its reduce_mean hook
    (gfl_head_increment_erd.py:390-391,406-409, mmdet/utils/dist_utils.py:59-65).
tests/test_gpu_multi.py launches it when >= 2 GPUs exist."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erd_b200 import _native as N                      # noqa: E402
from erd_b200.dist_utils import peer_exchange, reduce_mean_   # noqa: E402


def fused_step_against_oracle(rank, world, dev):
    from erd_b200.ops import ErdPath
    from erd_b200.synth import make_batch
    from oracle import erd_oracle as O
    batch = make_batch(2, (320, 480), ori=40, seed=500 + rank, mode='trained' if rank % 2 else 'gaussian',
                       num_gt=[1 + rank, 5], gt_size_pow=2.0)
    b = batch.to(dev)
    path = ErdPath()
    for rep in range(3):   # more than one step: epochs and slot parity advance
        p = path.plan(b.s_cls, b.num_classes, b.ori, b.reg_max)
        p.set_targets(b.gt_bboxes, b.gt_labels, b.pad_shapes)
        g_cls = [torch.empty_like(t) for t in b.s_cls]
        g_box = [torch.empty_like(t) for t in b.s_box]
        losses = torch.empty(p.num_losses, device=dev)
        path.prepare(p, b.t_cls, b.t_box, b.s_cls, b.s_box)
        local = p.avg.clone()          # stream-ordered: this rank's factors before the student pass averages them
        path.reduce_avg(p)             # nothing left to do when the exchange is fused
        path.loss_fwd_bwd(p, b.t_cls, b.t_box, b.s_cls, b.s_box, g_cls, g_box, losses, 1.0)
        torch.cuda.synchronize()
        every = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(every, local)
        want = torch.zeros_like(local)
        for t in every:
            want += t / world
        assert torch.equal(p.avg, want), (rank, p.avg.tolist(), want.tolist())
    means = [float(x) for x in want.tolist()]
    calls = []

    def reduce_mean(t):
        calls.append(float(t))
        return torch.tensor(means[len(calls) - 1], dtype=torch.float)
    s_cls = [t.clone().requires_grad_() for t in batch.s_cls]
    s_box = [t.clone().requires_grad_() for t in batch.s_box]
    ref, _, _ = O.erd_step(batch.t_cls, batch.t_box, s_cls, s_box, batch.gt_bboxes, batch.gt_labels, batch.pad_shapes,
                           batch.ori, 1.0, batch.num_classes, batch.reg_max, reduce_mean=reduce_mean)
    assert abs(calls[0] - float(every[rank][0])) < 1e-6 and abs(calls[1] - float(every[rank][1])) <= 1e-5 * max(1.0, calls[1])
    flat = ref['loss_cls'] + ref['loss_bbox'] + ref['loss_dfl'] + ref['loss_dist_cls'] + ref['loss_dist_bbox']
    got = losses.cpu()
    for j, x in enumerate(flat):
        assert abs(float(got[j]) - float(x)) <= 1e-5 * max(abs(float(x)), 1e-7), (rank, j, float(got[j]), float(x))
    for a, r in zip(g_cls + g_box, [t.grad for t in s_cls + s_box]):
        scale = float(r.abs().max())
        assert float((a.cpu() - r).abs().max()) <= 1e-5 * max(scale, 1e-12), rank


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
    fused_step_against_oracle(rank, world, dev)
    assert not ex.timed_out()
    dist.barrier()
    if rank == 0:
        print(f'peer exchange ok: world={world}, worst rel err vs NCCL {worst:.2e}; fused step matches the oracle '
              f'with world-averaged factors on every rank', flush=True)
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == '__main__':
    main()
