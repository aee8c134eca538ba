"""The N > 1 host logic on CPU: two gloo ranks shard sequences without overlap, reduce timings as a
max and work as a sum, and all-reduce a flat gradient buffer (the training path's only collective)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from trackmpnn_b200 import parallel


def test_partition_is_a_balanced_cover():
    rs = np.random.RandomState(0)
    costs = rs.uniform(1, 100, size=37)
    for w in (1, 2, 4, 8):
        bins = parallel.partition_sequences(costs, w)
        assert sorted(i for b in bins for i in b) == list(range(37))
        load = [sum(costs[i] for i in b) for b in bins]
        assert max(load) - min(load) <= costs.max() + 1e-9
    assert parallel.partition_sequences(costs, 2) == parallel.partition_sequences(costs, 2)


def test_sequence_cost_grows_with_density():
    assert parallel.sequence_cost([10] * 20) < parallel.sequence_cost([20] * 20)
    assert parallel.sequence_cost([]) == 0.0


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        lin = torch.nn.Linear(5, 3)
        frozen = torch.nn.Parameter(torch.zeros(2), requires_grad=False)
        x = torch.full((4, 5), float(rank + 1))
        lin(x).sum().backward()
        want_w = sum(torch.full((3, 5), 4.0 * (r + 1)) for r in range(world))
        lin.bias.grad = None if rank == 0 else lin.bias.grad  # a missing gradient counts as zero
        n = parallel.allreduce_gradients(list(lin.parameters()) + [frozen])
        assert n == 18
        assert torch.allclose(lin.weight.grad, want_w)
        assert torch.allclose(lin.bias.grad, torch.full((3,), 4.0 * (world - 1)))
        # the flat gradient buffer of the training path: p.grad are views, the all-reduce exchanges the buffer itself
        from trackmpnn_b200.models.track_mpnn import TrackMPNN
        torch.manual_seed(5)
        model = TrackMPNN('2d', 3, 64, 0, 'diff')
        flat = parallel.FlatGradients(model)
        assert flat.flat.numel() == sum(p.numel() for p in model.parameters()) == 54914
        for k, p in enumerate(model.parameters()):
            assert flat.view_of(p) is not None
            p.grad.add_(float(rank + 1) * (k + 1))        # in place: what the backward kernels do
        assert flat.allreduce(average=True) == 54914
        for k, p in enumerate(model.parameters()):
            assert p.grad.data_ptr() == flat.views[id(p)].data_ptr()
            assert torch.allclose(p.grad, torch.full_like(p, (k + 1) * sum(r + 1 for r in range(world)) / world))
        flat.zero()
        assert float(flat.flat.abs().max()) == 0.0
        model.zero_grad()                                 # set_to_none: the views are gone -> the Function falls back to autograd
        assert all(flat.view_of(p) is None for p in model.parameters())
        costs = [parallel.sequence_cost([5 + (i % 3)] * 10) for i in range(9)]
        mine = parallel.partition_sequences(costs, world)[rank]
        tot = parallel.reduce_sum(len(mine))
        tmax = parallel.timed_max(10.0 * (rank + 1))
        if rank == 0:
            out.put((tot, tmax))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    tot, tmax = out.get()
    assert tot == 9 and tmax == 20.0
