"""Multi-GPU plumbing for the hot path (one process per GPU, ``torch.distributed``).

The path shards only across independent sequences / training chunks:

* inference: ``partition_sequences`` bin-packs the sequences onto the ranks by cost; every rank runs its
  own ``TrackEngine`` -- no data-path collective;
* training: ``allreduce_gradients`` sums one flat fp32 buffer of all parameter gradients across the ranks
  (NCCL over NVLink on GPUs, gloo in the CPU tests) -- the only exchange step the path has
  (data-parallel training, BASELINE.json configs[4]); BatchNorm statistics stay per chunk as in the
  reference (no SyncBN);
* ``timed_max`` / ``reduce_sum`` are the bench's "max over ranks" and "sum over ranks" reductions.
"""
import numpy as np
import torch
import torch.distributed as dist


def sequence_cost(frame_counts, window=5):
    """Work estimate of one sequence: edge rows summed over its frames ~ sum_t D_t * (detections of the
    previous window-1 frames), the quantity the MP step is linear in."""
    c = np.asarray(frame_counts, dtype=np.float64)
    cost = 0.0
    for t in range(len(c)):
        cost += c[t] * c[max(0, t - window + 1):t].sum() + c[t]
    return float(cost)


def partition_sequences(costs, world_size):
    """Greedy longest-processing-time bin packing.  Returns ``world_size`` lists of sequence indices;
    deterministic (ties broken by index), every index appears exactly once."""
    costs = [float(c) for c in costs]
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    bins = [[] for _ in range(world_size)]
    load = [0.0] * world_size
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        bins[r].append(i)
        load[r] += costs[i]
    return [sorted(b) for b in bins]


class FlatGradients:
    """ONE flat fp32 buffer holding every parameter gradient of a model; ``p.grad`` of every parameter is a view into it.

    The backward kernels of the message-passing step (``functional._MPStepFn.backward``) accumulate straight into these
    views, so a data-parallel step is: ``flat.zero()`` -> forward / backward -> ``flat.allreduce()`` (a single
    ``all_reduce`` of the buffer itself: no pack, no unpack, no per-parameter copies) -> ``optimizer.step()``.
    Use ``flat.zero()`` instead of ``optimizer.zero_grad()`` (whose default drops the views)."""

    def __init__(self, model):
        from .functional import _param_list
        step_params = _param_list(model)            # the order the autograd Function receives them in
        seen = {id(p) for p in step_params}
        self.params = step_params + [p for p in model.parameters() if id(p) not in seen]
        self.params = [p for p in self.params if p.requires_grad]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = {}
        off = 0
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            p.grad = v
            self.views[id(p)] = v
            off += p.numel()
        model.__dict__['_tmpnn_flat_grads'] = self

    def view_of(self, p):
        """The gradient view of parameter ``p`` while ``p.grad`` still is that view (None otherwise: the caller falls back
        to returning the gradient through autograd)."""
        v = self.views.get(id(p))
        if v is None or p.grad is None or p.grad.data_ptr() != v.data_ptr():
            return None
        return v

    def zero(self):
        self.flat.zero_()

    def allreduce(self, group=None, average=False):
        """Sums the buffer over the ranks in place; returns the number of floats exchanged."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                self.flat.mul_(1.0 / dist.get_world_size(group))
        return int(self.flat.numel())


def allreduce_gradients(params, group=None, average=False):
    """One all-reduce(sum) of a flat buffer holding every parameter gradient (missing grads count as 0),
    then scatter back into ``p.grad``.  Returns the number of floats exchanged."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    dev, dt = params[0].device, params[0].dtype
    n = sum(p.numel() for p in params)
    flat = torch.zeros(n, dtype=dt, device=dev)
    off = 0
    for p in params:
        if p.grad is not None:
            flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
        off += p.numel()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        g = flat[off:off + p.numel()].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += p.numel()
    return n


def _reduce(x, op, device):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=op)
    return float(t.item())


def timed_max(ms, device='cpu'):
    """A timing is the max over ranks (never the wall clock of one rank)."""
    return _reduce(ms, dist.ReduceOp.MAX, device)


def reduce_sum(x, device='cpu'):
    return _reduce(x, dist.ReduceOp.SUM, device)
