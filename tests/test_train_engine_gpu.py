"""GPU: the batched trainer (B chunks as one block-diagonal graph per step) against the drop-in path run chunk by
chunk (itself pinned to the reference's golden gradients in test_cuda_golden.py): the batch loss is the sum of the
chunk losses and every parameter gradient the sum of the chunk gradients."""
import copy

import numpy as np
import pytest
import torch

from trackmpnn_b200 import synth

pytestmark = pytest.mark.gpu


def _chunk_loss_dropin(model, X, y, tp=True):
    """train.py:65-127 for one chunk through the drop-in modules (as bench.py's training leg does)."""
    from trackmpnn_b200.utils.graph import initialize_graph, update_graph
    from trackmpnn_b200.models.loss import create_targets, CELoss, FocalLoss
    ce, fn, fe = CELoss(), FocalLoss(gamma=0), FocalLoss(gamma=0)

    def losses(scores, logits, y_pred, labels, node_adj):
        idx_edge = torch.nonzero((y_pred[:, 0] == -1))[:, 0]
        idx_node = torch.nonzero((y_pred[:, 0] != -1))[:, 0]
        targets = create_targets(labels, node_adj, idx_node)
        l = ce(logits, targets, node_adj, idx_node) + fe(scores[idx_edge, 0], targets[idx_edge])
        if tp:
            l = l + fn(scores[idx_node, 0], targets[idx_node])
        return torch.cat((1 - scores, scores), dim=1), l

    y_pred, feats, node_adj, edge_adj, labels, t_st, t_end = initialize_graph(X, y, t_st=0, mode='train', cuda=True)
    scores, logits, states, _ = model(feats, None, node_adj, edge_adj)
    scores, loss = losses(scores, logits, y_pred, labels, node_adj)
    for t in range(t_st, t_end):
        y_pred, feats, node_adj, edge_adj, labels = update_graph(node_adj, labels, scores, y_pred, X, y, t,
                                                                 use_hungraian=False, mode='train', cuda=True)
        scores, logits, states, _ = model(feats, states, node_adj, edge_adj)
        scores, l = losses(scores, logits, y_pred, labels, node_adj)
        loss = loss + l
    return loss


@pytest.mark.parametrize('builder,msg_type,heads', [('slab', 'diff', 0), ('chunk', 'diff', 0), ('slab', 'concat', 0),
                                                     ('chunk', 'concat', 0), ('slab', 'diff', 2)])
def test_batch_equals_sum_of_chunks(msg_type, builder, heads):
    features = '2d'
    from trackmpnn_b200.models.track_mpnn import TrackMPNN
    from trackmpnn_b200.train_engine import TrainBatch, batch_loss
    dev = torch.device('cuda:0')
    torch.manual_seed(5)
    model = TrackMPNN(features, 3, 64, heads, msg_type).to(dev)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2 and '.gat.' not in name:
                p.mul_(5.0)
    for gru in model.factor_grus:
        # attention heads: dropout that keeps every entry (x 2), so that the batch and the chunk layouts agree
        gru.attention_keep_fn = lambda head, index, device: torch.ones(index.cap_inc, dtype=torch.uint8, device=device)
    model.train()
    chunks = []
    for i, dets in enumerate((6, 9, 4, 7)):
        ts = synth.train_chunk_timestamps(40 + i, 5, 2)
        Xn, yn = synth.make_sequence(40 + i, None, dets, 'kitti', timestamps=ts)
        chunks.append((torch.from_numpy(Xn).to(dev), torch.from_numpy(yn).to(dev)))
    ref = copy.deepcopy(model)
    tot = 0.0
    for X, y in chunks:
        l = _chunk_loss_dropin(ref, X, y)
        l.backward()                      # gradients accumulate over the chunks
        tot += float(l)
    batch = TrainBatch(chunks, dev, builder=builder)
    assert batch.builder == builder
    assert batch.num_chunks == len(chunks) and len(batch.steps) >= 7
    loss = batch_loss(model, batch)
    loss.backward()
    assert abs(float(loss) - tot) <= 1e-4 * max(1.0, abs(tot))
    gmax = max(float(p.grad.abs().max()) for p in ref.parameters())
    for (name, p), q in zip(model.named_parameters(), ref.parameters()):
        tol = 2e-3 * float(q.grad.abs().max()) + 1e-6 * gmax
        np.testing.assert_allclose(p.grad.cpu().numpy(), q.grad.cpu().numpy(), atol=tol, rtol=0, err_msg=name)
    # BatchNorm saw one batch per chunk and step in both runs
    for bm, br in zip(model.input_transforms, ref.input_transforms):
        assert int(bm[1].num_batches_tracked) == int(br[1].num_batches_tracked)


def test_training_steps_do_not_accumulate_device_memory():
    """The autograd step must not keep its graph alive after backward (a ctx attribute holding an output tensor closes
    a reference cycle that only the cycle collector frees: GBs per batched step)."""
    import gc
    from trackmpnn_b200.models.track_mpnn import TrackMPNN
    from trackmpnn_b200.train_engine import TrainBatch, batch_loss
    dev = torch.device('cuda:0')
    torch.manual_seed(5)
    model = TrackMPNN('2d', 3, 64, 0, 'diff').to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    chunks = []
    for i in range(4):
        ts = synth.train_chunk_timestamps(60 + i, 5, 2)
        Xn, yn = synth.make_sequence(60 + i, None, 12, 'kitti', timestamps=ts)
        chunks.append((torch.from_numpy(Xn).to(dev), torch.from_numpy(yn).to(dev)))
    batch = TrainBatch(chunks, dev)
    gc.collect()
    gc.disable()
    try:
        used = []
        for _ in range(4):
            opt.zero_grad()
            loss = batch_loss(model, batch)
            loss.backward()
            opt.step()
            del loss
            torch.cuda.synchronize()
            used.append(torch.cuda.memory_allocated())
        assert used[-1] <= used[1] + (1 << 20), used   # steady after the optimizer state exists
    finally:
        gc.enable()


def _kitti_chunks(dev, seeds, dets):
    chunks, host = [], []
    for sd in seeds:
        ts = synth.train_chunk_timestamps(sd, 5, 2)
        Xn, yn = synth.make_sequence(sd, None, dets, 'kitti', timestamps=ts)
        host.append((Xn, yn))
        chunks.append((torch.from_numpy(Xn).to(dev), torch.from_numpy(yn).to(dev)))
    return chunks, host


def test_flat_gradient_buffer_equals_autograd():
    """parallel.FlatGradients: the backward kernels accumulate straight into views of ONE flat buffer (what the data-parallel
    all-reduce exchanges, no pack / unpack); the result equals the gradients returned through autograd, over two
    accumulated batches, and zero() resets them."""
    from trackmpnn_b200 import parallel
    from trackmpnn_b200.models.track_mpnn import TrackMPNN
    from trackmpnn_b200.train_engine import TrainBatch, batch_loss
    dev = torch.device('cuda:0')
    torch.manual_seed(5)
    model = TrackMPNN('2d', 3, 64, 0, 'diff').to(dev).train()
    ref = copy.deepcopy(model)
    chunks, _ = _kitti_chunks(dev, [70, 71, 72, 73], 9)
    batch = TrainBatch(chunks, dev)
    for _ in range(2):
        batch_loss(ref, batch).backward()
    flat = parallel.FlatGradients(model)
    assert flat.flat.numel() == sum(p.numel() for p in model.parameters())
    for _ in range(2):
        batch_loss(model, batch).backward()
    gmax = max(float(p.grad.abs().max()) for p in ref.parameters())
    for (name, p), q in zip(model.named_parameters(), ref.parameters()):
        assert p.grad.data_ptr() == flat.views[id(p)].data_ptr(), name   # still the view: nothing was re-allocated
        tol = 1e-4 * float(q.grad.abs().max()) + 1e-6 * gmax               # float atomics: sums differ in their last bits
        np.testing.assert_allclose(p.grad.cpu().numpy(), q.grad.cpu().numpy(), atol=tol, rtol=0, err_msg=name)
    assert float(flat.flat.abs().sum()) > 0
    flat.zero()
    assert all(float(p.grad.abs().max()) == 0.0 for p in model.parameters())


def test_batched_trainer_against_training_oracle_at_bench_size():
    """The bench's training workload (BASELINE configs[1]: KITTI-shaped chunks, ~40 detections / frame), 8 chunks in one batch:
    loss and every parameter gradient against oracle/train_ref.py (torch fp32 ops on edge lists + autograd, itself pinned to
    the live reference's backward pass) run chunk by chunk and summed."""
    from oracle import train_ref as T
    from trackmpnn_b200.models.track_mpnn import TrackMPNN
    from trackmpnn_b200.train_engine import TrainBatch, batch_loss
    dev = torch.device('cuda:0')
    torch.manual_seed(5)
    model = TrackMPNN('2d', 3, 64, 0, 'diff').to(dev).train()
    params = {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}
    chunks, host = _kitti_chunks(dev, list(range(1000, 1008)), 40)
    loss_ref, grads_ref = 0.0, None
    for Xn, yn in host:
        r = T.train_chunk(params, Xn, yn)
        loss_ref += r['loss']
        grads_ref = r['grads'] if grads_ref is None else {k: grads_ref[k] + v for k, v in r['grads'].items()}
    batch = TrainBatch(chunks, dev)
    assert batch.edge_rows > 300000
    loss = batch_loss(model, batch)
    loss.backward()
    assert abs(float(loss) - loss_ref) <= 1e-4 * abs(loss_ref)
    # the golden bar (tests/golden_util.assert_grads_close): per parameter 2e-3 of its own largest entry + 1e-6 of the model's
    gmax = max(float(np.abs(v).max()) for v in grads_ref.values())
    for name, p in model.named_parameters():
        want = grads_ref[name]
        atol = 2e-3 * float(np.abs(want).max()) + 1e-6 * gmax + 1e-9
        np.testing.assert_allclose(p.grad.detach().cpu().numpy().reshape(want.shape), want, atol=atol, rtol=0, err_msg=name)


def test_graphed_train_step_equals_eager_steps():
    """GraphedTrainStep: [zero + forward + losses + backward] and [Adam] replayed as CUDA graphs (the weight images are
    re-packed inside the captured work).  Five optimizer steps -- three eager ones during capture, two replays -- leave the
    same parameters and loss as five eager steps of a second model (float atomics in a few small gradient sums: 1e-5)."""
    from trackmpnn_b200.models.track_mpnn import TrackMPNN
    from trackmpnn_b200.train_engine import GraphedTrainStep, TrainBatch
    dev = torch.device('cuda:0')
    torch.manual_seed(5)
    model = TrackMPNN('2d', 3, 64, 0, 'diff').to(dev).train()
    ref = copy.deepcopy(model)
    chunks, _ = _kitti_chunks(dev, [80, 81, 82, 83, 84, 85], 30)     # > 8192 rows per step: the tcgen05 kernels run
    batch = TrainBatch(chunks, dev)
    a = GraphedTrainStep(model, batch, lr=1e-3, weight_decay=5e-4)
    b = GraphedTrainStep(ref, batch, lr=1e-3, weight_decay=5e-4)
    a.capture(warmup=3)
    a.replay(); a.replay()
    for _ in range(5):
        b.eager()
    torch.cuda.synchronize()
    assert abs(float(a.loss) - float(b.loss)) <= 1e-5 * abs(float(b.loss))
    moved = 0.0
    for (name, p), q in zip(model.named_parameters(), ref.parameters()):
        if name.endswith('.0.bias'):
            # Linear bias in front of the train-mode BatchNorm: its gradient is analytically zero, what the kernels (and the
            # reference) produce is round-off, and Adam turns any round-off into steps of +-lr
            continue
        # 5 steps of lr 1e-3 move a weight by up to 5e-3; float atomics in a few gradient sums + Adam's normalisation: 2e-5
        np.testing.assert_allclose(p.detach().cpu().numpy(), q.detach().cpu().numpy(), rtol=0, atol=2e-5, err_msg=name)
    torch.manual_seed(5)
    fresh = TrackMPNN('2d', 3, 64, 0, 'diff').to(dev)
    for p, q in zip(model.parameters(), fresh.parameters()):
        moved = max(moved, float((p - q).abs().max()))
    assert moved > 1e-3     # the replays really trained
    for bm, br in zip(model.input_transforms, ref.input_transforms):   # BatchNorm bookkeeping advanced inside the graph too
        # the running mean follows Linear1's bias, which only moves by Adam-normalised round-off (see above): +-lr per step
        np.testing.assert_allclose(bm[1].running_mean.cpu().numpy(), br[1].running_mean.cpu().numpy(), rtol=0, atol=5e-3)
        np.testing.assert_allclose(bm[1].running_var.cpu().numpy(), br[1].running_var.cpu().numpy(), rtol=1e-3, atol=1e-6)
        assert int(bm[1].num_batches_tracked) == int(br[1].num_batches_tracked)


def test_batched_training_step_is_bit_reproducible():
    """Two optimizer steps of the batched trainer on the tensor-core path (40 detections / frame: > 8192 rows per step), run
    twice from the same weights: every parameter equal BIT FOR BIT.  The weight gradients of the GRU cells come from the
    tcgen05 contraction's per-SM partials, those of the input transform from per-chunk partials, the bias / head gradients
    from tmpnn_gate_bwd's per-CTA partials -- all added in a fixed order, no float atomics on this path."""
    from trackmpnn_b200.models.track_mpnn import TrackMPNN
    from trackmpnn_b200.train_engine import TrainBatch, GraphedTrainStep
    dev = torch.device('cuda:0')
    chunks, _ = _kitti_chunks(dev, [80, 81, 82, 83, 84, 85], 40)
    batch = TrainBatch(chunks, dev)
    finals = []
    for run in range(2):
        torch.manual_seed(5)
        model = TrackMPNN('2d', 3, 64, 0, 'diff').to(dev).train()
        step = GraphedTrainStep(model, batch, lr=1e-3, weight_decay=5e-4)
        for _ in range(2):
            step.eager()
        torch.cuda.synchronize()
        finals.append([p.detach().cpu().numpy().copy() for p in model.parameters()])
    changed = 0
    for a, b in zip(*finals):
        np.testing.assert_array_equal(a, b)
        changed += int(np.abs(a).sum() > 0)
    assert changed > 10
