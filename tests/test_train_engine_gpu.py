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
