"""GPU parity: the drop-in API (TrackMPNN.forward + utils.graph functions, which call the
C ABI of libtmpnn_sm100a.so) replayed against the reference's golden vectors.

Integer outputs (y_pred, adjacency, labels, y_out) bit-exact; logits / hidden states
max-abs <= 1e-4 (BASELINE.json north_star tolerance)."""
import numpy as np
import pytest
import torch

from golden_util import Golden, golden_names, pin_attention_dropout

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _model(gold, dev, train=False, tensor=False):
    from trackmpnn_b200.models.track_mpnn import TrackMPNN
    m = gold.meta
    model = TrackMPNN(m['features'], m['ncategories'], 64, m.get('nattheads', 0), m['msg_type'], use_tensor_cores=bool(tensor))
    if tensor:
        model.tensor_core_kernel = tensor   # 'gather' | 'pre'
    sd = {k: torch.from_numpy(v) for k, v in gold.params().items()}
    model.load_state_dict(sd, strict=True)
    model.to(dev)
    model.train(train)
    return model


def _dense(adj):
    return adj.to_dense().cpu().numpy()


def _ref_dense(gold, s, prefix=''):
    r, c, v = gold.get(s, prefix + 'adj').astype(np.int64)
    n = gold.get(s, prefix + 'y_pred').shape[0]
    a = np.zeros((n, n), np.float32)
    a[r, c] = v
    return a


def _fix(scores, y_pred, tp):
    scores = torch.cat((1 - scores, scores), dim=1)
    if not tp:
        idn = torch.nonzero((y_pred[:, 0] != -1))[:, 0]
        scores[idn, 0] = 0
        scores[idn, 1] = 1
    return scores


INFER = golden_names('infer')


@pytest.mark.parametrize('tensor', [False, 'gather', 'pre'], ids=['fma', 'tcgen05', 'tcgen05-pre'])
@pytest.mark.parametrize('name', INFER)
def test_infer_free_running(name, tensor):
    """tensor=False: fp32 FMA kernel; 'gather' / 'pre': the tcgen05 kernels (3-term fp16 split; endpoints gathered
    per association row, or prepared once per detection row) -- same 1e-4 bar."""
    from trackmpnn_b200.utils.graph import initialize_graph, update_graph, prune_graph, decode_tracks
    gold = Golden(name)
    m = gold.meta
    if tensor == 'gather' and m['msg_type'] != 'diff':
        pytest.skip('the gather-and-split tensor-core kernel covers msg_type diff')
    dev = torch.device('cuda:0')
    model = _model(gold, dev, tensor=tensor)
    X, y = torch.from_numpy(gold.X).to(dev), torch.from_numpy(gold.y).to(dev)
    y_out = gold.y[0].astype(np.int64); y_out[:, 1] = -1
    with torch.no_grad():
        y_pred, feats, node_adj, edge_adj, labels, t_st, t_end = initialize_graph(X, y, 0, 'test', True)
        assert t_st == int(gold.z['t_st']) and t_end == int(gold.z['t_end'])
        np.testing.assert_array_equal(y_pred.cpu().numpy(), gold.get(0, 'y_pred'))
        np.testing.assert_array_equal(feats.cpu().numpy(), gold.get(0, 'feats'))
        np.testing.assert_array_equal(_dense(node_adj), _ref_dense(gold, 0))
        np.testing.assert_array_equal(_dense(edge_adj), _ref_dense(gold, 0).T * (1 - np.eye(y_pred.shape[0], dtype=np.float32))
                                      + np.diag((gold.get(0, 'y_pred')[:, 0] == -1).astype(np.float32)))
        np.testing.assert_array_equal(labels.cpu().numpy(), gold.get(0, 'labels'))
        scores, logits, states, att = model(feats, None, node_adj, edge_adj)
        assert len(att) == len(model.feature_idx) and all((a is None) == (m.get('nattheads', 0) == 0) for a in att)
        np.testing.assert_allclose(logits.cpu().numpy(), gold.get(0, 'logits'), atol=TOL, rtol=0)
        np.testing.assert_allclose(states.cpu().numpy(), gold.get(0, 'h'), atol=TOL, rtol=0)
        scores = _fix(scores, y_pred, m['tp_classifier'])
        s = 0
        t_skip = t_st
        for t_cur in range(t_st, t_end):
            if t_cur < t_skip:
                continue
            s += 1
            if feats.size()[0] == 0 and states.size()[0] == 0:
                y_pred, feats, node_adj, edge_adj, labels, t_skip, _ = initialize_graph(X, y, t_cur, 'test', True)
                if y_pred is None:
                    break
                states = None
            else:
                y_pred, feats, node_adj, edge_adj, labels = update_graph(
                    node_adj, labels, scores, y_pred, X, y, t_cur, use_hungraian=m['hungarian'], mode='test', cuda=True)
            np.testing.assert_array_equal(y_pred.cpu().numpy(), gold.get(s, 'y_pred'), err_msg=f'step {s}')
            np.testing.assert_array_equal(feats.cpu().numpy(), gold.get(s, 'feats'))
            np.testing.assert_array_equal(_dense(node_adj), _ref_dense(gold, s))
            np.testing.assert_array_equal(labels.cpu().numpy(), gold.get(s, 'labels'))
            scores, logits, states, att = model(feats, states, node_adj, edge_adj)
            np.testing.assert_allclose(logits.cpu().numpy(), gold.get(s, 'logits'), atol=TOL, rtol=0)
            np.testing.assert_allclose(states.cpu().numpy(), gold.get(s, 'h'), atol=TOL, rtol=0)
            if gold.has(s, 'att'):  # the reference's dense attention [groups, heads, N, N]
                want = gold.get(s, 'att')
                for gi in range(want.shape[0]):
                    for k in range(want.shape[1]):
                        np.testing.assert_allclose(att[gi][k].to_dense().cpu().numpy(), want[gi, k], atol=1e-5, rtol=0)
            scores = _fix(scores, y_pred, m['tp_classifier'])
            if gold.has(s, 'prune_y_pred'):
                t_lo, t_hi = gold.get(s, 'prune_t')
                y_pred, states, node_adj, labels, scores = prune_graph(
                    states, node_adj, labels, scores, y_pred, int(t_lo), int(t_hi), threshold=0.5, cuda=True)
                np.testing.assert_array_equal(y_pred.cpu().numpy(), gold.get(s, 'prune_y_pred'))
                np.testing.assert_array_equal(_dense(node_adj), _ref_dense(gold, s, 'prune_'))
                np.testing.assert_allclose(states.cpu().numpy(), gold.get(s, 'prune_h'), atol=TOL, rtol=0)
                np.testing.assert_array_equal(labels.cpu().numpy(), gold.get(s, 'prune_labels'))
            t_upto = t_end if t_cur == t_end - 1 else t_cur - m['cur_win_size'] + 2
            y_pred, y_out, states, node_adj, labels, scores = decode_tracks(
                states, node_adj, labels, scores, y_pred, y_out, t_upto, m['ret_win_size'], use_hungraian=m['hungarian'], cuda=True)
            np.testing.assert_array_equal(y_pred.cpu().numpy(), gold.get(s, 'dec_y_pred'), err_msg=f'decode step {s}')
            np.testing.assert_array_equal(y_out, gold.get(s, 'y_out'), err_msg=f'y_out step {s}')
            np.testing.assert_array_equal(_dense(node_adj), _ref_dense(gold, s, 'dec_'))
            np.testing.assert_allclose(states.cpu().numpy(), gold.get(s, 'dec_h'), atol=TOL, rtol=0)
            np.testing.assert_allclose(scores.cpu().numpy(), gold.get(s, 'dec_scores'), atol=TOL, rtol=0)
            np.testing.assert_array_equal(labels.cpu().numpy(), gold.get(s, 'dec_labels'))
        assert s + 1 == gold.n_steps


@pytest.mark.parametrize('name', golden_names('train'))
def test_train_forward(name):
    """Teacher-forced graph growth and the train-mode forward (BatchNorm batch statistics incl. the
    zero edge rows, running-stat updates)."""
    from trackmpnn_b200.utils.graph import initialize_graph, update_graph
    gold = Golden(name)
    m = gold.meta
    dev = torch.device('cuda:0')
    model = _model(gold, dev, train=True)
    state = {'s': 0, 'n': 0}
    pin_attention_dropout(model, gold, state)
    X, y = torch.from_numpy(gold.X).to(dev), torch.from_numpy(gold.y).to(dev)
    with torch.no_grad():
        y_pred, feats, node_adj, edge_adj, labels, t_st, t_end = initialize_graph(X, y, 0, 'train', True)
        states = None
        s = 0
        for t_cur in [None] + list(range(t_st, t_end)):
            if t_cur is not None:
                s += 1
                sc = torch.from_numpy(gold.get(s - 1, 'scores')).to(dev)
                y_pred, feats, node_adj, edge_adj, labels = update_graph(
                    node_adj, labels, sc, y_pred, X, y, t_cur, use_hungraian=False, mode='train', cuda=True)
            np.testing.assert_array_equal(y_pred.cpu().numpy(), gold.get(s, 'y_pred'), err_msg=f'step {s}')
            np.testing.assert_array_equal(labels.cpu().numpy(), gold.get(s, 'labels'))
            np.testing.assert_array_equal(_dense(node_adj), _ref_dense(gold, s))
            state.update(s=s, n=int(y_pred.shape[0]))
            scores, logits, states, _ = model(feats, states, node_adj, edge_adj)
            np.testing.assert_allclose(logits.cpu().numpy(), gold.get(s, 'logits'), atol=TOL, rtol=0)
            np.testing.assert_allclose(states.cpu().numpy(), gold.get(s, 'h'), atol=TOL, rtol=0)
    sd = model.state_dict()
    for k in gold.z.files:
        if k.startswith('w_after/'):
            np.testing.assert_allclose(sd[k[len('w_after/'):]].cpu().numpy(), gold.z[k], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize('name', golden_names('train'))
def test_train_backward(name):
    """One BPTT chunk exactly as train.py:65-134 drives it, through the drop-in modules
    (TrackMPNN.forward as an autograd Function, create_targets / CELoss / FocalLoss, loss.backward()):
    targets bit-exact, per-step losses, the total loss and every parameter gradient against the
    reference's own autograd."""
    from golden_util import assert_grads_close
    from trackmpnn_b200.utils.graph import initialize_graph, update_graph
    from trackmpnn_b200.models.loss import create_targets, CELoss, FocalLoss
    gold = Golden(name)
    m = gold.meta
    dev = torch.device('cuda:0')
    model = _model(gold, dev, train=True)
    state = {'s': 0, 'n': 0}
    pin_attention_dropout(model, gold, state)
    X, y = torch.from_numpy(gold.X).to(dev), torch.from_numpy(gold.y).to(dev)
    ce, focal_node, focal_edge = CELoss(), FocalLoss(gamma=0), FocalLoss(gamma=0)

    def losses(scores, logits, y_pred, labels, node_adj, s):
        idx_edge = torch.nonzero((y_pred[:, 0] == -1))[:, 0]
        idx_node = torch.nonzero((y_pred[:, 0] != -1))[:, 0]
        targets = create_targets(labels, node_adj, idx_node)
        np.testing.assert_array_equal(targets.cpu().numpy(), gold.get(s, 'targets'), err_msg=f'targets step {s}')
        loss_c = ce(logits, targets, node_adj, idx_node)
        if m['tp_classifier']:
            loss_f = focal_node(scores[idx_node, 0], targets[idx_node]) + focal_edge(scores[idx_edge, 0], targets[idx_edge])
            scores = torch.cat((1 - scores, scores), dim=1)
        else:
            loss_f = focal_edge(scores[idx_edge, 0], targets[idx_edge])
            scores = torch.cat((1 - scores, scores), dim=1)
            scores[idx_node, 0] = 0
            scores[idx_node, 1] = 1
        np.testing.assert_allclose(loss_c.item(), gold.get(s, 'loss_c'), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(loss_f.item(), gold.get(s, 'loss_f'), rtol=1e-4, atol=1e-5)
        return scores, loss_c, loss_f

    y_pred, feats, node_adj, edge_adj, labels, t_st, t_end = initialize_graph(X, y, t_st=0, mode='train', cuda=True)
    state.update(s=0, n=int(y_pred.shape[0]))
    scores, logits, states, _ = model(feats, None, node_adj, edge_adj)
    assert logits.requires_grad and states.requires_grad
    s = 0
    scores, loss_c, loss_f = losses(scores, logits, y_pred, labels, node_adj, s)
    for t_cur in range(t_st, t_end):
        s += 1
        y_pred, feats, node_adj, edge_adj, labels = update_graph(
            node_adj, labels, scores, y_pred, X, y, t_cur, use_hungraian=False, mode='train', cuda=True)
        state.update(s=s, n=int(y_pred.shape[0]))
        scores, logits, states, att = model(feats, states, node_adj, edge_adj)
        # attention slots: one SparseAttention per head (after dropout), None without heads
        assert all((a is None) == (m.get('nattheads', 0) == 0) for a in att)
        assert all(a is None or len(a) == m['nattheads'] for a in att)
        np.testing.assert_allclose(logits.detach().cpu().numpy(), gold.get(s, 'logits'), atol=TOL, rtol=0)
        scores, lc, lf = losses(scores, logits, y_pred, labels, node_adj, s)
        loss_c = loss_c + lc
        loss_f = loss_f + lf
    assert s + 1 == gold.n_steps
    loss = loss_c + loss_f
    np.testing.assert_allclose(loss.item(), float(gold.z['loss']), rtol=1e-4)
    loss.backward()
    got = {k: (torch.zeros_like(p) if p.grad is None else p.grad).cpu().numpy() for k, p in model.named_parameters()}
    assert_grads_close(got, gold)
    # an optimizer step on these gradients runs and changes the packed cells on the next forward
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=5e-4)
    opt.step()
