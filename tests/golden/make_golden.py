"""Generates the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports ``models.track_mpnn``, ``models.loss`` and ``utils.graph`` from
``/root/reference`` (read-only), drives them with the loops of the reference's
``infer.py:48-87`` and ``train.py:65-134`` (4-value unpack of ``forward``) on the
seeded synthetic streams of ``trackmpnn_b200/synth.py`` and stores every tensor that
crosses the hot-path boundary.  The fixtures pin ``oracle/trackmpnn_oracle.py``
(``tests/test_oracle_golden.py``) and, through it and directly, the CUDA path.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

from models.track_mpnn import TrackMPNN  # noqa: E402  (reference)
from models.loss import create_targets, CELoss, FocalLoss  # noqa: E402  (reference)
from utils.graph import initialize_graph, update_graph, prune_graph, decode_tracks  # noqa: E402  (reference)
from trackmpnn_b200 import synth  # noqa: E402
sys.path.insert(0, os.path.dirname(HERE))
from golden_util import attention_keep_matrix  # noqa: E402  (the pinned dropout decisions of the attention heads)


def coo(adj):
    """Off-diagonal entries of an adjacency (dense or sparse) as int32 [3, nnz] rows (r, c, v)."""
    a = adj.to_dense() if adj.is_sparse else adj
    a = a.detach().cpu().numpy()
    r, c = np.nonzero(a)
    return np.stack((r, c, a[r, c])).astype(np.int32)


def make_model(features, ncat, msg_type, scale, edge_bias, seed=5, nattheads=0):
    torch.manual_seed(seed)
    m = TrackMPNN(features=features, ncategories=ncat, nhidden=64, nattheads=nattheads, msg_type=msg_type)
    if scale != 1.0:
        with torch.no_grad():
            for name, p in m.named_parameters():
                if p.dim() >= 2 and '.gat.' not in name:  # attention parameters keep their xavier init
                    p.mul_(scale)
    if edge_bias is not None:
        with torch.no_grad():
            m.output_transform_edge.bias.fill_(edge_bias)
    return m


def add_features(X, y, features):
    """Pads X with deterministic pseudo 'temp' (2) / 'vis' (128) columns when requested."""
    cols = [X]
    rs = np.random.RandomState(123)
    if 'temp' in features:
        t = y[0, :, 0:1]
        cols.append(np.concatenate((np.sin(t / 10.0), np.cos(t / 10.0)), 1)[None].astype(np.float32))
    if 'vis' in features:
        v = rs.normal(0, 1, (1, X.shape[1], 128)).astype(np.float32)
        cols.append(v)
    return np.concatenate(cols, 2)


def run_infer(name, seed, frames, dets, dataset, features, msg_type, scale, edge_bias, hungarian,
              ret_win, cur_win, tp_classifier, prune_at=None, nattheads=0):
    Xn, yn = synth.make_sequence(seed, frames, dets, dataset)
    Xn = add_features(Xn, yn, features)
    ncat = synth.num_categories(dataset)
    model = make_model(features, ncat, msg_type, scale, edge_bias, nattheads=nattheads)
    model.eval()
    X = torch.from_numpy(Xn); y = torch.from_numpy(yn)
    out = {'X': Xn, 'y': yn}
    meta = dict(kind='infer', features=features, ncategories=ncat, msg_type=msg_type, hungarian=hungarian,
                ret_win_size=ret_win, cur_win_size=cur_win, tp_classifier=tp_classifier, dataset=dataset,
                nattheads=nattheads)
    for k, v in model.state_dict().items():
        out['w/' + k] = v.detach().cpu().numpy().copy()

    y_out = yn[0].astype(np.int64)
    y_out[:, 1] = -1
    steps = []

    def fix(scores, y_pred):
        scores = torch.cat((1 - scores, scores), dim=1)
        if not tp_classifier:
            idn = torch.nonzero((y_pred[:, 0] != -1))[:, 0]
            scores[idn, 0] = 0
            scores[idn, 1] = 1
        return scores

    with torch.no_grad():
        y_pred, feats, node_adj, edge_adj, labels, t_st, t_end = initialize_graph(X, y, t_st=0, mode='test', cuda=False)
        scores, logits, states, _ = model(feats, None, node_adj, edge_adj)
        scores = fix(scores, y_pred)
        s = 0
        out[f's{s}/op'] = np.array(0)  # 0 = init
        out[f's{s}/t'] = np.array(0)
        out[f's{s}/y_pred'] = y_pred.numpy(); out[f's{s}/feats'] = feats.numpy()
        out[f's{s}/adj'] = coo(node_adj); out[f's{s}/labels'] = labels.numpy()
        out[f's{s}/scores'] = scores.numpy(); out[f's{s}/logits'] = logits.numpy(); out[f's{s}/h'] = states.numpy()
        out['t_st'] = np.array(t_st); out['t_end'] = np.array(t_end)
        t_skip = t_st
        for t_cur in range(t_st, t_end):
            if t_cur < t_skip:
                continue
            s += 1
            if feats.size()[0] == 0 and states.size()[0] == 0:
                r = initialize_graph(X, y, t_st=t_cur, mode='test', cuda=False)
                y_pred, feats, node_adj, edge_adj, labels, t_skip, _ = r
                if y_pred is None:
                    s -= 1
                    break
                states = None
                out[f's{s}/op'] = np.array(2)  # 2 = re-init
                out[f's{s}/t_skip'] = np.array(t_skip)
            else:
                y_pred, feats, node_adj, edge_adj, labels = update_graph(
                    node_adj, labels, scores, y_pred, X, y, t_cur, use_hungraian=hungarian, mode='test', cuda=False)
                out[f's{s}/op'] = np.array(1)  # 1 = update
            out[f's{s}/t'] = np.array(t_cur)
            out[f's{s}/y_pred'] = y_pred.numpy(); out[f's{s}/feats'] = feats.numpy()
            out[f's{s}/adj'] = coo(node_adj); out[f's{s}/labels'] = labels.numpy()
            scores, logits, states, att = model(feats, states, node_adj, edge_adj)
            scores = fix(scores, y_pred)
            out[f's{s}/scores'] = scores.numpy(); out[f's{s}/logits'] = logits.numpy(); out[f's{s}/h'] = states.numpy()
            if nattheads > 0 and s in (1, 3):  # dense attention [groups, heads, N, N] of two small steps
                out[f's{s}/att'] = np.stack([np.stack([a.numpy() for a in grp]) for grp in att])
            if prune_at is not None and t_cur == prune_at:
                t_lo, t_hi = t_cur - 2, t_cur - 1
                y_pred, states, node_adj, labels, scores = prune_graph(
                    states, node_adj, labels, scores, y_pred, t_lo, t_hi, threshold=0.5, cuda=False)
                out[f's{s}/prune_t'] = np.array([t_lo, t_hi])
                out[f's{s}/prune_y_pred'] = y_pred.numpy(); out[f's{s}/prune_adj'] = coo(node_adj)
                out[f's{s}/prune_h'] = states.numpy(); out[f's{s}/prune_scores'] = scores.numpy()
                out[f's{s}/prune_labels'] = labels.numpy()
            t_upto = t_end if t_cur == t_end - 1 else t_cur - cur_win + 2
            y_pred, y_out, states, node_adj, labels, scores = decode_tracks(
                states, node_adj, labels, scores, y_pred, y_out, t_upto, ret_win, use_hungraian=hungarian, cuda=False)
            out[f's{s}/t_upto'] = np.array(t_upto)
            out[f's{s}/dec_y_pred'] = y_pred.numpy(); out[f's{s}/dec_adj'] = coo(node_adj)
            out[f's{s}/dec_h'] = states.numpy(); out[f's{s}/dec_scores'] = scores.numpy()
            out[f's{s}/dec_labels'] = labels.numpy(); out[f's{s}/y_out'] = y_out.copy()
    out['n_steps'] = np.array(s + 1)
    out['meta'] = np.array(repr(meta))
    ntr = len(set(y_out[:, 1].tolist()) - {-1})
    print(f'{name}: {Xn.shape[1]} dets, {s + 1} steps, {ntr} tracks, max rows '
          f'{max(out[f"s{k}/y_pred"].shape[0] for k in range(s + 1))}')
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)


class _PinnedDropout(torch.nn.Module):
    """Stands in for ``GraphAttentionLayer.dropout`` (reference models/layers.py:24) with the pinned mask."""

    def __init__(self, seed, group, head, step_ref):
        super().__init__()
        self.seed, self.group, self.head, self.step_ref = seed, group, head, step_ref

    def forward(self, attention):
        keep = attention_keep_matrix(self.seed, self.step_ref[0], self.group, self.head, attention.shape[0])
        return attention * torch.from_numpy(keep.astype(np.float32) * 2.0)


def run_train(name, seed, dets, dataset, features, msg_type, scale, edge_bias, tp_classifier, ret_win=0, cur_win=5,
              nattheads=0):
    ts = synth.train_chunk_timestamps(seed, cur_win, ret_win + 2)
    Xn, yn = synth.make_sequence(seed, None, dets, dataset, timestamps=ts)
    Xn = add_features(Xn, yn, features)
    ncat = synth.num_categories(dataset)
    model = make_model(features, ncat, msg_type, scale, edge_bias, nattheads=nattheads)
    model.train()
    step_ref = [0]
    for g, gru in enumerate(model.factor_grus):
        for k, head in enumerate(gru.gat or ()):
            head.dropout = _PinnedDropout(seed, g, k, step_ref)
    X = torch.from_numpy(Xn); y = torch.from_numpy(yn)
    out = {'X': Xn, 'y': yn}
    meta = dict(kind='train', features=features, ncategories=ncat, msg_type=msg_type, tp_classifier=tp_classifier,
                ret_win_size=ret_win, cur_win_size=cur_win, dataset=dataset, timestamps=ts, nattheads=nattheads, seed=seed)
    for k, v in model.state_dict().items():
        out['w/' + k] = v.detach().cpu().numpy().copy()
    focal_node, focal_edge, ce = FocalLoss(gamma=0), FocalLoss(gamma=0), CELoss()

    def losses(scores, logits, y_pred, labels, node_adj, s):
        idx_edge = torch.nonzero((y_pred[:, 0] == -1))[:, 0]
        idx_node = torch.nonzero((y_pred[:, 0] != -1))[:, 0]
        targets = create_targets(labels, node_adj, idx_node)
        loss_c = ce(logits, targets, node_adj, idx_node)
        if tp_classifier:
            loss_f = focal_node(scores[idx_node, 0], targets[idx_node]) + focal_edge(scores[idx_edge, 0], targets[idx_edge])
            scores = torch.cat((1 - scores, scores), dim=1)
        else:
            loss_f = focal_edge(scores[idx_edge, 0], targets[idx_edge])
            scores = torch.cat((1 - scores, scores), dim=1)
            scores[idx_node, 0] = 0
            scores[idx_node, 1] = 1
        out[f's{s}/targets'] = targets.numpy()
        out[f's{s}/loss_c'] = loss_c.detach().numpy(); out[f's{s}/loss_f'] = loss_f.detach().numpy()
        return scores, loss_c, loss_f

    y_pred, feats, node_adj, edge_adj, labels, t_st, t_end = initialize_graph(X, y, t_st=0, mode='train', cuda=False)
    scores, logits, states, _ = model(feats, None, node_adj, edge_adj)
    s = 0
    out[f's{s}/op'] = np.array(0); out[f's{s}/t'] = np.array(0)
    out[f's{s}/y_pred'] = y_pred.numpy(); out[f's{s}/feats'] = feats.numpy()
    out[f's{s}/adj'] = coo(node_adj); out[f's{s}/labels'] = labels.numpy()
    out[f's{s}/logits'] = logits.detach().numpy(); out[f's{s}/h'] = states.detach().numpy()
    scores, loss_c, loss_f = losses(scores, logits, y_pred, labels, node_adj, s)
    out[f's{s}/scores'] = scores.detach().numpy()
    out['t_st'] = np.array(t_st); out['t_end'] = np.array(t_end)
    t_skip = t_st
    for t_cur in range(t_st, t_end):
        if t_cur < t_skip:
            continue
        s += 1
        step_ref[0] = s
        if feats.size()[0] == 0 and states.size()[0] == 0:
            raise RuntimeError('unexpected re-init in a training chunk')
        y_pred, feats, node_adj, edge_adj, labels = update_graph(
            node_adj, labels, scores, y_pred, X, y, t_cur, use_hungraian=False, mode='train', cuda=False)
        out[f's{s}/op'] = np.array(1); out[f's{s}/t'] = np.array(t_cur)
        out[f's{s}/y_pred'] = y_pred.numpy(); out[f's{s}/feats'] = feats.numpy()
        out[f's{s}/adj'] = coo(node_adj); out[f's{s}/labels'] = labels.numpy()
        scores, logits, states, _ = model(feats, states, node_adj, edge_adj)
        out[f's{s}/logits'] = logits.detach().numpy(); out[f's{s}/h'] = states.detach().numpy()
        scores, lc, lf = losses(scores, logits, y_pred, labels, node_adj, s)
        out[f's{s}/scores'] = scores.detach().numpy()
        loss_c = loss_c + lc; loss_f = loss_f + lf
    loss = loss_c + loss_f
    loss.backward()
    out['loss'] = loss.detach().numpy(); out['loss_c'] = loss_c.detach().numpy(); out['loss_f'] = loss_f.detach().numpy()
    for k, p in model.named_parameters():
        out['g/' + k] = (torch.zeros_like(p) if p.grad is None else p.grad).numpy()
    for k, v in model.state_dict().items():
        if 'running' in k or 'num_batches' in k:
            out['w_after/' + k] = v.detach().cpu().numpy()
    out['n_steps'] = np.array(s + 1)
    out['meta'] = np.array(repr(meta))
    print(f'{name}: {Xn.shape[1]} dets, {s + 1} MP steps, loss {float(loss):.5f}, max rows '
          f'{max(out[f"s{k}/y_pred"].shape[0] for k in range(s + 1))}')
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'gat_train':
    # attention heads in train mode: dropout mask pinned (attention_keep_matrix), gradients of W_att / a included
    torch.set_num_threads(4)
    run_train('train_gat2', 24, 6, 'kitti', '2d', 'diff', 20.0, 0.0, True, nattheads=2)
    run_train('train_gat3_concat', 25, 5, 'kitti', '2d+temp', 'concat', 20.0, 0.0, True, nattheads=3)
elif __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'gat':
    # added after the first batch of fixtures: attention heads (--num-att-heads 2), eval mode
    torch.set_num_threads(4)
    run_infer('infer_gat2', 18, 10, 5, 'kitti', '2d', 'diff', 20.0, 0.0, False, 0, 5, True, nattheads=2)
    run_infer('infer_gat3_concat', 19, 9, 4, 'kitti', '2d+temp', 'concat', 20.0, 0.0, False, 0, 4, True, nattheads=3)
elif __name__ == '__main__':
    torch.set_num_threads(4)
    # stock init: no association is ever made (edge score ~ 0.01), every detection is its own track
    run_infer('infer_stock_kitti', 11, 10, 8, 'kitti', '2d', 'diff', 1.0, None, False, 0, 5, True)
    # decision-exercising weights (all >=2-D params x20, edge-head bias 0): associations, walks, deletion
    run_infer('infer_decide_kitti', 12, 14, 6, 'kitti', '2d', 'diff', 20.0, 0.0, False, 0, 5, True)
    run_infer('infer_decide_ret2', 13, 14, 6, 'kitti', '2d', 'diff', 20.0, 0.0, False, 2, 4, True, prune_at=6)
    run_infer('infer_decide_notp', 14, 12, 6, 'kitti', '2d', 'diff', 20.0, 0.0, False, 0, 5, False)
    run_infer('infer_decide_hung', 15, 12, 6, 'kitti', '2d', 'diff', 20.0, 0.0, True, 0, 5, True)
    run_infer('infer_concat_bdd', 16, 10, 5, 'bdd', '2d', 'concat', 20.0, 0.0, False, 0, 5, True)
    run_infer('infer_groups3', 17, 8, 4, 'kitti', '2d+temp+vis', 'diff', 20.0, 0.0, False, 0, 5, True)
    run_train('train_stock_kitti', 21, 6, 'kitti', '2d', 'diff', 1.0, None, True)
    run_train('train_decide_kitti', 22, 6, 'kitti', '2d', 'diff', 20.0, 0.0, True)
    run_train('train_concat_notp', 23, 5, 'kitti', '2d+temp', 'concat', 20.0, 0.0, False)
