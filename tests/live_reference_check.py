"""Runs the UNMODIFIED reference (``/root/reference``, imported read-only) and the oracle side by side on fresh random
streams -- not the committed fixtures -- and prints one JSON line per case.  Executed by
``tests/test_oracle_live.py`` in a subprocess (the reference's top-level module names ``models`` / ``utils`` must not
leak into the pytest process); only in the build container, the GPU box has no reference tree.

    python tests/live_reference_check.py SEED [SEED ...]            # inference loops
    python tests/live_reference_check.py train SEED [SEED ...]      # one BPTT chunk each: loss and all gradients
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, '/root/reference')

from models.track_mpnn import TrackMPNN  # noqa: E402  (reference)
from models.loss import create_targets, CELoss, FocalLoss  # noqa: E402  (reference)
from utils.graph import initialize_graph, update_graph, decode_tracks  # noqa: E402  (reference)
from oracle.infer_loop import run_infer  # noqa: E402
from trackmpnn_b200 import synth  # noqa: E402


def reference_loop(model, X, y, cur_win, ret_win, hungarian):
    """infer.py:48-87 with the 4-value unpack of forward."""
    y_out = y[0].numpy().astype(np.int64)
    y_out[:, 1] = -1
    fix = lambda s: torch.cat((1 - s, s), dim=1)
    edges = frames = 0
    with torch.no_grad():
        y_pred, feats, node_adj, edge_adj, labels, t_st, t_end = initialize_graph(X, y, t_st=0, mode='test', cuda=False)
        if y_pred is None:
            return y_out, 0, 0
        scores, logits, states, _ = model(feats, None, node_adj, edge_adj)
        edges += int((y_pred[:, 0] == -1).sum())
        scores = fix(scores)
        t_skip = t_st
        for t_cur in range(t_st, t_end):
            if t_cur < t_skip:
                continue
            if feats.size()[0] == 0 and states.size()[0] == 0:
                y_pred, feats, node_adj, edge_adj, labels, t_skip, _ = initialize_graph(X, y, t_st=t_cur, mode='test', cuda=False)
                if y_pred is None:
                    break
                states = None
            else:
                y_pred, feats, node_adj, edge_adj, labels = update_graph(node_adj, labels, scores, y_pred, X, y, t_cur,
                                                                         use_hungraian=hungarian, mode='test', cuda=False)
            scores, logits, states, _ = model(feats, states, node_adj, edge_adj)
            edges += int((y_pred[:, 0] == -1).sum())
            scores = fix(scores)
            t_upto = t_end if t_cur == t_end - 1 else t_cur - cur_win + 2
            y_pred, y_out, states, node_adj, labels, scores = decode_tracks(states, node_adj, labels, scores, y_pred, y_out,
                                                                            t_upto, ret_win, use_hungraian=hungarian, cuda=False)
            frames += 1
    return y_out, edges, frames


def case(seed):
    rs = np.random.RandomState(seed)
    msg_type = ('diff', 'concat')[rs.randint(2)]
    hungarian = bool(rs.randint(2))
    cur_win, ret_win = int(rs.randint(3, 6)), int(rs.randint(0, 3))
    frames, dets = int(rs.randint(6, 12)), int(rs.randint(2, 6))
    timestamps = None
    if rs.rand() < 0.3:   # a hole longer than the window: the re-initialisation path
        timestamps = list(range(0, 3)) + list(range(3 + cur_win + 2, 3 + cur_win + 2 + frames))
    Xn, yn = synth.make_sequence(seed, frames, dets, 'kitti', timestamps=timestamps)
    torch.manual_seed(seed)
    model = TrackMPNN(features='2d', ncategories=3, nhidden=64, nattheads=0, msg_type=msg_type)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() >= 2:
                p.mul_(20.0)
        model.output_transform_edge.bias.fill_(0.0)
    model.eval()
    want, edges, nfr = reference_loop(model, torch.from_numpy(Xn), torch.from_numpy(yn), cur_win, ret_win, hungarian)
    params = {k: v.numpy() for k, v in model.state_dict().items()}
    got, st = run_infer(params, Xn, yn, msg_type=msg_type, cur_win_size=cur_win, ret_win_size=ret_win,
                        use_hungarian=hungarian, record_margin=True)
    return dict(seed=seed, msg_type=msg_type, hungarian=hungarian, cur_win=cur_win, ret_win=ret_win, dets=int(Xn.shape[1]),
                tracks=len(set(want[:, 1].tolist()) - {-1}), margin=float(st['margin']),
                same_tracks=bool(np.array_equal(got, want)), same_edges=bool(edges == st['edge_updates']),
                same_frames=bool(nfr == st['frames']))


def train_case(seed):
    """train.py:65-134 for one chunk on the reference vs oracle/train_ref.py: loss and every parameter gradient."""
    from oracle import train_ref as T
    rs = np.random.RandomState(seed)
    msg_type = ('diff', 'concat')[rs.randint(2)]
    tp = bool(rs.randint(2))
    ts = synth.train_chunk_timestamps(seed, 5, int(rs.randint(2, 5)))
    Xn, yn = synth.make_sequence(seed, None, int(rs.randint(3, 7)), 'kitti', timestamps=ts)
    torch.manual_seed(seed)
    model = TrackMPNN(features='2d', ncategories=3, nhidden=64, nattheads=0, msg_type=msg_type)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() >= 2:
                p.mul_(10.0)
        model.output_transform_edge.bias.fill_(0.0)
    model.train()
    params = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
    X, y = torch.from_numpy(Xn), torch.from_numpy(yn)
    ce, fn, fe = CELoss(), FocalLoss(gamma=0), FocalLoss(gamma=0)

    def losses(scores, logits, y_pred, labels, node_adj):
        idx_edge = torch.nonzero((y_pred[:, 0] == -1))[:, 0]
        idx_node = torch.nonzero((y_pred[:, 0] != -1))[:, 0]
        targets = create_targets(labels, node_adj, idx_node)
        l = ce(logits, targets, node_adj, idx_node) + fe(scores[idx_edge, 0], targets[idx_edge])
        if tp:
            l = l + fn(scores[idx_node, 0], targets[idx_node])
        return torch.cat((1 - scores, scores), dim=1), l

    y_pred, feats, node_adj, edge_adj, labels, t_st, t_end = initialize_graph(X, y, t_st=0, mode='train', cuda=False)
    scores, logits, states, _ = model(feats, None, node_adj, edge_adj)
    scores, loss = losses(scores, logits, y_pred, labels, node_adj)
    for t_cur in range(t_st, t_end):
        y_pred, feats, node_adj, edge_adj, labels = update_graph(node_adj, labels, scores, y_pred, X, y, t_cur,
                                                                 use_hungraian=False, mode='train', cuda=False)
        scores, logits, states, _ = model(feats, states, node_adj, edge_adj)
        scores, l = losses(scores, logits, y_pred, labels, node_adj)
        loss = loss + l
    loss.backward()
    out = T.train_chunk(params, Xn, yn, msg_type=msg_type, tp_classifier=tp)
    gmax = max(float(p.grad.abs().max()) for p in model.parameters() if p.grad is not None)
    worst = 0.0
    for k, p in model.named_parameters():
        want = np.zeros(tuple(p.shape), np.float32) if p.grad is None else p.grad.numpy()
        err = float(np.abs(out['grads'][k] - want).max())
        worst = max(worst, err / (2e-3 * float(np.abs(want).max()) + 1e-6 * gmax + 1e-9))   # the bar of assert_grads_close
    return dict(seed=seed, train=True, msg_type=msg_type, tp_classifier=tp, dets=int(Xn.shape[1]), steps=len(out['graphs']),
                loss=float(loss), loss_rel_err=abs(out['loss'] - float(loss)) / max(1.0, abs(float(loss))), grad_err_over_bar=worst)


if __name__ == '__main__':
    torch.set_num_threads(2)
    args = sys.argv[1:]
    fn = case
    if args and args[0] == 'train':
        fn, args = train_case, args[1:]
    for sd in args:
        print(json.dumps(fn(int(sd))), flush=True)
