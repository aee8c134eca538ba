"""Batched training: B chunks advance in lock step through forward, losses and backward.

The reference trains one chunk at a time (``train.py:49-173``: batch size 1, ~10 message-passing steps per
chunk, one ``loss.backward()``); at its chunk sizes every kernel is launch-latency bound (~2600 launches per
chunk).  Here the window graphs of B chunks are concatenated per step into ONE block-diagonal single-slab graph,
so every kernel of the step -- aggregation, the fused GRU step with stored gates, targets / CE, the whole backward
-- runs once per step for all chunks, through the same C-ABI entry points and autograd Function as the drop-in path.

What stays per chunk, because the reference's semantics are per chunk:
  * BatchNorm statistics of the input transform (one chunk = one batch, ``models/track_mpnn.py:59``; no SyncBN),
  * the means of the BCE terms (``train.py:76-85``: one mean over the chunk's edge rows, one over its detections).
The loss of a batch is the sum of its chunks' losses, so its gradient is the sum of the per-chunk gradients.

Training graphs never depend on the model (teacher forcing from the labels, ``utils/graph.py:229-245``), so a
``TrainBatch`` is built once per set of chunks from the drop-in graph functions and can be replayed for any number
of optimizer steps.
"""
import types

import numpy as np
import torch

from . import _lib as L
from .device_graph import WindowGraph
from .functional import H, NewRowGroups, _MPStepFn, _param_list
from .models.loss import CELoss, create_targets
from .utils.graph import initialize_graph, update_graph

_I32 = torch.int32


class _Step:
    pass


class TrainBatch:
    def __init__(self, chunks, device=None):
        """chunks: list of ``(X [1, ND, F], y [1, ND, 2])`` tensors (``y`` = [timestamp, track id])."""
        dev = device if device is not None else chunks[0][0].device
        if dev.type != 'cuda':
            raise L.TmpnnError('TrainBatch needs a CUDA device; there is no CPU path')
        per_chunk = [self._chunk_graphs(X.to(dev), y.to(dev)) for X, y in chunks]
        per_chunk = [c for c in per_chunk if c]
        self.device, self.num_chunks = dev, len(per_chunk)
        self.steps = []
        n_steps = max(len(c) for c in per_chunk) if per_chunk else 0
        prev = None
        for i in range(n_steps):
            st = self._concat(i, per_chunk, prev, dev)
            self.steps.append(st)
            prev = st
        self.edge_rows = sum(int(s.n_edge_rows) for s in self.steps)

    # ---- one chunk: its window graph at every step (drop-in graph functions, labels only) ----------------
    @staticmethod
    def _chunk_graphs(X, y):
        y_pred, feats, node_adj, edge_adj, labels, t_st, t_end = initialize_graph(X, y, 0, 'train', True)
        if y_pred is None:
            return []
        out = []

        def snap(node_adj, feats, n_old):
            wg = node_adj._tmpnn
            n = wg.n
            g = wg.g
            ts = g.ts[:n].clone()
            new_det = torch.nonzero(ts[n_old:] >= 0)[:, 0]
            out.append(dict(n=n, n_old=n_old, ts=ts, det=g.det[:n].clone(), src=g.src[:n].clone(), dst=g.dst[:n].clone(),
                            label=g.label[:n].clone(), new_det=(new_det + n_old).to(_I32),
                            x_new=feats[new_det].to(torch.float32), n_new_edges=(n - n_old) - int(new_det.numel())))

        snap(node_adj, feats, 0)
        for t in range(t_st, t_end):
            n_old = int(y_pred.shape[0])
            dummy = torch.zeros((n_old, 2), dtype=torch.float32, device=X.device)   # train mode never reads the scores
            y_pred, feats, node_adj, edge_adj, labels = update_graph(node_adj, labels, dummy, y_pred, X, y, t,
                                                                     use_hungraian=False, mode='train', cuda=True)
            snap(node_adj, feats, n_old)
        return out

    # ---- one step: block-diagonal concatenation of the active chunks' graphs ----------------------------------
    @staticmethod
    def _concat(i, per_chunk, prev, dev):
        act = [c for c in range(len(per_chunk)) if len(per_chunk[c]) > i]
        st = _Step()
        st.chunks = act
        base, off = {}, 0
        for c in act:
            base[c] = off
            off += per_chunk[c][i]['n']
        n = off
        st.n, st.base = n, base
        cat = lambda key: torch.cat([per_chunk[c][i][key] for c in act])
        shift = lambda key: torch.cat([torch.where(per_chunk[c][i][key] >= 0, per_chunk[c][i][key] + base[c],
                                                   per_chunk[c][i][key]) for c in act])
        wg = WindowGraph(n, n, dev, with_labels=True)
        g = wg.g
        g.ts[:n] = cat('ts'); g.det[:n] = cat('det'); g.ass[:n] = -1
        g.src[:n] = shift('src'); g.dst[:n] = shift('dst'); g.label[:n] = cat('label')
        st.wg, st.holder = wg, types.SimpleNamespace(_tmpnn=wg)
        # new detection rows, chunk by chunk
        xs, groups, xo = [], [], 0
        for c in act:
            d = per_chunk[c][i]
            nd = int(d['new_det'].numel())
            groups.append((torch.arange(xo, xo + nd, dtype=_I32, device=dev), (d['new_det'] + base[c]).to(_I32).contiguous(),
                           nd, int(d['n_new_edges'])))
            xs.append(d['x_new'])
            xo += nd
        xd = torch.cat(xs).contiguous() if xo else torch.zeros((1, per_chunk[act[0]][i]['x_new'].shape[1]), device=dev)
        st.new_rows = NewRowGroups(xd, groups, torch.arange(max(1, xo), dtype=_I32, device=dev))
        # rows carried over from the previous step's layout
        if prev is not None:
            src_idx, dst_idx = [], []
            for c in act:
                n_old = per_chunk[c][i]['n_old']
                src_idx.append(np.arange(prev.base[c], prev.base[c] + n_old))
                dst_idx.append(np.arange(base[c], base[c] + n_old))
            st.carry_from = torch.from_numpy(np.concatenate(src_idx)).to(dev)
            st.carry_to = torch.from_numpy(np.concatenate(dst_idx)).to(dev)
        # loss bookkeeping: targets (labels only), row sets, per-chunk mean weights of the two BCE terms
        labels = g.label[:n].to(torch.int64)
        is_det = g.ts[:n] >= 0
        st.idx_node = torch.nonzero(is_det)[:, 0]
        st.idx_edge = torch.nonzero(~is_det)[:, 0]
        st.targets = create_targets(labels, st.holder, st.idx_node)
        w = torch.zeros(n, dtype=torch.float32, device=dev)
        for c in act:
            d = per_chunk[c][i]
            rows = slice(base[c], base[c] + d['n'])
            det_c = is_det[rows]
            nd_c, ne_c = int(det_c.sum()), d['n'] - int(det_c.sum())
            wc = torch.where(det_c, torch.full((), 1.0 / max(1, nd_c), device=dev), torch.full((), 1.0 / max(1, ne_c), device=dev))
            w[rows] = wc
        st.bce_w = w
        st.n_edge_rows = int(st.idx_edge.numel())
        return st


def batch_loss(model, batch, tp_classifier=True):
    """Forward + losses of every step of ``batch`` (``train.py:65-127`` for all chunks at once): returns the scalar
    loss = sum over chunks of (CE + BCE terms summed over the chunk's steps).  Call ``.backward()`` on it."""
    ce = CELoss()
    params = _param_list(model)
    loss = None
    h_prev = None
    ldh = len(model.feature_idx) * H
    for st in batch.steps:
        h_in = torch.zeros((st.n, ldh), dtype=torch.float32, device=batch.device)
        if h_prev is not None:
            h_in = h_in.index_copy(0, st.carry_to, h_prev.index_select(0, st.carry_from))
        scores, logits, h_prev = _MPStepFn.apply(model, st.holder, st.new_rows, h_in, *params)
        p = scores[:, 0]
        tgt = st.targets.to(p.dtype)
        p_t = p * tgt + (1 - p) * (1 - tgt)                      # FocalLoss(gamma=0): mean(-log(p_t + 1e-10)), per chunk
        bce = -(torch.log(p_t + 1e-10) * st.bce_w)
        l = ce(logits, st.targets, st.holder, st.idx_node) + bce[st.idx_edge].sum()
        if tp_classifier:
            l = l + bce[st.idx_node].sum()
        loss = l if loss is None else loss + l
    return loss
