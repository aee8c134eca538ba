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

_I32 = torch.int32


class _Step:
    pass


class TrainBatch:
    def __init__(self, chunks, device=None):
        """chunks: list of ``(X [1, ND, F], y [1, ND, 2])`` tensors (``y`` = [timestamp, track id])."""
        dev = device if device is not None else chunks[0][0].device
        if dev.type != 'cuda':
            raise L.TmpnnError('TrainBatch needs a CUDA device; there is no CPU path')
        built = [self._chunk_graphs(X.to(dev), y.to(dev)) for X, y in chunks]
        built = [c for c in built if c is not None]
        graphs = [c[0] for c in built]
        per_chunk = [c[1] for c in built]
        self.device, self.num_chunks = dev, len(per_chunk)
        self.steps = []
        n_steps = max(len(c) for c in per_chunk) if per_chunk else 0
        prev = None
        for i in range(n_steps):
            st = self._concat(i, per_chunk, graphs, prev, dev)
            self.steps.append(st)
            prev = st
        self.edge_rows = sum(int(s.n_edge_rows) for s in self.steps)

    # ---- one chunk: its window graph at every step (labels only) --------------------------------------------------
    @staticmethod
    def _chunk_graphs(X, y):
        """The training loop's graph bookkeeping for one chunk (``train.py:65,92-104``: ``initialize_graph`` then
        ``update_graph(mode='train')`` for every integer timestep), on ONE growing single-slab graph: training never
        deletes rows, so the graph of step i is the first ``n_i`` rows of the final one.  Returns ``(graph, steps)``
        or None when the reference would skip the chunk (``utils/graph.py:132-133``)."""
        from .utils.graph import _append, _associate, _frames_of, _seq_state
        dev = X.device
        if bool((y[0, :, 1] == -1).all()):
            return None
        ft = _frames_of(y, dev)
        fp = ft.host_frame_ptr[0]
        counts = np.diff(fp[:ft.t_max + 2]).astype(np.int64)
        nz = [t for t in range(len(counts)) if counts[t] > 0]
        if len(nz) < 2:
            return None
        # every detection can at most be connected to every later one
        cap = int(counts.sum() + sum(int(counts[:t].sum()) * int(counts[t]) for t in range(len(counts))))
        wg = WindowGraph(0, cap, dev, with_labels=True)
        st_t, st_c = _seq_state(dev)
        Xd = X[0].to(torch.float32)
        steps = []

        def snap(n_old, new_rows, new_x):
            steps.append(dict(n=wg.n, n_old=n_old, new_det=new_rows.clone(), x_new=Xd[new_x.long()],
                              n_new_edges=(wg.n - n_old) - int(new_rows.numel())))

        new_rows, new_x = _append(wg, ft, st_c, 0, 1, 0, int(counts[nz[0]] + counts[nz[1]]))
        t_st, t_end = int(st_t['skip_until'].item()), int(st_t['t_end'].item())
        snap(0, new_rows, new_x)
        for t in range(t_st, t_end):
            n_old = wg.n
            _associate(wg, None, False, 'train')          # teacher forcing from the labels (utils/graph.py:229-245)
            new_rows, new_x = _append(wg, ft, None, t, 0, 1, ft.count(0, t))
            snap(n_old, new_rows, new_x)
        return wg, steps

    # ---- one step: block-diagonal concatenation of the active chunks' graphs ----------------------------------
    @staticmethod
    def _concat(i, per_chunk, graphs, prev, dev):
        act = [c for c in range(len(per_chunk)) if len(per_chunk[c]) > i]
        st = _Step()
        st.chunks = act
        base, off = {}, 0
        for c in act:
            base[c] = off
            off += per_chunk[c][i]['n']
        n = off
        st.n, st.base = n, base
        col = lambda c, key: getattr(graphs[c].g, key)[:per_chunk[c][i]['n']]   # step i = a prefix of the final graph
        cat = lambda key: torch.cat([col(c, key) for c in act])
        shift = lambda key: torch.cat([torch.where(col(c, key) >= 0, col(c, key) + base[c], col(c, key)) for c in act])
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
