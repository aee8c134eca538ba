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
from .models.loss import _CEFn, create_targets

_I32 = torch.int32


class _Step:
    pass


class TrainBatch:
    def __init__(self, chunks, device=None, builder='slab'):
        """chunks: list of ``(X [1, ND, F], y [1, ND, 2])`` tensors (``y`` = [timestamp, track id]).  builder 'slab':
        all chunks' graphs grow in lock step in one multi-slab graph (one host read-back of the row counts per step);
        'chunk': one chunk at a time (also the fallback when the chunks do not start tracking at the same timestep)."""
        dev = device if device is not None else chunks[0][0].device
        if dev.type != 'cuda':
            raise L.TmpnnError('TrainBatch needs a CUDA device; there is no CPU path')
        fast = self._build_slabs(chunks, dev) if builder == 'slab' else None
        if fast is not None:
            graphs, per_chunk = fast
        else:
            built = [self._chunk_graphs(X.to(dev), y.to(dev)) for X, y in chunks]
            built = [c for c in built if c is not None]
            graphs = [c[0] for c in built]
            per_chunk = [c[1] for c in built]
        self.builder = 'slab' if fast is not None else 'chunk'
        self.device, self.num_chunks = dev, len(per_chunk)
        self.steps = []
        n_steps = max(len(c) for c in per_chunk) if per_chunk else 0
        prev = None
        for i in range(n_steps):
            st = self._concat(i, per_chunk, graphs, prev, dev)
            self.steps.append(st)
            prev = st
        self.edge_rows = sum(int(s.n_edge_rows) for s in self.steps)

    # ---- all chunks in lock step: slabs of one device graph ---------------------------------------------------------
    @staticmethod
    def _build_slabs(chunks, dev):
        """The bookkeeping of ``_chunk_graphs`` for all chunks at once with the batched engine's kernels (train mode:
        teacher-forced association, active set of ``utils/graph.py:271-275``): chunk c owns slab c, every step is one
        index build + associate + append for all slabs and ONE read-back of the row counts.  Returns
        ``(per-chunk graph views, per-chunk step records)`` or None when a chunk cannot be tracked / the chunks do not
        share the first tracked timestep."""
        import ctypes as C
        from .device_graph import FrameTable, SlabGraph, SlabIndex
        B = len(chunks)
        ys = [y[0].detach().cpu().numpy() for _, y in chunks]
        xs = [X[0].detach().cpu().numpy().astype(np.float32) for X, _ in chunks]
        if any((y[:, 1] == -1).all() for y in ys):
            return None
        ft = FrameTable(ys, dev, xs)
        fp = ft.host_frame_ptr
        counts = (fp[:, 1:] - fp[:, :-1]).astype(np.int64)               # [B, t_max + 1]
        if any((counts[c] > 0).sum() < 2 for c in range(B)):
            return None
        cap = max(int(counts[c].sum() + sum(int(counts[c, :t].sum()) * int(counts[c, t]) for t in range(counts.shape[1])))
                  for c in range(B))
        cap = (cap + 63) // 64 * 64
        g = SlabGraph(B, cap, dev, with_labels=True)
        max_dets = int(counts.sum(1).max())
        index = SlabIndex(g, cap_dets=B * max_dets, cap_inc=2 * B * cap)
        z = lambda n: torch.zeros(n, dtype=_I32, device=dev)
        st = {k: z(B) for k in L.SEQ_STATE_FIELDS}
        st_c = L.SeqState(*[L.ptr(st[k]) for k in L.SEQ_STATE_FIELDS])
        cap_new = 2 * B * int(counts.max())
        new_rows, new_x, n_new, n_app = z(cap_new), z(cap_new), z(2), z(B)
        scratch = z(int(L.lib().tmpnn_graph_append_scratch_ints(B, cap)))
        t_dev = z(1)
        per_chunk = [[] for _ in range(B)]
        n_prev = np.zeros(B, dtype=np.int64)

        def append(start, mode):
            L.call('tmpnn_graph_append', g.c, ft.c, C.byref(st_c), L.ptr(t_dev), int(start), 0, int(mode), None, 0,
                   L.ptr(new_rows), L.ptr(new_x), L.ptr(n_new), cap_new, L.ptr(n_app), L.ptr(scratch), L.stream())
            # the one host read-back of the step: row counts, who stepped, how many detections are new
            host = torch.cat((g.n_rows, st['active'], n_new, g.status)).cpu().numpy()
            if host[-1]:
                g.check_status()
            n_rows, active, nn = host[:B].astype(np.int64), host[B:2 * B], int(host[2 * B])
            rows = new_rows[:nn].long()
            order = torch.argsort(rows)                                   # slab-major, ascending rows inside a slab
            rows, xrow = rows[order], new_x[:nn].long()[order]
            per = torch.bincount(rows // cap, minlength=B).cpu().numpy()
            off = 0
            for c in range(B):
                k = int(per[c])
                if active[c]:
                    per_chunk[c].append(dict(n=int(n_rows[c]), n_old=int(n_prev[c]), new_det=(rows[off:off + k] - c * cap).to(_I32),
                                             x_new=ft.x[xrow[off:off + k]],
                                             n_new_edges=int(n_rows[c] - n_prev[c]) - k))
                    n_prev[c] = n_rows[c]
                off += k

        append(1, 0)
        t_first = st['skip_until'].cpu().numpy()
        t_last = st['t_end'].cpu().numpy()
        if len(set(int(t) for t in t_first)) != 1:
            return None
        for t in range(int(t_first[0]), int(t_last.max())):
            t_dev.fill_(t)
            # who is active at t is decided inside the append; the association needs the mask of the sequences still
            # running: t < t_end (a finished chunk keeps its graph)
            running = (st['t_end'] > t).to(_I32)
            index.build(g, running, structured=False)
            L.call('tmpnn_graph_associate', g.c, index.c, 1, L.ptr(running), L.stream())
            append(0, 1)

        class _View:   # chunk c's columns of the slab arrays: what _concat reads
            def __init__(self, c):
                self.g = types.SimpleNamespace(**{k: getattr(g, k)[c * cap:(c + 1) * cap] for k in ('ts', 'det', 'src', 'dst', 'label')})

        return [_View(c) for c in range(B)], per_chunk

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
        ns = torch.tensor([per_chunk[c][i]['n'] for c in act], dtype=torch.int64, device=dev)
        chunk_of_row = torch.repeat_interleave(torch.arange(len(act), device=dev), ns)
        base_of_row = torch.tensor([base[c] for c in act], dtype=_I32, device=dev)[chunk_of_row]
        shift = lambda key: (lambda v: torch.where(v >= 0, v + base_of_row, v))(cat(key))
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
        # rows carried over from the previous step's layout: one contiguous segment per chunk (tmpnn_rows_move)
        carry = None
        if prev is not None:
            seg = np.array([[prev.base[c], base[c], per_chunk[c][i]['n_old']] for c in act], dtype=np.int32)
            t = torch.from_numpy(np.ascontiguousarray(seg.T)).to(dev)
            carry = (t[0].contiguous(), t[1].contiguous(), t[2].contiguous(), len(act), int(seg[:, 2].max()))
        st.new_rows = NewRowGroups(xd, groups, torch.arange(max(1, xo), dtype=_I32, device=dev), carry=carry)
        # loss bookkeeping: targets (labels only), row sets, per-chunk mean weights of the two BCE terms
        labels = g.label[:n].to(torch.int64)
        is_det = g.ts[:n] >= 0
        st.idx_node = torch.nonzero(is_det)[:, 0]
        st.idx_edge = torch.nonzero(~is_det)[:, 0]
        st.targets = create_targets(labels, st.holder, st.idx_node)
        # detections per chunk are known on the host (the new-detection counts add up): no read-back
        nd_c = np.array([sum(int(per_chunk[c][k]['new_det'].numel()) for k in range(i + 1)) for c in act], dtype=np.float64)
        ne_c = np.array([per_chunk[c][i]['n'] for c in act], dtype=np.float64) - nd_c
        wd = torch.from_numpy(1.0 / np.maximum(nd_c, 1)).to(device=dev, dtype=torch.float32)[chunk_of_row]
        we = torch.from_numpy(1.0 / np.maximum(ne_c, 1)).to(device=dev, dtype=torch.float32)[chunk_of_row]
        w = torch.where(is_det, wd, we)
        st.bce_w = w.contiguous()
        st.bce_w_edges = torch.where(is_det, torch.zeros_like(w), w).contiguous()   # without the TP classifier: edge rows only
        st.targets64 = st.targets.to(torch.int64).contiguous()
        st.n_edge_rows = int(ne_c.sum())
        return st


class _WeightedBCE(torch.autograd.Function):
    """sum_i w_i (-log(p_t,i + 1e-10)): FocalLoss(gamma = 0) of every chunk's edge rows and detection rows, each with its own
    per-chunk mean (train.py:76-85), as one kernel pair over all rows of the batch."""

    @staticmethod
    def forward(ctx, p, targets, w):
        n = int(p.numel())
        pc = p.detach().to(torch.float32).contiguous().view(-1)
        per = torch.empty(n, dtype=torch.float32, device=p.device)
        loss = L.zeros(1, p.device)
        L.call('tmpnn_loss_wbce_fwd', n, L.ptr(pc), L.ptr(targets), L.ptr(w), L.ptr(per), L.ptr(loss), L.stream())
        ctx.pc, ctx.t, ctx.w, ctx.n, ctx.shape = pc, targets, w, n, p.shape
        return loss[0]

    @staticmethod
    def backward(ctx, gout):
        g = gout.detach().to(torch.float32).contiguous().view(1)
        dp = torch.empty(ctx.n, dtype=torch.float32, device=ctx.pc.device)
        L.call('tmpnn_loss_wbce_bwd', ctx.n, L.ptr(ctx.pc), L.ptr(ctx.t), L.ptr(ctx.w), L.ptr(g), L.ptr(dp), L.stream())
        return dp.view(ctx.shape), None, None


def batch_loss(model, batch, tp_classifier=True):
    """Forward + losses of every step of ``batch`` (``train.py:65-127`` for all chunks at once): returns the scalar
    loss = sum over chunks of (CE + BCE terms summed over the chunk's steps).  Call ``.backward()`` on it.

    Per step: ONE autograd Function for the whole message-passing step (it also moves the carried rows from the previous
    step's block-diagonal layout into this one), one for the CE term, one for both BCE terms -- no torch indexing ops."""
    params = _param_list(model)
    loss = None
    h_prev = None
    for st in batch.steps:
        scores, logits, h_prev = _MPStepFn.apply(model, st.holder, st.new_rows, h_prev, *params)[:3]
        l = _CEFn.apply(logits, st.targets, st.wg) + \
            _WeightedBCE.apply(scores, st.targets64, st.bce_w if tp_classifier else st.bce_w_edges)
        loss = l if loss is None else loss + l
    return loss


def invalidate_weight_caches(model):
    """Drops the packed weight images cached on ``model`` (functional.packed_cells*, the backward images).  They are keyed by
    the parameters' version counters, which a CUDA-graph replay of an optimizer step does not advance."""
    for k in ('_tmpnn_pack_cache', '_tmpnn_pack_cache_tc', '_tmpnn_pack_cache_tc_node'):
        model.__dict__.pop(k, None)
    sc = model.__dict__.get('_tmpnn_bwd_tc')
    if sc:
        for k in [k for k in sc if isinstance(k, tuple)]:
            sc.pop(k)


class GraphedTrainStep:
    """One optimizer step of the batched trainer -- zero the flat gradient buffer, forward + losses of every message-passing
    step, backward through autograd (the kernels accumulate straight into the flat buffer), Adam -- replayed as CUDA
    graphs: ~650 launches and ~12 ms of host time per step become two replays.

    The step is captured in two graphs so that the data-parallel all-reduce of the flat buffer (NCCL, ``allreduce=True``)
    sits between them: [zero, forward, losses, backward] -> all-reduce -> [Adam].  ``capture()`` first runs three REAL steps
    eagerly (on a side stream, as PyTorch requires before capturing autograd work); the weight images are re-packed inside
    the captured work, so every replay sees the weights the previous replay's optimizer step left.  Training graphs do not
    depend on the model (teacher forcing), so one ``TrainBatch`` can be replayed for any number of steps."""

    def __init__(self, model, batch, lr=1e-4, weight_decay=5e-4, tp_classifier=True, allreduce=False):
        from . import parallel
        self.model, self.batch, self.tp, self.allreduce = model, batch, bool(tp_classifier), bool(allreduce)
        self.flat = parallel.FlatGradients(model)
        self.opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay, capturable=True)  # train.py:329
        self.loss = torch.zeros((), dtype=torch.float32, device=batch.device)
        self._g_fb = self._g_opt = None
        self.arena = L.ZeroArena(batch.device)

    def _forward_backward(self):
        self.flat.zero()
        self.arena.begin()      # one fill for the step's small zero-initialised scratch tensors
        try:
            loss = batch_loss(self.model, self.batch, self.tp)
            loss.backward()
        finally:
            L.ZeroArena.end()
        self.loss.copy_(loss.detach())

    def eager(self):
        self._forward_backward()
        if self.allreduce:
            self.flat.allreduce(average=True)
        self.opt.step()
        return self.loss

    def capture(self, warmup=3):
        L.check(L.lib().tmpnn_init())
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.eager()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        invalidate_weight_caches(self.model)      # the captured forward must contain the re-packing of the weights
        g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            self._forward_backward()
        with torch.cuda.graph(g2, pool=g1.pool()):
            self.opt.step()
        self._g_fb, self._g_opt = g1, g2
        invalidate_weight_caches(self.model)      # the cached images belong to the graph now

    def replay(self):
        if self._g_fb is None:
            self.capture()
        self._g_fb.replay()
        if self.allreduce:
            self.flat.allreduce(average=True)
        self._g_opt.replay()
        return self.loss
