"""Device-resident window graphs: thin owners of the arrays described in ``include/tmpnn.h``.

``SlabGraph`` holds S slabs of ``cap_rows`` rows (structure of arrays, int32 / fp32),
``SlabIndex`` the per-step detection list + incidence CSR + tile table.  Both only
allocate torch tensors and expose the matching ctypes structs; all work on them is done
by the CUDA library.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


class SlabGraph:
    def __init__(self, num_seqs, cap_rows, device, with_labels=True, status=None, deferred=False):
        self.num_seqs, self.cap_rows, self.device = int(num_seqs), int(cap_rows), device
        n = self.num_seqs * self.cap_rows
        i32 = dict(dtype=torch.int32, device=device)
        self.n_rows = torch.zeros(self.num_seqs, **i32)
        self.ts = torch.empty(n, **i32)
        self.det = torch.empty(n, **i32)
        self.ass = torch.empty(n, **i32)
        self.src = torch.empty(n, **i32)
        self.dst = torch.empty(n, **i32)
        self.label = torch.zeros(n, **i32) if with_labels else None
        self.score = torch.zeros(n, dtype=torch.float32, device=device)
        self.logit = torch.zeros(n, dtype=torch.float32, device=device)
        self.status = status if status is not None else torch.zeros(1, **i32)
        # deferred compaction of h (TrackEngine): physical positions of every row and of its endpoints
        self.phys = torch.zeros(n, **i32) if deferred else None
        self.psrc = torch.zeros(n, **i32) if deferred else None
        self.pdst = torch.zeros(n, **i32) if deferred else None
        self.phys_end = torch.zeros(self.num_seqs, **i32) if deferred else None
        self._c = None

    @property
    def c(self):
        if self._c is None:
            self._c = L.Graph(self.num_seqs, self.cap_rows, L.ptr(self.n_rows), L.ptr(self.ts), L.ptr(self.det),
                              L.ptr(self.ass), L.ptr(self.src), L.ptr(self.dst), L.ptr(self.label),
                              L.ptr(self.score), L.ptr(self.logit), L.ptr(self.status), L.ptr(self.phys),
                              L.ptr(self.psrc), L.ptr(self.pdst), L.ptr(self.phys_end))
        return C.byref(self._c)

    def check_status(self):
        """Host read of the sticky flags (synchronises)."""
        flags = int(self.status.item())
        self.notes = flags & L.NOTE_TC_RANGE_RERUN   # handled on the device (the step was re-run on the FMA kernel)
        flags &= ~L.NOTE_TC_RANGE_RERUN
        if flags:
            msgs = [m for b, m in L.FLAG_NAMES.items() if flags & b]
            if flags & 16:
                raise AssertionError('More than one GT edge from same node!')
            raise L.TmpnnError('device-side capacity error: ' + '; '.join(msgs))


class SlabIndex:
    def __init__(self, graph, cap_dets, cap_inc):
        dev = graph.device
        S, cap = graph.num_seqs, graph.cap_rows
        self.cap_dets, self.cap_inc = max(1, int(cap_dets)), max(1, int(cap_inc))
        i32 = dict(dtype=torch.int32, device=dev)
        self.n_dets = torch.zeros(1, **i32)
        self.n_edges = torch.zeros(1, **i32)
        self.det_rows = torch.empty(self.cap_dets, **i32)
        self.det_of_row = torch.empty(S * cap, **i32)
        self.seq_det_ptr = torch.zeros(S + 1, **i32)
        self.seg_ptr = torch.zeros(2 * self.cap_dets + 2, **i32)
        self.inc = torch.empty(self.cap_inc, **i32)
        self.tile_ptr = torch.zeros(S + 1, **i32)
        self.tile128_ptr = torch.zeros(S + 1, **i32)
        self.scratch = torch.empty(int(L.lib().tmpnn_index_scratch_ints(S, cap, self.cap_dets)), **i32)
        self._c = L.Index(self.cap_dets, self.cap_inc, L.ptr(self.n_dets), L.ptr(self.n_edges), L.ptr(self.det_rows),
                          L.ptr(self.det_of_row), L.ptr(self.seq_det_ptr), L.ptr(self.seg_ptr), L.ptr(self.inc),
                          L.ptr(self.tile_ptr), L.ptr(self.tile128_ptr), L.ptr(self.scratch))

    @property
    def c(self):
        return C.byref(self._c)

    def build(self, graph, active=None, structured=False):
        """structured=True: the graph only ever changed through tmpnn_graph_append / tmpnn_graph_decode
        (TrackEngine), so the incidence lists follow from the block boundaries (no atomics, no sort)."""
        if structured:
            if getattr(self, '_scratch2', None) is None:
                nbytes = int(L.lib().tmpnn_index_structured_scratch_bytes(graph.num_seqs, self.cap_dets))
                self._scratch2 = torch.zeros((nbytes + 3) // 4, dtype=torch.int32, device=graph.device)  # boundary counts start at 0
            L.call('tmpnn_index_build_structured', graph.c, self.c, L.ptr(active), L.ptr(self._scratch2), L.stream())
        else:
            L.call('tmpnn_index_build', graph.c, self.c, L.ptr(active), L.stream())


class FrameTable:
    """Detections of S sequences grouped by (sequence, timestamp) -- ``tmpnn_frames``."""

    def __init__(self, ys, device, xs=None):
        """ys: list of [ND_s, 2] arrays ([ts, track_id]); xs: optional list of [ND_s, F] features."""
        S = len(ys)
        ts_list = [np.asarray(y)[:, 0].astype(np.int64) for y in ys]
        self.t_max = int(max([int(t.max()) if t.size else 0 for t in ts_list] + [0]))
        T2 = self.t_max + 2
        frame_ptr = np.zeros(S * T2, np.int32)
        det_ptr = np.zeros(S + 1, np.int32)
        frame_dets, tracks = [], []
        off = 0
        for s, (y, ts) in enumerate(zip(ys, ts_list)):
            order = np.argsort(ts, kind='stable').astype(np.int32)  # ascending detection ids inside a timestep
            counts = np.bincount(ts[ts >= 0], minlength=self.t_max + 1) if ts.size else np.zeros(self.t_max + 1, np.int64)
            assert (ts >= 0).all(), 'negative timestamps are not supported'
            fp = off + np.concatenate(([0], np.cumsum(counts)))
            frame_ptr[s * T2:(s + 1) * T2] = fp
            frame_dets.append(order)
            tracks.append(np.asarray(y)[:, 1].astype(np.int32))
            off += ts.size
            det_ptr[s + 1] = off
        self.host_frame_ptr = frame_ptr.reshape(S, T2)
        self.num_seqs = S
        self.total_dets = off
        cat = lambda a, dt: np.concatenate(a).astype(dt) if off else np.zeros(0, dt)
        self.frame_ptr = torch.from_numpy(frame_ptr).to(device)
        self.frame_dets = torch.from_numpy(cat(frame_dets, np.int32)).to(device)
        self.det_ptr = torch.from_numpy(det_ptr).to(device)
        self.det_track = torch.from_numpy(cat(tracks, np.int32)).to(device)
        self.host_det_ptr = det_ptr
        self.x = None
        self.ldx = 0
        if xs is not None:
            self.x = torch.cat([torch.as_tensor(x, dtype=torch.float32) for x in xs], 0).to(device).contiguous()
            self.ldx = int(self.x.shape[1])
        self._c = L.Frames(self.t_max, self.ldx, L.ptr(self.frame_ptr), L.ptr(self.frame_dets), L.ptr(self.det_ptr),
                           L.ptr(self.det_track))

    @property
    def c(self):
        return C.byref(self._c)

    def count(self, s, t):
        if t < 0 or t > self.t_max:
            return 0
        return int(self.host_frame_ptr[s, t + 1] - self.host_frame_ptr[s, t])


class WindowGraph:
    """One sequence's window graph (S = 1) as handed around by the drop-in API.

    The reference threads ``y_pred`` / ``node_adj`` / ``labels`` tensors from call to call;
    our ``node_adj`` tensors additionally carry one of these (attribute ``_tmpnn``) so the
    next call can skip the COO -> edge-list conversion.
    """

    def __init__(self, n, cap, device, with_labels=True):
        self.g = SlabGraph(1, max(1, int(cap)), device, with_labels)
        self.n = int(n)
        self.g.n_rows.fill_(self.n)
        self._index = None

    @property
    def device(self):
        return self.g.device

    def index(self):
        if self._index is None:
            self._index = SlabIndex(self.g, cap_dets=max(1, self.n), cap_inc=max(1, 2 * self.n))
            self._index.build(self.g)
        return self._index

    def invalidate(self):
        self._index = None

    # ---- conversions to / from the reference's tensors --------------------------------
    @staticmethod
    def from_tensors(y_pred, node_adj, labels=None, scores=None, extra_cap=0):
        """Builds the device graph from the reference-style tensors of one call."""
        dev = _cuda_device(y_pred, node_adj)
        n = int(y_pred.shape[0])
        wg = WindowGraph(n, n + int(extra_cap), dev, with_labels=True)
        g = wg.g
        if n:
            yp = y_pred.to(device=dev, dtype=torch.int64).contiguous()
            L.call('tmpnn_ypred_unpack', L.ptr(yp), n, L.ptr(g.ts), L.ptr(g.det), L.ptr(g.ass), L.stream())
            h = getattr(node_adj, '_tmpnn', None)
            if h is not None and h.n == n:
                g.src[:n].copy_(h.g.src[:n])
                g.dst[:n].copy_(h.g.dst[:n])
            else:
                sp = node_adj if node_adj.is_sparse else node_adj.to_sparse()
                sp = sp.to(dev)
                idx = sp._indices().contiguous()
                val = sp._values().to(torch.float32).contiguous()
                L.call('tmpnn_edges_from_coo', L.ptr(idx), L.ptr(val), int(val.numel()), n, L.ptr(g.src), L.ptr(g.dst),
                       L.stream())
            if labels is not None:
                g.label[:n].copy_(labels.to(device=dev, dtype=torch.int32))
            if scores is not None:
                g.score[:n].copy_(scores.to(device=dev, dtype=torch.float32)[:, 1])
        return wg

    def y_pred(self):
        out = torch.empty((self.n, 3), dtype=torch.int64, device=self.device)
        if self.n:
            L.call('tmpnn_ypred_pack', L.ptr(self.g.ts), L.ptr(self.g.det), L.ptr(self.g.ass), self.n, L.ptr(out), L.stream())
        return out

    def labels(self):
        return self.g.label[:self.n].to(torch.int64)

    def adjacency(self, transpose=False):
        """Sparse COO ``node_adj`` (with I_node) or ``edge_adj`` (= node_adj^T off-diagonal + I_edge),
        built in O(nnz) from the edge list (reference ``utils/graph.py:152-163, 299-308``)."""
        n = self.n
        n_det = int((self.g.ts[:n] >= 0).sum().item()) if n else 0
        n_edge = n - n_det
        nnz = 3 * n_edge if transpose else 2 * n_edge + n_det
        idx = torch.empty((2, nnz), dtype=torch.int64, device=self.device)
        val = torch.empty(nnz, dtype=torch.float32, device=self.device)
        if n:
            scratch = torch.empty(2 * n + 8 + (n + 2048) // 2048, dtype=torch.int32, device=self.device)
            L.call('tmpnn_coo_from_edges', L.ptr(self.g.ts), L.ptr(self.g.src), L.ptr(self.g.dst), n, int(transpose),
                   L.ptr(idx), L.ptr(val), nnz, L.ptr(scratch), L.stream())
        adj = torch.sparse_coo_tensor(idx, val, (n, n), check_invariants=False, is_coalesced=not transpose)
        adj._tmpnn = self
        return adj


def _cuda_device(*tensors):
    for t in tensors:
        if t is not None and getattr(t, 'is_cuda', False):
            return t.device
    if not torch.cuda.is_available():
        raise L.TmpnnError('trackmpnn_b200 needs a CUDA device (B200, sm_100a); there is no CPU path')
    return torch.device('cuda', torch.cuda.current_device())


def window_graph_of(node_adj):
    """The device graph behind a ``node_adj`` tensor: the attached handle when it came from our
    own graph functions, else rebuilt from the COO entries (detections are the rows with a
    diagonal entry, reference ``models/track_mpnn.py:55``)."""
    h = getattr(node_adj, '_tmpnn', None)
    if h is not None:
        return h
    dev = _cuda_device(node_adj)
    n = int(node_adj.shape[0])
    sp = (node_adj if node_adj.is_sparse else node_adj.to_sparse()).to(dev)
    idx = sp._indices().contiguous()
    val = sp._values().to(torch.float32).contiguous()
    wg = WindowGraph(n, n, dev, with_labels=False)
    if n:
        L.call('tmpnn_edges_from_coo', L.ptr(idx), L.ptr(val), int(val.numel()), n, L.ptr(wg.g.src), L.ptr(wg.g.dst),
               L.stream())
        wg.g.ts.fill_(-1)
        diag = (idx[0] == idx[1]) & (val != 0)
        wg.g.ts[idx[0][diag]] = 0
        wg.g.det.fill_(-1)
        wg.g.ass.fill_(-1)
    return wg
