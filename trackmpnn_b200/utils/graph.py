"""Drop-in replacements for the free functions of reference ``utils/graph.py``.

Same names, positional order, keyword names (including ``use_hungraian``), return tuples and
error behaviour; tensors in, tensors out.  All graph logic runs in ``libtmpnn_sm100a.so`` on
the edge list -- no dense N x N matrix is ever built.  The returned ``node_adj`` / ``edge_adj``
are genuine sparse COO tensors (built in O(nnz)) that also carry the device graph
(``._tmpnn``) so the next call skips the COO -> edge-list conversion.

These wrappers serve one sequence per call and read a couple of counts back to size their
outputs; the batched, sync-free form is ``trackmpnn_b200.engine.TrackEngine``.
"""
import numpy as np
import torch

from .. import _lib as L
from ..device_graph import FrameTable, SlabGraph, WindowGraph, _cuda_device

_I32 = torch.int32


def _frames_of(y, device):
    """Frame table of one sequence, cached on the y tensor (the drivers pass the same y every frame)."""
    cached = getattr(y, '_tmpnn_frames', None)
    if cached is not None and cached[0] == y._version:
        return cached[1]
    ft = FrameTable([y[0].detach().cpu().numpy()], device)
    try:
        y._tmpnn_frames = (y._version, ft)
    except Exception:
        pass
    return ft


def _place(t, cuda, device):
    if t is None:
        return None
    return t if cuda else t.cpu()


def _seq_state(device):
    z = lambda: torch.zeros(1, dtype=_I32, device=device)
    t = dict(phase=z(), skip_until=z(), t_end=z(), active=z(), t_upto=z(), fresh=z(), last_new=z())
    c = L.SeqState(*[L.ptr(t[k]) for k in L.SEQ_STATE_FIELDS])
    return t, c


def _append(wg, ft, st_c, t, start, mode, cap_new):
    """Runs tmpnn_graph_append on a single-slab graph; returns (rows of new dets, feature rows of new dets)."""
    import ctypes as C
    dev = wg.device
    g = wg.g
    new_rows = torch.empty(max(1, cap_new), dtype=_I32, device=dev)
    new_x = torch.empty(max(1, cap_new), dtype=_I32, device=dev)
    n_new = torch.zeros(2, dtype=_I32, device=dev)
    n_app = torch.zeros(1, dtype=_I32, device=dev)
    scratch = torch.empty(int(L.lib().tmpnn_graph_append_scratch_ints(1, g.cap_rows)), dtype=_I32, device=dev)
    t_dev = torch.tensor([int(t)], dtype=_I32, device=dev)
    L.call('tmpnn_graph_append', g.c, ft.c, C.byref(st_c) if st_c is not None else None, L.ptr(t_dev), int(start), 0,
           int(mode), None, 0, L.ptr(new_rows), L.ptr(new_x), L.ptr(n_new), int(cap_new), L.ptr(n_app), L.ptr(scratch),
           L.stream())
    g.check_status()
    wg.n = int(g.n_rows.item())
    wg.invalidate()
    nn_ = int(n_new[0].item())
    return new_rows[:nn_], new_x[:nn_]


def _feats(wg, n_old, X, new_rows, new_x):
    """Input features of the new rows: zeros for edge rows, X[0, det] for detection rows."""
    n_new = wg.n - n_old
    Xd = X[0]
    feats = torch.zeros((n_new, X.size()[2]), dtype=X.dtype, device=X.device)
    if new_rows.numel():
        feats[(new_rows.long() - n_old).to(X.device)] = Xd[new_x.long().to(X.device)]
    return feats


def initialize_graph(X, y, t_st=0, mode='test', cuda=True):
    """Reference ``utils/graph.py:96-186``: graph over the first two non-empty timesteps >= t_st.

    Returns ``(y_pred, feats, node_adj, edge_adj, labels, t1+1, tN+1)`` or 7 x ``None``."""
    assert (X.size()[0] == y.size()[0] == 1), "Only batch size 1 supported!"
    assert (X.size()[1] == y.size()[1]), "Input dimension mismatch!"
    dev = _cuda_device(X, y)
    if (y[0, :, 1] == -1).all() and mode == 'train':
        return None, None, None, None, None, None, None
    ft = _frames_of(y, dev)
    # sizes of the two frames are known on the host (y is host-visible metadata)
    fp = ft.host_frame_ptr[0]
    nz = [t for t in range(int(t_st), ft.t_max + 1) if fp[t + 1] > fp[t]]
    if len(nz) < 2:
        return None, None, None, None, None, None, None
    n0, n1 = int(fp[nz[0] + 1] - fp[nz[0]]), int(fp[nz[1] + 1] - fp[nz[1]])
    total = n0 + n0 * n1 + n1
    wg = WindowGraph(0, total, dev, with_labels=True)
    st_t, st_c = _seq_state(dev)
    new_rows, new_x = _append(wg, ft, st_c, t_st, 1, 0, n0 + n1)
    assert wg.n == total
    t1p, tNp = int(st_t['skip_until'].item()), int(st_t['t_end'].item())
    feats = _feats(wg, 0, X, new_rows, new_x)
    y_pred = wg.y_pred()
    node_adj, edge_adj, labels = wg.adjacency(False), wg.adjacency(True), wg.labels()
    dup = _duplicate_track(y, nz[0])
    assert not dup, "More than one detection from same timestep assinged to same track!"
    return (_place(y_pred, cuda, dev), feats, _place_adj(node_adj, cuda), _place_adj(edge_adj, cuda),
            _place(labels, cuda, dev), t1p, tNp)


def _duplicate_track(y, t0):
    tr = y[0, y[0, :, 0] == t0, 1]
    tr = tr[tr >= 0]
    return bool(tr.numel() != torch.unique(tr).numel())


def _place_adj(adj, cuda):
    if cuda:
        return adj
    h = adj._tmpnn
    out = adj.cpu()
    out._tmpnn = h
    return out


def update_graph(node_adj, labels, scores, y_pred, X, y, t, use_hungraian=True, mode='test', cuda=True):
    """Reference ``utils/graph.py:189-334``: re-associate, then append timestep t.

    Returns ``(y_pred', feats_new, node_adj', edge_adj', labels')``."""
    assert (X.size()[0] == y.size()[0] == 1), "Only batch size 1 supported!"
    assert (X.size()[1] == y.size()[1]), "Input dimension mismatch!"
    dev = _cuda_device(y_pred, node_adj, scores)
    ft = _frames_of(y, dev)
    nt = ft.count(0, int(t))
    n = int(y_pred.shape[0])
    n_det = int((y_pred[:, 0] >= 0).sum().item())
    wg = WindowGraph.from_tensors(y_pred, node_adj, labels, scores, extra_cap=n_det * nt + nt)
    _associate(wg, scores, use_hungraian, mode)
    new_rows, new_x = _append(wg, ft, None, t, 0, 1 if mode == 'train' else 0, nt)
    feats = _feats(wg, n, X, new_rows, new_x)
    out_labels = None if labels is None else _place(wg.labels(), cuda, dev)
    return (_place(wg.y_pred(), cuda, dev), feats, _place_adj(wg.adjacency(False), cuda),
            _place_adj(wg.adjacency(True), cuda), out_labels)


def _hungarian(wg, scores, only_t=False, t=0, threshold=0.5):
    """tmpnn_graph_associate_hungarian on a single-slab graph; the cost column is scores[:, 0] as the
    reference reads it (utils/graph.py:81)."""
    dev = wg.device
    ix = wg.index()
    nd = max(1, int(ix.n_dets.item()))
    cost = scores.detach().to(device=dev, dtype=torch.float32)[:, 0].contiguous()
    nbytes = int(L.lib().tmpnn_hungarian_scratch_bytes(1, nd))
    scratch = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=dev)
    L.call('tmpnn_graph_associate_hungarian', wg.g.c, ix.c, L.ptr(cost), None, nd, int(only_t), int(t), float(threshold),
           L.ptr(scratch), L.stream())
    wg.g.check_status()


def _associate(wg, scores, use_hungarian, mode):
    if mode == 'train':
        L.call('tmpnn_graph_associate', wg.g.c, wg.index().c, 1, None, L.stream())
        wg.g.check_status()
    elif use_hungarian:
        _hungarian(wg, scores)
    else:
        L.call('tmpnn_graph_associate', wg.g.c, wg.index().c, 0, None, L.stream())


def _compact(wg, keep, states, scores):
    """Deletes the rows with keep == 0; returns (WindowGraph', states', scores')."""
    dev = wg.device
    n = wg.n
    out = WindowGraph(0, max(1, n), dev, with_labels=True)
    ldh = int(states.shape[1]) if states is not None else 0
    pad = (-ldh) % 4
    h_src = h_dst = None
    if states is not None:
        h_src = states.detach().to(device=dev, dtype=torch.float32)
        if pad:
            h_src = torch.nn.functional.pad(h_src, (0, pad))
        h_src = h_src.contiguous()
        h_dst = torch.empty_like(h_src)
    new_of_old = torch.empty(wg.g.cap_rows, dtype=_I32, device=dev)
    scratch = torch.empty(int(L.lib().tmpnn_graph_compact_scratch_ints(1, wg.g.cap_rows)), dtype=_I32, device=dev)
    L.call('tmpnn_graph_compact', wg.g.c, out.g.c, L.ptr(keep), L.ptr(h_src), None, None, L.ptr(h_dst), ldh + pad,
           L.ptr(new_of_old), L.ptr(scratch), L.stream())
    out.n = int(out.g.n_rows.item())
    new_states = None if states is None else h_dst[:out.n, :ldh]
    new_scores = None
    if scores is not None:
        keep_idx = torch.nonzero(new_of_old[:n] >= 0)[:, 0]
        new_scores = scores.to(dev)[keep_idx]
    return out, new_states, new_scores


def prune_graph(states, node_adj, labels, scores, y_pred, t_st, t_ed, threshold=0.5, cuda=True):
    """Reference ``utils/graph.py:337-389``: drop low-probability edge rows inside [t_st, t_ed].

    Returns ``(y_pred', states', node_adj', labels', scores')``."""
    assert (t_st <= t_ed), "t_st must be lesser than or equal to t_ed!"
    dev = _cuda_device(y_pred, node_adj, scores)
    wg = WindowGraph.from_tensors(y_pred, node_adj, labels, scores)
    # the reference keeps y_pred[:, 2] as is (no re-association here)
    keep = torch.empty(wg.g.cap_rows, dtype=torch.uint8, device=dev)
    scratch = torch.empty(2, dtype=_I32, device=dev)
    L.call('tmpnn_graph_prune_mask', wg.g.c, wg.index().c, int(t_st), int(t_ed), float(threshold), L.ptr(keep),
           L.ptr(scratch), L.stream())
    out, new_states, new_scores = _compact(wg, keep, states, scores)
    out_labels = None if labels is None else out.labels()
    return out.y_pred(), new_states, out.adjacency(False), out_labels, new_scores


def decode_tracks(states, node_adj, labels, scores, y_pred, y_out, t_upto, ret_win_size, use_hungraian=True, cuda=True):
    """Reference ``utils/graph.py:392-539``: re-associate, walk chains to assign track ids into
    ``y_out`` (host ndarray, mutated in place), delete everything before ``t_upto``.

    Returns ``(y_pred', y_out, states', node_adj', labels', scores')``."""
    dev = _cuda_device(y_pred, node_adj, scores, states)
    wg = WindowGraph.from_tensors(y_pred, node_adj, labels, scores)
    _associate(wg, scores, use_hungraian, 'test')
    nd_seq = int(y_out.shape[0])
    track = torch.from_numpy(np.ascontiguousarray(y_out[:, 1]).astype(np.int32)).to(dev)
    next_id = torch.tensor([int(np.amax(y_out[:, 1])) + 1], dtype=_I32, device=dev)
    det_ptr = torch.tensor([0, nd_seq], dtype=_I32, device=dev)
    fr = L.Frames(0, 0, None, None, L.ptr(det_ptr), None)
    keep = torch.empty(wg.g.cap_rows, dtype=torch.uint8, device=dev)
    max_dets = max(1, wg.n)   # bound of the window's detection rows known without a read-back: its row count
    scratch = torch.empty(int(L.lib().tmpnn_graph_decode_scratch_ints(1, max_dets)), dtype=_I32, device=dev)
    import ctypes as C
    L.call('tmpnn_graph_decode', wg.g.c, wg.index().c, C.byref(fr), L.ptr(track), L.ptr(next_id), None, int(t_upto),
           None, int(ret_win_size), L.ptr(keep), max_dets, L.ptr(scratch), L.stream())
    wg.g.check_status()
    y_out[:, 1] = track.cpu().numpy().astype(y_out.dtype)
    out, new_states, new_scores = _compact(wg, keep, states, scores)
    new_labels = out.labels()
    return (_place(out.y_pred(), cuda, dev), y_out, _place(new_states, cuda, dev), _place_adj(out.adjacency(False), cuda),
            _place(new_labels, cuda, dev), _place(new_scores, cuda, dev))


def hungarian(node_adj, scores, y_pred, t, threshold=0.5):
    """Reference ``utils/graph.py:33-93``: optimal assignment for the detections of timestep t only;
    returns ``y_pred`` with the new associations written into column 2 (earlier ones are kept)."""
    wg = WindowGraph.from_tensors(torch.as_tensor(y_pred), node_adj, None, None)
    _hungarian(wg, torch.as_tensor(scores), only_t=True, t=int(t), threshold=threshold)
    out = wg.y_pred()
    if isinstance(y_pred, np.ndarray):
        return out.cpu().numpy().astype(y_pred.dtype)
    return out.to(device=y_pred.device, dtype=y_pred.dtype)
