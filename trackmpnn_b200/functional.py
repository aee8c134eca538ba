"""Host-side glue between the reference-shaped modules and the CUDA library.

Nothing here computes on the host or through PyTorch ops beyond allocating / slicing
tensors: the input transform, the aggregation, the GRU step and the heads all run in
``libtmpnn_sm100a.so`` (``include/tmpnn.h``).
"""
import torch

from . import _lib as L
from .device_graph import window_graph_of

H = L.HIDDEN


class _PackCache:
    """Packed GRU cells per feature group, rebuilt when a parameter changes in place."""

    def __init__(self):
        self.key = None
        self.packs = None


def packed_cells(model):
    cache = model.__dict__.setdefault('_tmpnn_pack_cache', _PackCache())
    params = []
    for gru in model.factor_grus:
        for cell in (gru.edge_gru, gru.node_gru):
            params += [cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh]
    params += [model.output_transform_edge.weight, model.output_transform_edge.bias,
               model.output_transform_node.weight, model.output_transform_node.bias]
    key = tuple((p.data_ptr(), p._version) for p in params)
    if cache.key != key:
        packs = []
        for g, gru in enumerate(model.factor_grus):
            pair = []
            for cell, head in ((gru.edge_gru, model.output_transform_edge), (gru.node_gru, model.output_transform_node)):
                kx = int(cell.weight_ih.shape[1])
                out = torch.empty(int(L.lib().tmpnn_gru_pack_floats(kx)), dtype=torch.float32, device=cell.weight_ih.device)
                hw = head.weight.detach()[0, g * H:(g + 1) * H]
                L.call('tmpnn_pack_gru', L.ptr(cell.weight_ih.detach()), L.ptr(cell.weight_hh.detach()),
                       L.ptr(cell.bias_ih.detach()), L.ptr(cell.bias_hh.detach()), L.ptr(hw), L.ptr(head.bias.detach()),
                       kx, L.ptr(out), L.stream())
                pair.append(out)
            packs.append(tuple(pair))
        cache.key, cache.packs = key, packs
    return cache.packs


def packed_cells_tc(model):
    """fp16 hi/lo UMMA weight images of the edge cells (tensor-core path, msg_type 'diff' only)."""
    cache = model.__dict__.setdefault('_tmpnn_pack_cache_tc', _PackCache())
    params = []
    for gru in model.factor_grus:
        cell = gru.edge_gru
        params += [cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh]
    params += [model.output_transform_edge.weight, model.output_transform_edge.bias]
    key = tuple((p.data_ptr(), p._version) for p in params)
    if cache.key != key:
        packs = []
        nbytes = int(L.lib().tmpnn_gru_tc_pack_bytes())
        for g, gru in enumerate(model.factor_grus):
            cell, head = gru.edge_gru, model.output_transform_edge
            if gru.msg_type != 'diff':
                packs.append(None)
                continue
            out = torch.empty(nbytes, dtype=torch.uint8, device=cell.weight_ih.device)
            hw = head.weight.detach()[0, g * H:(g + 1) * H]
            L.call('tmpnn_pack_gru_tc', L.ptr(cell.weight_ih.detach()), L.ptr(cell.weight_hh.detach()),
                   L.ptr(cell.bias_ih.detach()), L.ptr(cell.bias_hh.detach()), L.ptr(hw), L.ptr(head.bias.detach()),
                   L.ptr(out), L.stream())
            packs.append(out)
        cache.key, cache.packs = key, packs
    return cache.packs


TENSOR_MIN_ROWS = 8192  # below this a window graph cannot fill the 148 x 128-row tiles; the FMA path wins


def use_tensor_path(model, n_rows):
    mode = getattr(model, 'use_tensor_cores', 'auto')
    if any(g.msg_type != 'diff' for g in model.factor_grus):
        return False
    if mode == 'auto':
        return n_rows >= TENSOR_MIN_ROWS
    return bool(mode)


def mp_step(model, g_c, ix_c, h_in, h_out, ldh, agg, tensor):
    """One message-passing step over all feature groups (K1-det, edge rows, detection rows)."""
    G = len(model.feature_idx)
    packs = packed_cells(model)
    tc = packed_cells_tc(model) if tensor else None
    st = L.stream()
    for g in range(G):
        concat = int(model.factor_grus[g].msg_type == 'concat')
        L.call('tmpnn_aggregate_dets', g_c, ix_c, L.ptr(h_in), ldh, g * H, L.ptr(agg), st)
        if tensor:
            L.call('tmpnn_mp_edge_fwd_tc', g_c, ix_c, L.ptr(h_in), L.ptr(h_out), ldh, g, G, L.ptr(tc[g]), st)
        else:
            L.call('tmpnn_mp_edge_fwd', g_c, ix_c, L.ptr(h_in), L.ptr(h_out), ldh, g, G, concat, L.ptr(packs[g][0]), st)
        L.call('tmpnn_mp_det_fwd', g_c, ix_c, L.ptr(h_in), L.ptr(h_out), ldh, g, G, L.ptr(packs[g][1]), L.ptr(agg), st)


def input_transform_rows(model, g, x, x_idx, n, n_edge_rows, h, ldh, out_rows, n_dev=None):
    """K0 for feature group g on ``n`` detection feature rows ``x[x_idx]`` -> ``h[out_rows, 64g:64g+64]``."""
    seq = model.input_transforms[g]
    lin1, bn, lin2 = seq[0], seq[1], seq[3]
    cols = model.feature_idx[g]
    dev = h.device
    cap = n if n_dev is None else int(x_idx.numel())
    a = torch.empty((max(1, cap), H), dtype=torch.float32, device=dev)
    L.call('tmpnn_input_linear1', L.ptr(x), int(x.shape[1]), int(cols[0]), len(cols), L.ptr(x_idx),
           L.ptr(lin1.weight.detach()), L.ptr(lin1.bias.detach()), L.ptr(a), L.ptr(n_dev), int(n), L.stream())
    if bn.training:
        if n_dev is not None:
            raise L.TmpnnError('train-mode BatchNorm needs host-known row counts')
        if n + n_edge_rows <= 1:
            raise ValueError('Expected more than 1 value per channel when training, got input size '
                             f'torch.Size([{n + n_edge_rows}, {H}])')
        assert bn.momentum == 0.1 and bn.eps == 1e-5, 'kernels assume BatchNorm1d defaults'
        stats = torch.empty((2, H), dtype=torch.float32, device=dev)
        L.call('tmpnn_input_bn_stats', L.ptr(a), int(n), int(n_edge_rows), L.ptr(lin1.bias.detach()), L.ptr(stats),
               L.ptr(bn.running_mean), L.ptr(bn.running_var), L.stream())
        bn.num_batches_tracked += 1
        mean, var = stats[0], stats[1]
    else:
        mean, var = bn.running_mean, bn.running_var
    L.call('tmpnn_input_bn_relu_linear2', L.ptr(a), L.ptr(mean), L.ptr(var), L.ptr(bn.weight.detach()),
           L.ptr(bn.bias.detach()), L.ptr(lin2.weight.detach()), L.ptr(lin2.bias.detach()), L.ptr(h), int(ldh), g * H,
           L.ptr(out_rows), L.ptr(n_dev), int(n), L.stream())
    return a


def track_mpnn_forward(model, x, h_in, node_adj, edge_adj):
    """``TrackMPNN.forward`` (reference ``models/track_mpnn.py:54-75``) on the CUDA library."""
    wg = window_graph_of(node_adj)
    dev = wg.device
    n_tot = wg.n
    G = len(model.feature_idx)
    ldh = G * H
    n_new = int(x.size()[0])
    n_old = n_tot - n_new
    if (0 if h_in is None else int(h_in.shape[0])) != n_old:
        raise ValueError(f'h_in has {0 if h_in is None else int(h_in.shape[0])} rows, graph has {n_tot} rows of '
                         f'which {n_new} are new')
    h_cur = torch.zeros((n_tot, ldh), dtype=torch.float32, device=dev)
    if n_old:
        h_cur[:n_old].copy_(h_in.detach())
    if n_new > 0:
        new_det = torch.nonzero(wg.g.ts[n_old:n_tot] >= 0)[:, 0].to(torch.int32)
        nd = int(new_det.numel())
        out_rows = (new_det + n_old).contiguous()
        xd = x.detach().to(device=dev, dtype=torch.float32).contiguous()
        for g in range(G):
            input_transform_rows(model, g, xd, new_det, nd, n_new - nd, h_cur, ldh, out_rows)
    ix = wg.index()
    h_out = torch.empty_like(h_cur)
    agg = torch.empty((ix.cap_dets, H), dtype=torch.float32, device=dev)
    tensor = use_tensor_path(model, n_tot)
    mp_step(model, wg.g.c, ix.c, h_cur, h_out, ldh, agg, tensor)
    if tensor:
        wg.g.check_status()
    logits = wg.g.logit[:n_tot].clone().unsqueeze(1)
    scores = wg.g.score[:n_tot].clone().unsqueeze(1)
    return scores, logits, h_out, tuple(None for _ in range(G))


def mp_step_single_group(gru, h, node_adj):
    """``FactorGraphGRU.forward`` (reference ``models/layers.py:84-116``) for one 64-wide group.
    The heads are fused in the kernel; with no head attached here a zero head is packed."""
    wg = window_graph_of(node_adj)
    dev = wg.device
    n = wg.n
    zero_w = torch.zeros(H, dtype=torch.float32, device=dev)
    zero_b = torch.zeros(1, dtype=torch.float32, device=dev)
    packs = []
    for cell in (gru.edge_gru, gru.node_gru):
        kx = int(cell.weight_ih.shape[1])
        out = torch.empty(int(L.lib().tmpnn_gru_pack_floats(kx)), dtype=torch.float32, device=dev)
        L.call('tmpnn_pack_gru', L.ptr(cell.weight_ih.detach()), L.ptr(cell.weight_hh.detach()),
               L.ptr(cell.bias_ih.detach()), L.ptr(cell.bias_hh.detach()), L.ptr(zero_w), L.ptr(zero_b), kx,
               L.ptr(out), L.stream())
        packs.append(out)
    ix = wg.index()
    h_in = h.detach().to(device=dev, dtype=torch.float32).contiguous()
    h_out = torch.empty_like(h_in)
    agg = torch.empty((ix.cap_dets, H), dtype=torch.float32, device=dev)
    L.call('tmpnn_mp_step_fwd', wg.g.c, ix.c, L.ptr(h_in), L.ptr(h_out), H, 0, 1, int(gru.msg_type == 'concat'),
           L.ptr(packs[0]), L.ptr(packs[1]), L.ptr(agg), L.stream())
    return h_out
