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


def packed_cells_tc(model, node=False):
    """fp16 hi/lo UMMA weight images of the edge cells (tensor-core path); node=True: of the node cells with the detection
    head (tmpnn_mp_det_fwd_tc)."""
    cache = model.__dict__.setdefault('_tmpnn_pack_cache_tc_node' if node else '_tmpnn_pack_cache_tc', _PackCache())
    head = model.output_transform_node if node else model.output_transform_edge
    params = []
    for gru in model.factor_grus:
        cell = gru.node_gru if node else gru.edge_gru
        params += [cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh]
    params += [head.weight, head.bias]
    key = tuple((p.data_ptr(), p._version) for p in params)
    if cache.key != key:
        packs = []
        nbytes = int(L.lib().tmpnn_gru_tc_pack_bytes())
        for g, gru in enumerate(model.factor_grus):
            cell = gru.node_gru if node else gru.edge_gru
            out = torch.empty(nbytes, dtype=torch.uint8, device=cell.weight_ih.device)
            hw = head.weight.detach()[0, g * H:(g + 1) * H]
            # 'concat': the image holds the far-endpoint half W_ih[:, 64:128] (only tmpnn_mp_edge_fwd_tc_pre takes it);
            # the node cell's input is always the 64-wide aggregate
            L.call('tmpnn_pack_gru_tc', L.ptr(cell.weight_ih.detach()), L.ptr(cell.weight_hh.detach()),
                   L.ptr(cell.bias_ih.detach()), L.ptr(cell.bias_hh.detach()), L.ptr(hw), L.ptr(head.bias.detach()),
                   int(gru.msg_type == 'concat' and not node), L.ptr(out), L.stream())
            packs.append(out)
        cache.key, cache.packs = key, packs
    return cache.packs


TENSOR_MIN_ROWS = 8192  # below this a window graph cannot fill the 148 x 128-row tiles; the FMA path wins


def use_tensor_path(model, n_rows):
    mode = getattr(model, 'use_tensor_cores', 'auto')
    if mode == 'auto':
        return n_rows >= TENSOR_MIN_ROWS
    return bool(mode)


class SparseAttention:
    """Attention of one head, one weight per incidence entry (detection, incident edge).  ``to_dense()``
    gives the reference's [N, N] matrix (``models/layers.py:35-37``): softmax rows for the detections and
    the uniform 1/N rows its masked softmax yields for rows without incident edges."""

    def __init__(self, n, n_dets, det_rows, seg_ptr, inc, alpha):
        self.n, self.n_dets, self.det_rows, self.seg_ptr, self.inc, self.alpha = n, n_dets, det_rows, seg_ptr, inc, alpha

    def to_dense(self):
        nd = int(self.n_dets.item())
        A = torch.full((self.n, self.n), 1.0 / self.n, dtype=torch.float32, device=self.alpha.device)
        seg = self.seg_ptr[:2 * nd + 1].long()
        lens = seg[2::2] - seg[0:-1:2]           # incident edges per detection (past + future)
        rows = self.det_rows[:nd].long()
        A[rows[lens > 0]] = 0
        tot = int(seg[-1].item())
        owner = torch.repeat_interleave(rows, lens)
        A[owner, self.inc[:tot].long()] = self.alpha[:tot]
        return A


def aggregate_for_dets(gru, graph, index, h_in, ldh, col, agg, scratch=None, keep_attention=False, blocks=None):
    """Input of the node GRU for the detection rows: the plain signed sum, or -- with attention heads --
    the attention-weighted one (reference ``models/layers.py:101-112``).  Returns a list of per-head
    ``alpha`` tensors when ``keep_attention`` (else None)."""
    st = L.stream()
    if gru.gat is None:
        if blocks is not None:
            # structured index (TrackEngine): one pass over the dense edge blocks, every association row read once
            L.call('tmpnn_aggregate_dets_blocks', graph.c, index.c, L.ptr(index._scratch2), L.ptr(h_in), ldh, col, L.ptr(agg),
                   L.ptr(blocks['scratch']), blocks['cap_runs'], blocks['cap_cpart'], st)
        else:
            L.call('tmpnn_aggregate_dets', graph.c, index.c, L.ptr(h_in), ldh, col, L.ptr(agg), st)
        return None
    if gru.training:
        # train mode: dropout on the attention (same kernels as the autograd path; the saved tensors are dropped)
        saved = gat_forward_train(gru, graph, index, h_in, ldh, col, agg)
        return [attention_after_dropout(sv) for sv in saved] if keep_attention else None
    dev = h_in.device
    if scratch is None:
        scratch = {}
    hatt = scratch.get('hatt')
    if hatt is None or hatt.shape[0] < index.cap_dets:
        hatt = scratch['hatt'] = torch.empty((index.cap_dets, H), dtype=torch.float32, device=dev)
    esc = scratch.get('escore')
    n_all = graph.num_seqs * graph.cap_rows
    if esc is None or esc.numel() < n_all:
        esc = scratch['escore'] = torch.empty(n_all, dtype=torch.float32, device=dev)
    alphas = []
    nh = len(gru.gat)
    for k, head in enumerate(gru.gat):
        alpha = torch.empty(index.cap_inc, dtype=torch.float32, device=dev) if keep_attention else None
        L.call('tmpnn_gat_aggregate_dets', graph.c, index.c, L.ptr(h_in), ldh, col, L.ptr(head.W_att.detach().contiguous()),
               L.ptr(head.a.detach().contiguous()), k, nh, L.ptr(hatt), L.ptr(esc), L.ptr(agg), L.ptr(alpha), st)
        alphas.append(alpha)
    return alphas if keep_attention else None


ATTENTION_DROPOUT_P = 0.5  # nn.Dropout(p=0.5) of the reference's GraphAttentionLayer (models/layers.py:24)


def attention_keep_mask(gru, head, index, device):
    """One keep byte per incidence entry (``index.inc`` order) for ``head`` of ``gru``.  Drawn from torch's CUDA
    generator; a test (or a caller that wants reproducible masks) can set
    ``gru.attention_keep_fn = fn(head, index, device) -> uint8 [index.cap_inc]``."""
    n = index.cap_inc
    fn = getattr(gru, 'attention_keep_fn', None)
    if fn is not None:
        keep = fn(head, index, device)
        if keep.dtype != torch.uint8 or keep.numel() < n or keep.device != device:
            raise ValueError('attention_keep_fn must return a uint8 tensor with one entry per incidence entry on the device')
        return keep.contiguous()
    return (torch.rand(n, device=device) >= ATTENTION_DROPOUT_P).to(torch.uint8)


def gat_forward_train(gru, graph, index, h_in, ldh, col, agg):
    """The attention-weighted aggregate of every head with dropout when ``gru.training`` (reference
    ``models/layers.py:26-43, 105-112``); returns per head what ``tmpnn_gat_bwd`` needs."""
    dev = h_in.device
    st = L.stream()
    nh = len(gru.gat)
    training = bool(gru.training)
    scale = 1.0 / (1.0 - ATTENTION_DROPOUT_P) if training else 1.0
    saved = []
    for k, head in enumerate(gru.gat):
        f32 = dict(dtype=torch.float32, device=dev)
        sv = dict(w=head.W_att.detach().contiguous(), a=head.a.detach().contiguous(),
                  keep=attention_keep_mask(gru, k, index, dev) if training else None, scale=scale,
                  hatt=torch.empty((index.cap_dets, H), **f32), esc=torch.empty(graph.cap_rows, **f32),
                  alpha=torch.empty(index.cap_inc, **f32), att_edge=torch.empty(2 * graph.cap_rows, **f32))
        L.call('tmpnn_gat_aggregate_dets_train', graph.c, index.c, L.ptr(h_in), ldh, col, L.ptr(sv['w']), L.ptr(sv['a']), k, nh,
               L.ptr(sv['keep']), scale, L.ptr(sv['hatt']), L.ptr(sv['esc']), L.ptr(agg), L.ptr(sv['alpha']),
               L.ptr(sv['att_edge']), st)
        saved.append(sv)
    return saved


def attention_after_dropout(sv):
    """What the reference returns as ``attention`` in train mode: the softmax output after dropout, per incidence entry."""
    return sv['alpha'] if sv['keep'] is None else sv['alpha'] * sv['keep'] * sv['scale']


def edge_step_tc(model, graph, index, h_in, h_out, ldh, g, G, image, scratch, kernel='auto'):
    """Edge rows of feature group g on the tensor cores.  kernel 'pre' (= 'auto'): endpoints prepared once per
    detection row, re-staged kernel (tmpnn_mp_edge_fwd_tc_pre; both msg_types); 'gather': endpoints gathered,
    subtracted and split per association row (tmpnn_mp_edge_fwd_tc, the first tensor-core kernel, 'diff' only)."""
    gru = model.factor_grus[g]
    concat = int(gru.msg_type == 'concat')
    st = L.stream()
    if kernel == 'auto':
        kernel = 'pre'
    if kernel not in ('pre', 'gather'):
        raise ValueError(f'unknown tensor-core kernel {kernel!r}')
    if kernel == 'gather':
        if concat:
            raise L.TmpnnError("tmpnn_mp_edge_fwd_tc handles msg_type 'diff' only")
        L.call('tmpnn_mp_edge_fwd_tc', graph.c, index.c, L.ptr(h_in), L.ptr(h_out), ldh, g, G, L.ptr(image), st)
        return
    n_all = graph.num_seqs * graph.cap_rows
    img = scratch.get('det_img')
    if img is None or img.shape != (n_all, ldh):
        img = scratch['det_img'] = torch.empty((n_all, ldh), dtype=torch.float32, device=h_in.device)
    dp = scratch.get('det_p')
    if dp is None or dp.shape[0] < index.cap_dets:
        dp = scratch['det_p'] = torch.empty((index.cap_dets, 3 * H), dtype=torch.float32, device=h_in.device)
    nb = int(L.lib().tmpnn_tc_tile_table_bytes(graph.num_seqs, graph.cap_rows))
    tab = scratch.get('tile_tab')
    if tab is None or tab.numel() * 4 < nb:
        tab = scratch['tile_tab'] = torch.empty(((nb + 15) // 16, 4), dtype=torch.int32, device=h_in.device)
    L.call('tmpnn_mp_edge_fwd_tc_pre', graph.c, index.c, L.ptr(h_in), L.ptr(h_out), ldh, g, G, concat, L.ptr(image),
           L.ptr(img), L.ptr(dp), L.ptr(tab), st)


def det_step_tc(graph, index, h_in, h_out, ldh, g, G, node_image, agg, scratch):
    """Detection rows of feature group g on the tensor cores (tmpnn_mp_det_fwd_tc): after edge_step_tc(kernel='pre') of the
    same group, whose det_img / det_p scratch it reuses; needs the structured index (segment tables in index._scratch2)."""
    nb = int(L.lib().tmpnn_tc_det_tile_table_bytes(graph.num_seqs, index.cap_dets))
    tab = scratch.get('det_tile_tab')
    if tab is None or tab.numel() * 4 < nb:
        tab = scratch['det_tile_tab'] = torch.empty(((nb + 15) // 16, 4), dtype=torch.int32, device=h_in.device)
    L.call('tmpnn_mp_det_fwd_tc', graph.c, index.c, L.ptr(index._scratch2), L.ptr(h_in), L.ptr(h_out), ldh, g, G,
           L.ptr(node_image), L.ptr(agg), L.ptr(scratch['det_img']), L.ptr(scratch['det_p']), L.ptr(tab), L.stream())


def rerun_edges_if_out_of_range(model, graph, index, h_in, h_out, ldh, packs, aggs=None):
    """Behind the tensor-core edge step: if an activation left the range of the fp16 split (|h| > 6e4, e.g. a trained
    input transform with large outputs; TMPNN_FLAG_TC_RANGE), the association rows of every feature group are re-run by
    the fp32 FMA kernel.  The test happens on the device (the launches exit at once otherwise), so this sits inside the
    engine's CUDA graph; the flag then becomes the note NOTE_TC_RANGE_RERUN."""
    st = L.stream()
    G = len(model.feature_idx)
    for g in range(G):
        concat = int(model.factor_grus[g].msg_type == 'concat')
        L.call('tmpnn_mp_edge_fwd_on_flag', graph.c, index.c, L.ptr(h_in), L.ptr(h_out), ldh, g, G, concat, L.ptr(packs[g][0]),
               L.FLAG_TC_RANGE, st)
    if aggs is not None:   # the detection rows went through the tensor cores too (aggs[g]: group g's aggregates, kept)
        for g in range(G):
            L.call('tmpnn_mp_det_fwd_on_flag', graph.c, index.c, L.ptr(h_in), L.ptr(h_out), ldh, g, G, L.ptr(packs[g][1]),
                   L.ptr(aggs[g]), L.FLAG_TC_RANGE, st)
    L.call('tmpnn_status_ack', graph.c, L.FLAG_TC_RANGE, L.NOTE_TC_RANGE_RERUN, st)


def mp_step(model, graph, index, h_in, h_out, ldh, agg, tensor, keep_attention=False):
    """One message-passing step over all feature groups (K1-det, edge rows, detection rows).
    Returns the per-group attention (list of per-head alpha tensors, or None)."""
    G = len(model.feature_idx)
    g_c, ix_c = graph.c, index.c
    packs = packed_cells(model)
    tc = packed_cells_tc(model) if tensor else None
    st = L.stream()
    att = []
    for g in range(G):
        concat = int(model.factor_grus[g].msg_type == 'concat')
        att.append(aggregate_for_dets(model.factor_grus[g], graph, index, h_in, ldh, g * H, agg, None, keep_attention))
        if tensor:
            scratch = model.__dict__.setdefault('_tmpnn_tc_scratch', {})
            edge_step_tc(model, graph, index, h_in, h_out, ldh, g, G, tc[g], scratch,
                         kernel=getattr(model, 'tensor_core_kernel', 'auto'))
        else:
            L.call('tmpnn_mp_edge_fwd', g_c, ix_c, L.ptr(h_in), L.ptr(h_out), ldh, g, G, concat, L.ptr(packs[g][0]), st)
        L.call('tmpnn_mp_det_fwd', g_c, ix_c, L.ptr(h_in), L.ptr(h_out), ldh, g, G, L.ptr(packs[g][1]), L.ptr(agg), st)
    if tensor:
        rerun_edges_if_out_of_range(model, graph, index, h_in, h_out, ldh, packs)
    return att


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
    return a, mean, var


def track_mpnn_forward(model, x, h_in, node_adj, edge_adj):
    """``TrackMPNN.forward`` (reference ``models/track_mpnn.py:54-75``) on the CUDA library."""
    # The autograd Function (FMA kernel, stored gates) runs only where a backward pass can follow: grad mode on AND
    # (train mode with trainable parameters, or a state that already carries a graph).  The reference's infer.py never
    # enters torch.no_grad() but never calls backward either: in eval mode it gets the inference path (tensor-core
    # kernel on large graphs, nothing saved).  Gradients in eval mode: pass an h_in that requires grad, or call .train().
    if torch.is_grad_enabled() and ((model.training and any(p.requires_grad for p in model.parameters()))
                                    or (h_in is not None and h_in.requires_grad)):
        scores, logits, h_out, *alphas = _MPStepFn.apply(model, node_adj, x, h_in, *_param_list(model))
        # attention slots: per group None or one SparseAttention per head (after dropout, as the reference returns it)
        attention, wg = [], None
        for gru in model.factor_grus:
            if gru.gat is None:
                attention.append(None)
                continue
            if wg is None:
                wg = window_graph_of(node_adj)
                ix = wg.index()
            attention.append([SparseAttention(wg.n, ix.n_dets, ix.det_rows, ix.seg_ptr, ix.inc, alphas.pop(0)) for _ in gru.gat])
        return scores, logits, h_out, tuple(attention)
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
    att = mp_step(model, wg.g, ix, h_cur, h_out, ldh, agg, tensor, keep_attention=True)
    if tensor:
        wg.g.check_status()
    logits = wg.g.logit[:n_tot].clone().unsqueeze(1)
    scores = wg.g.score[:n_tot].clone().unsqueeze(1)
    attention = tuple(None if a is None else [SparseAttention(n_tot, ix.n_dets, ix.det_rows, ix.seg_ptr, ix.inc, al) for al in a]
                      for a in att)
    return scores, logits, h_out, attention


def mp_step_single_group(gru, h, node_adj):
    """``FactorGraphGRU.forward`` (reference ``models/layers.py:84-116``) for one 64-wide group ->
    ``(h', attention)``.  The heads are fused in the kernel; with no head attached here a zero head is packed."""
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
    st = L.stream()
    alphas = aggregate_for_dets(gru, wg.g, ix, h_in, H, 0, agg, None, keep_attention=True)
    L.call('tmpnn_mp_edge_fwd', wg.g.c, ix.c, L.ptr(h_in), L.ptr(h_out), H, 0, 1, int(gru.msg_type == 'concat'),
           L.ptr(packs[0]), st)
    L.call('tmpnn_mp_det_fwd', wg.g.c, ix.c, L.ptr(h_in), L.ptr(h_out), H, 0, 1, L.ptr(packs[1]), L.ptr(agg), st)
    att = None if alphas is None else [SparseAttention(n, ix.n_dets, ix.det_rows, ix.seg_ptr, ix.inc, a) for a in alphas]
    return h_out, att


# ---------------------------------------------------------------------------------------------
# training: the step as a torch.autograd.Function (reference: loss.backward() in train.py:65-134)
# ---------------------------------------------------------------------------------------------
_PER_GROUP = 14  # lin1.w, lin1.b, bn.w, bn.b, lin2.w, lin2.b, edge_gru x4, node_gru x4


def _param_list(model):
    ps = []
    for g in range(len(model.feature_idx)):
        seq, gru = model.input_transforms[g], model.factor_grus[g]
        ps += [seq[0].weight, seq[0].bias, seq[1].weight, seq[1].bias, seq[3].weight, seq[3].bias]
        for cell in (gru.edge_gru, gru.node_gru):
            ps += [cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh]
    for gru in model.factor_grus:   # attention heads sit between the per-group blocks and the four head tensors
        for head in (gru.gat or ()):
            ps += [head.W_att, head.a]
    ps += [model.output_transform_node.weight, model.output_transform_node.bias,
           model.output_transform_edge.weight, model.output_transform_edge.bias]
    return ps


class NewRowGroups:
    """New detection rows of a block-diagonal batch of window graphs (trackmpnn_b200/train_engine.py): ``xd`` holds
    the feature rows, every group ``(x_idx, out_rows, n_dets, n_edge_rows)`` is one chunk's share -- one BatchNorm
    batch, as in the reference, which feeds one chunk per forward."""

    def __init__(self, xd, groups, x_idx_all=None, carry=None):
        self.xd, self.groups, self.x_idx_all = xd, groups, x_idx_all   # x_idx_all: the groups' x_idx concatenated
        # carry = (seg_src, seg_dst, seg_len int32 device tensors, number of segments, longest segment): h_in is then the
        # PREVIOUS step's state in the previous block-diagonal layout; segment s (one chunk) moves seg_len[s] rows from
        # row seg_src[s] there to row seg_dst[s] of this step's layout, everything else starts at zero
        self.carry = carry


def _input_rows_forward(model, xd, x_idx, out_rows, nd, n_edge_rows, h_cur, ldh):
    """K0 of every feature group on one BatchNorm batch of new rows (nd detections + n_edge_rows all-zero edge
    rows); returns what tmpnn_input_bwd needs."""
    per_group = []
    for g in range(len(model.feature_idx)):
        training = model.input_transforms[g][1].training
        a, mean, var = input_transform_rows(model, g, xd, x_idx, nd, n_edge_rows, h_cur, ldh, out_rows)
        per_group.append((a, mean.clone(), var.clone(), training))
    return (xd, x_idx, out_rows, nd, n_edge_rows, per_group)


def _input_groups_forward(model, x, h_cur, ldh):
    """K0 for the row groups of a NewRowGroups (one BatchNorm batch each).  Train-mode BatchNorm: Linear1 over all
    rows at once, then statistics / running averages / BatchNorm-ReLU-Linear2 of all groups in three launches
    (tmpnn_input_bn_groups_fwd); otherwise group by group.  Returns the per-group tuples tmpnn_input_bwd needs."""
    groups = [t for t in x.groups if t[2] > 0]
    G = len(model.feature_idx)
    if not groups:
        return []
    if not all(model.input_transforms[g][1].training for g in range(G)) or len(groups) == 1:
        return [_input_rows_forward(model, x.xd, xi, orow, nd, ne, h_cur, ldh) for xi, orow, nd, ne in groups]
    dev = h_cur.device
    tot = sum(t[2] for t in groups)
    x_all = torch.cat([t[0] for t in groups]) if x.x_idx_all is None else x.x_idx_all
    st = L.stream()
    per = [[] for _ in groups]
    for g in range(G):
        seq = model.input_transforms[g]
        lin1, bn, lin2 = seq[0], seq[1], seq[3]
        cols = model.feature_idx[g]
        assert bn.momentum == 0.1 and bn.eps == 1e-5, 'kernels assume BatchNorm1d defaults'
        a = torch.empty((tot, H), dtype=torch.float32, device=dev)
        L.call('tmpnn_input_linear1', L.ptr(x.xd), int(x.xd.shape[1]), int(cols[0]), len(cols), L.ptr(x_all),
               L.ptr(lin1.weight.detach()), L.ptr(lin1.bias.detach()), L.ptr(a), None, int(tot), st)
        stats = torch.empty((len(groups), 2, H), dtype=torch.float32, device=dev)
        desc = (L.InputGroup * len(groups))()
        off = 0
        for k, (xi, orow, nd, ne) in enumerate(groups):
            if nd + ne <= 1:
                raise ValueError('Expected more than 1 value per channel when training, got input size '
                                 f'torch.Size([{nd + ne}, {H}])')
            desc[k] = L.InputGroup(a.data_ptr() + 4 * H * off, stats[k, 0].data_ptr(), stats[k, 1].data_ptr(), xi.data_ptr(),
                                   orow.data_ptr(), None, nd, ne)
            per[k].append((a[off:off + nd], stats[k, 0], stats[k, 1], True))
            off += nd
        L.call('tmpnn_input_bn_groups_fwd', desc, len(groups), L.ptr(lin1.bias.detach()), L.ptr(bn.weight.detach()),
               L.ptr(bn.bias.detach()), L.ptr(lin2.weight.detach()), L.ptr(lin2.bias.detach()), L.ptr(bn.running_mean),
               L.ptr(bn.running_var), L.ptr(h_cur), int(ldh), g * H, st)
        bn.num_batches_tracked += len(groups)
    return [(x.xd, xi, orow, nd, ne, per[k]) for k, (xi, orow, nd, ne) in enumerate(groups)]


def _gate_bwd_partials(model, dev):
    """Scratch of tmpnn_gate_bwd: per-CTA bias / head-weight sums, added in CTA order (reproducible) instead of atomically."""
    sc = model.__dict__.setdefault('_tmpnn_bwd_tc', {})
    t = sc.get('gate_partials')
    if t is None or t.device != dev:
        t = sc['gate_partials'] = torch.empty(int(L.lib().tmpnn_gate_bwd_partial_floats()), dtype=torch.float32, device=dev)
    return t


def _bwd_tc_partials(model, dev):
    """Scratch of the tensor-core backward contraction: one weight-gradient partial per SM."""
    sc = model.__dict__.setdefault('_tmpnn_bwd_tc', {})
    t = sc.get('partials')
    if t is None or t.device != dev:
        t = sc['partials'] = torch.empty(int(L.lib().tmpnn_bwd_tc_partial_floats()), dtype=torch.float32, device=dev)
    return t


def _bwd_tc_image(model, w, col0):
    """bf16 hi / lo UMMA image of columns [col0, col0 + 64) of a GRU weight, re-packed when the weight changes."""
    sc = model.__dict__.setdefault('_tmpnn_bwd_tc', {})
    key = ('img', w.data_ptr(), col0)
    ent = sc.get(key)
    if ent is None or ent[0] != w._version:
        img = ent[1] if ent is not None else torch.empty(int(L.lib().tmpnn_bwd_tc_image_bytes()), dtype=torch.uint8, device=w.device)
        L.call('tmpnn_pack_w_tc', L.ptr(w), int(w.shape[1]), int(col0), L.ptr(img), L.stream())
        sc[key] = (w._version, img)
        return img
    return ent[1]


class _MPStepFn(torch.autograd.Function):
    """One ``TrackMPNN.forward`` with everything its backward needs kept on the device: the state the
    step consumed, the GRU gates of every row, the detection aggregates, Linear1 outputs and the
    BatchNorm statistics of the new rows, and the (immutable) window graph + incidence index."""

    @staticmethod
    def forward(ctx, model, node_adj, x, h_in, *params):
        wg = window_graph_of(node_adj)
        dev = wg.device
        n_tot = wg.n
        G = len(model.feature_idx)
        ldh = G * H
        saved_in = []
        if isinstance(x, NewRowGroups):
            # batched trainer: h_in already has one row per graph row (zeros where rows are new); the new detection rows
            # come in groups, one per chunk = one BatchNorm batch each (the reference normalises chunk by chunk)
            n_new = 0
            if x.carry is not None or h_in is None:
                h_cur = torch.zeros((n_tot, ldh), dtype=torch.float32, device=dev)
                n_old = 0 if h_in is None else int(h_in.shape[0])
                if h_in is not None and x.carry is not None:
                    seg_src, seg_dst, seg_len, n_seg, max_len = x.carry
                    L.call('tmpnn_rows_move', L.ptr(h_in.detach().contiguous()), L.ptr(h_cur), L.ptr(seg_src), L.ptr(seg_dst),
                           L.ptr(seg_len), n_seg, ldh, max_len, L.stream())
                ctx.carry = x.carry
            else:
                if int(h_in.shape[0]) != n_tot:
                    raise ValueError('with NewRowGroups and no carry map h_in must have one row per graph row')
                n_old = n_tot
                h_cur = h_in.detach().to(device=dev, dtype=torch.float32).clone().contiguous()
                ctx.carry = None
            saved_in = _input_groups_forward(model, x, h_cur, ldh)
        else:
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
            saved_in.append(_input_rows_forward(model, xd, new_det, out_rows, nd, n_new - nd, h_cur, ldh))
        ix = wg.index()
        h_out = torch.empty_like(h_cur)
        packs = packed_cells(model)
        gates, aggs, gats = [], [], []
        st = L.stream()
        for g in range(G):
            gru = model.factor_grus[g]
            concat = int(gru.msg_type == 'concat')
            gt = torch.empty((n_tot, 4 * H), dtype=torch.float32, device=dev)
            agg = torch.empty((ix.cap_dets, H), dtype=torch.float32, device=dev)
            if gru.gat is None and not concat and use_tensor_path(model, n_tot):
                # large graphs: the association rows' step on the tensor cores with the gates stored for the backward pass
                L.call('tmpnn_aggregate_dets', wg.g.c, ix.c, L.ptr(h_cur), ldh, g * H, L.ptr(agg), st)
                L.call('tmpnn_mp_edge_fwd_tc_train', wg.g.c, ix.c, L.ptr(h_cur), L.ptr(h_out), ldh, g, G,
                       L.ptr(packed_cells_tc(model)[g]), L.ptr(gt), st)
                L.call('tmpnn_mp_det_fwd_train', wg.g.c, ix.c, L.ptr(h_cur), L.ptr(h_out), ldh, g, G, L.ptr(packs[g][1]),
                       L.ptr(agg), L.ptr(gt), st)
                gats.append(None)
            elif gru.gat is None:
                L.call('tmpnn_mp_step_fwd_train', wg.g.c, ix.c, L.ptr(h_cur), L.ptr(h_out), ldh, g, G, concat,
                       L.ptr(packs[g][0]), L.ptr(packs[g][1]), L.ptr(agg), L.ptr(gt), st)
                gats.append(None)
            else:
                gats.append(gat_forward_train(gru, wg.g, ix, h_cur, ldh, g * H, agg))
                L.call('tmpnn_mp_step_fwd_train_agg', wg.g.c, ix.c, L.ptr(h_cur), L.ptr(h_out), ldh, g, G, concat,
                       L.ptr(packs[g][0]), L.ptr(packs[g][1]), L.ptr(agg), L.ptr(gt), st)
            gates.append(gt)
            aggs.append(agg)
        logits = wg.g.logit[:n_tot].clone().unsqueeze(1)
        scores = wg.g.score[:n_tot].clone().unsqueeze(1)
        ctx.model, ctx.wg, ctx.ix = model, wg, ix
        ctx.n_old, ctx.has_h_in = n_old, h_in is not None
        ctx.saved_in, ctx.gates, ctx.aggs, ctx.gats = saved_in, gates, aggs, gats
        # outputs are kept as detached aliases: an attribute holding an output tensor itself would close a reference
        # cycle (ctx -> output -> grad_fn -> ctx) that only the cycle collector frees -- GBs per batched step
        ctx.h_cur, ctx.h_out, ctx.p = h_cur, h_out.detach(), scores.detach()
        ctx.param_vals = [p.detach() for p in params]
        # the attention after dropout of every head, as extra non-differentiable outputs (copies: nothing ctx holds)
        alphas = [attention_after_dropout(sv).clone() for heads in gats if heads is not None for sv in heads]
        ctx.mark_non_differentiable(*alphas)
        return (scores, logits, h_out, *alphas)

    @staticmethod
    def backward(ctx, dscores, dlogits, dh_out, *_dalphas):
        model, wg, ix = ctx.model, ctx.wg, ctx.ix
        dev = wg.device
        n, G = wg.n, len(model.feature_idx)
        ldh = G * H
        st = L.stream()
        f32 = dict(dtype=torch.float32, device=dev)
        P = ctx.param_vals
        # parallel.FlatGradients: every p.grad is a view of one flat buffer and the kernels below accumulate straight into
        # it (nothing is returned through autograd for the parameters: no per-parameter zero-fill, add or copy)
        flat = model.__dict__.get('_tmpnn_flat_grads')
        direct = None
        if flat is not None:
            direct = [flat.view_of(p) for p in _param_list(model)]
            if len(direct) != len(P) or any(v is None for v in direct):
                direct = None
        grads = direct if direct is not None else [torch.zeros_like(p) for p in P]
        cont = lambda t: None if t is None else t.detach().to(**f32).contiguous()
        dscores, dlogits, dh_out = cont(dscores), cont(dlogits), cont(dh_out)
        dh_cur = torch.empty((n, ldh), **f32)   # every row / column block is assigned by tmpnn_scatter_bwd
        g_hw_node, g_hb_node, g_hw_edge, g_hb_edge = grads[-4], grads[-3], grads[-2], grads[-1]
        hw_node, hw_edge = P[-4], P[-2]
        h_cur, h_out = ctx.h_cur, ctx.h_out
        gat_base = G * _PER_GROUP   # W_att, a of every head, group by group
        for g in range(G):
            col = g * H
            b = g * _PER_GROUP
            e_wih, e_whh, n_wih, n_whh = P[b + 6], P[b + 7], P[b + 10], P[b + 11]
            kx = int(e_wih.shape[1])
            concat = int(kx == 2 * H)
            dgi = torch.empty((n, 3 * H), **f32)
            dgh = torch.empty((n, 3 * H), **f32)
            dhself = torch.empty((n, H), **f32)
            # bias gradients as [2][192] = d bias_ih | d bias_hh: in the flat gradient buffer the two vectors of a cell are
            # adjacent, so the kernel accumulates into them in place
            adj = lambda i: direct is not None and grads[i + 1].data_ptr() == grads[i].data_ptr() + 4 * 3 * H
            gb_e = grads[b + 8] if adj(b + 8) else torch.zeros((2, 3 * H), **f32)
            gb_d = grads[b + 12] if adj(b + 12) else torch.zeros((2, 3 * H), **f32)
            L.call('tmpnn_gate_bwd', n, L.ptr(wg.g.src), L.ptr(ctx.gates[g]), L.ptr(h_cur), L.ptr(h_out), ldh, col,
                   L.ptr(dh_out), L.ptr(dlogits), L.ptr(dscores), L.ptr(ctx.p), L.ptr(hw_edge) + 4 * col,
                   L.ptr(hw_node) + 4 * col, L.ptr(dgi), L.ptr(dgh), L.ptr(dhself), L.ptr(gb_e), L.ptr(gb_d),
                   L.ptr(g_hw_edge) + 4 * col, L.ptr(g_hw_node) + 4 * col,
                   L.ptr(g_hb_edge) if g == 0 else None, L.ptr(g_hb_node) if g == 0 else None, L.ptr(_gate_bwd_partials(model, dev)), st)
            if not adj(b + 8):
                grads[b + 8] += gb_e[0]; grads[b + 9] += gb_e[1]
            if not adj(b + 12):
                grads[b + 12] += gb_d[0]; grads[b + 13] += gb_d[1]
            # edge cell: rows with src >= 0
            dx = torch.empty((n, kx), **f32)
            xbuf = torch.empty((n, kx), **f32)
            L.call('tmpnn_aggregate_edges', wg.g.c, ix.c, L.ptr(h_cur), ldh, col, concat, L.ptr(xbuf), st)
            if use_tensor_path(model, n):
                # tcgen05 (csrc/train_tc.cu): dx = dgi W_ih and dW_ih += dgi^T x in one pass over dgi, then dh_self += dgh W_hh
                # and dW_hh += dgh^T h in one pass over dgh; a 128-wide W_ih (concat) is two 64-column passes
                part = _bwd_tc_partials(model, dev)
                for c0 in range(0, kx, H):
                    L.call('tmpnn_rows_gemm_tc', None, n, None, None, None, L.ptr(wg.g.src), L.ptr(dgi),
                           L.ptr(_bwd_tc_image(model, e_wih, c0)), L.ptr(dx) + 4 * c0, kx, 0, L.ptr(xbuf) + 4 * c0, kx,
                           L.ptr(part), L.ptr(grads[b + 6]) + 4 * c0, kx, L.ptr(wg.g.status), st)
                L.call('tmpnn_rows_gemm_tc', None, n, None, None, None, L.ptr(wg.g.src), L.ptr(dgh),
                       L.ptr(_bwd_tc_image(model, e_whh, 0)), L.ptr(dhself), H, 1, L.ptr(h_cur) + 4 * col, ldh,
                       L.ptr(part), L.ptr(grads[b + 7]), H, L.ptr(wg.g.status), st)
            else:
                L.call('tmpnn_rows_times_w', None, n, None, None, L.ptr(wg.g.src), L.ptr(dgi), L.ptr(e_wih), kx, L.ptr(dx), kx, 0, st)
                L.call('tmpnn_rows_times_w', None, n, None, None, L.ptr(wg.g.src), L.ptr(dgh), L.ptr(e_whh), H, L.ptr(dhself), H, 1, st)
                L.call('tmpnn_rows_outer', None, n, None, None, L.ptr(wg.g.src), L.ptr(dgi), L.ptr(xbuf), kx, kx, L.ptr(grads[b + 6]), st)
                L.call('tmpnn_rows_outer', None, n, None, None, L.ptr(wg.g.src), L.ptr(dgh), L.ptr(h_cur) + 4 * col, ldh, H,
                       L.ptr(grads[b + 7]), st)
            # node cell: the detection list (A gathered through det_rows; the count lives on the device)
            nd_dev, det_rows = L.ptr(ix.n_dets), L.ptr(ix.det_rows)
            dagg = (torch.empty if ctx.gats[g] is None else torch.zeros)((ix.cap_dets, H), **f32)
            if use_tensor_path(model, n):
                part = _bwd_tc_partials(model, dev)
                L.call('tmpnn_rows_gemm_tc', nd_dev, 0, det_rows, None, None, None, L.ptr(dgi), L.ptr(_bwd_tc_image(model, n_wih, 0)),
                       L.ptr(dagg), H, 0, L.ptr(ctx.aggs[g]), H, L.ptr(part), L.ptr(grads[b + 10]), H, L.ptr(wg.g.status), st)
                L.call('tmpnn_rows_gemm_tc', nd_dev, 0, det_rows, det_rows, det_rows, None, L.ptr(dgh),
                       L.ptr(_bwd_tc_image(model, n_whh, 0)), L.ptr(dhself), H, 1, L.ptr(h_cur) + 4 * col, ldh, L.ptr(part),
                       L.ptr(grads[b + 11]), H, L.ptr(wg.g.status), st)
            else:
                L.call('tmpnn_rows_times_w', nd_dev, 0, det_rows, None, None, L.ptr(dgi), L.ptr(n_wih), H, L.ptr(dagg), H, 0, st)
                L.call('tmpnn_rows_times_w', nd_dev, 0, det_rows, det_rows, None, L.ptr(dgh), L.ptr(n_whh), H, L.ptr(dhself), H, 1, st)
                L.call('tmpnn_rows_outer', nd_dev, 0, det_rows, None, None, L.ptr(dgi), L.ptr(ctx.aggs[g]), H, H, L.ptr(grads[b + 10]), st)
                L.call('tmpnn_rows_outer', nd_dev, 0, det_rows, det_rows, None, L.ptr(dgh), L.ptr(h_cur) + 4 * col, ldh, H,
                       L.ptr(grads[b + 11]), st)
            # through the gather / segmented sum, into the state this step consumed
            heads = ctx.gats[g]
            L.call('tmpnn_scatter_bwd', wg.g.c, ix.c, n, L.ptr(dhself), L.ptr(dx), kx,
                   L.ptr(dagg if heads is None else torch.zeros_like(dagg)), L.ptr(dh_cur), ldh, col, st)
            if heads is not None:
                # the aggregate came from the attention heads: through the weighted sum, the softmax and W_att, a
                dal = torch.empty(ix.cap_inc, **f32)
                de_side, dpre = torch.empty(2 * n, **f32), torch.empty(n, **f32)
                dhatt = torch.empty((ix.cap_dets, H), **f32)
                for k, sv in enumerate(heads):
                    L.call('tmpnn_gat_bwd', wg.g.c, ix.c, n, L.ptr(h_cur), ldh, col, L.ptr(sv['w']), L.ptr(sv['a']), len(heads),
                           L.ptr(sv['keep']), sv['scale'], L.ptr(sv['hatt']), L.ptr(sv['esc']), L.ptr(sv['alpha']),
                           L.ptr(sv['att_edge']), L.ptr(dagg), L.ptr(dal), L.ptr(de_side), L.ptr(dpre), L.ptr(dhatt),
                           L.ptr(dh_cur), L.ptr(grads[gat_base + 2 * k]), L.ptr(grads[gat_base + 2 * k + 1]), st)
                gat_base += 2 * len(heads)
            # new detection rows: through the input transform (one group of rows per BatchNorm batch)
            groups_in = [t for t in ctx.saved_in if t[3] > 0]
            if len(groups_in) > 1:
                # batched trainer: all chunks in one launch, one CTA per chunk (tmpnn_input_group descriptors)
                xd = groups_in[0][0]
                cols = model.feature_idx[g]
                tot = sum(t[3] for t in groups_in)
                scratch = torch.empty((2 * tot, H), **f32)
                # per-group partial gradients, added in group order by a second kernel (no atomics on the shared buffers)
                part = torch.empty(int(L.lib().tmpnn_input_bwd_partial_floats(len(groups_in))), **f32) if len(cols) <= H else None
                desc = (L.InputGroup * len(groups_in))()
                off = 0
                for k, (_, x_idx, out_rows, nd, n_edge_new, per_group) in enumerate(groups_in):
                    a, mean, var, training = per_group[g]
                    desc[k] = L.InputGroup(a.data_ptr(), mean.data_ptr(), var.data_ptr(), x_idx.data_ptr(), out_rows.data_ptr(),
                                           scratch.data_ptr() + 4 * H * 2 * off, nd, n_edge_new)
                    off += nd
                L.call('tmpnn_input_bwd_groups', L.ptr(xd), int(xd.shape[1]), int(cols[0]), len(cols), desc,
                       len(groups_in), L.ptr(P[b + 2]), L.ptr(P[b + 3]), L.ptr(P[b + 1]), L.ptr(P[b + 4]), L.ptr(dh_cur), ldh,
                       col, int(groups_in[0][5][g][3]), L.ptr(grads[b + 0]), L.ptr(grads[b + 1]), L.ptr(grads[b + 2]),
                       L.ptr(grads[b + 3]), L.ptr(grads[b + 4]), L.ptr(grads[b + 5]), L.ptr(part), st)
                groups_in = []
            for xd, new_det, out_rows, nd, n_edge_new, per_group in groups_in:
                a, mean, var, training = per_group[g]
                cols = model.feature_idx[g]
                scratch = torch.empty((2 * nd, H), **f32)
                L.call('tmpnn_input_bwd', L.ptr(xd), int(xd.shape[1]), int(cols[0]), len(cols), L.ptr(new_det), L.ptr(a),
                       L.ptr(mean), L.ptr(var), L.ptr(P[b + 2]), L.ptr(P[b + 3]), L.ptr(P[b + 1]), L.ptr(P[b + 4]),
                       L.ptr(dh_cur), ldh, col, L.ptr(out_rows), nd, n_edge_new, int(training), L.ptr(scratch),
                       L.ptr(grads[b + 0]), L.ptr(grads[b + 1]), L.ptr(grads[b + 2]), L.ptr(grads[b + 3]),
                       L.ptr(grads[b + 4]), L.ptr(grads[b + 5]), st)
        dh_in = dh_cur[:ctx.n_old] if ctx.has_h_in else None
        if ctx.has_h_in and getattr(ctx, 'carry', None) is not None:
            # back through the layout change: the rows of the previous step's layout collect the gradient of the rows they
            # were moved to (chunks that took no part in this step get none)
            seg_src, seg_dst, seg_len, n_seg, max_len = ctx.carry
            dh_in = torch.zeros((ctx.n_old, ldh), **f32)
            L.call('tmpnn_rows_move', L.ptr(dh_cur), L.ptr(dh_in), L.ptr(seg_dst), L.ptr(seg_src), L.ptr(seg_len), n_seg, ldh,
                   max_len, st)
        return (None, None, None, dh_in) + (tuple(None for _ in grads) if direct is not None else tuple(grads))
