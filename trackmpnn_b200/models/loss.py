"""Drop-in replacements for reference ``models/loss.py``: ``create_targets``, ``CELoss``, ``FocalLoss``.

Same names, argument order and return types; the per-detection segment logic runs on the incidence
CSR in ``libtmpnn_sm100a.so`` (``csrc/train.cu``) instead of dense N x N column scans, and both
losses are ``torch.autograd.Function``s whose backward is a CUDA kernel.
"""
import torch
import torch.nn as nn

from .. import _lib as L
from ..device_graph import window_graph_of


def _graph_with_labels(wg, labels):
    """ctypes view of the window graph whose label column is ``labels`` (the drivers pass the tensor
    that came back from update_graph, but any int tensor of the right length works)."""
    lab = labels.detach().to(device=wg.device, dtype=torch.int32).contiguous()
    g = wg.g
    c = L.Graph(g.num_seqs, g.cap_rows, L.ptr(g.n_rows), L.ptr(g.ts), L.ptr(g.det), L.ptr(g.ass), L.ptr(g.src),
                L.ptr(g.dst), L.ptr(lab), L.ptr(g.score), L.ptr(g.logit), L.ptr(g.status), None, None, None, None)
    return c, lab


def _check_idx_node(wg, idx_node):
    n_det = int(wg.index().n_dets.item())
    if int(idx_node.numel()) != n_det:
        raise NotImplementedError('idx_node must list every detection row (as train.py:71-74 passes it)')


def create_targets(labels, node_adj, idx_node):
    """Reference ``models/loss.py:8-44``: detection targets = labels; per detection the latest positive
    past edge and the earliest positive future edge get target 1.  Returns int64 ``targets[N]``."""
    import ctypes as C
    wg = window_graph_of(node_adj)
    _check_idx_node(wg, idx_node)
    c, lab = _graph_with_labels(wg, labels.view(-1))
    out = torch.zeros(max(1, wg.n), dtype=torch.int32, device=wg.device)
    L.call('tmpnn_loss_targets', C.byref(c), wg.index().c, wg.n, L.ptr(out), L.stream())
    return out[:wg.n].to(dtype=labels.dtype, device=labels.device)


class _CEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, wg):
        dev = wg.device
        ix = wg.index()
        n = wg.n
        lg = logits.detach().to(device=dev, dtype=torch.float32).contiguous().view(-1)
        tg = targets.detach().to(device=dev, dtype=torch.int32).contiguous().view(-1)
        nseg = 2 * ix.cap_dets + 2
        seg_lse = L.zeros(nseg, dev)
        seg_pos = torch.full((nseg,), -1, dtype=torch.int32, device=dev)
        seg_loss = L.zeros(nseg, dev)
        loss = L.zeros(1, dev)
        L.call('tmpnn_loss_ce_fwd', ix.c, n, L.ptr(tg), L.ptr(lg), L.ptr(seg_lse), L.ptr(seg_pos), L.ptr(seg_loss),
               L.ptr(loss), L.stream())
        ctx.wg, ctx.ix, ctx.lg, ctx.seg_lse, ctx.seg_pos = wg, ix, lg, seg_lse, seg_pos
        ctx.shape, ctx.dev = logits.shape, logits.device
        return loss[0].to(logits.device)

    @staticmethod
    def backward(ctx, gout):
        wg = ctx.wg
        g = gout.detach().to(device=wg.device, dtype=torch.float32).contiguous().view(1)
        d = L.zeros(max(1, wg.n), wg.device)
        L.call('tmpnn_loss_ce_bwd', wg.g.c, ctx.ix.c, wg.n, L.ptr(ctx.seg_lse), L.ptr(ctx.seg_pos), L.ptr(ctx.lg), L.ptr(g),
               L.ptr(d), L.stream())
        return d[:wg.n].view(ctx.shape).to(ctx.dev), None, None


class CELoss(nn.Module):
    """Reference ``models/loss.py:77-115``."""

    def forward(self, outputs, targets, node_adj, idx_node):
        wg = window_graph_of(node_adj)
        _check_idx_node(wg, idx_node)
        return _CEFn.apply(outputs, targets, wg)


class _FocalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, targets, mean):
        n = int(p.numel())
        pc = p.detach().to(torch.float32).contiguous().view(-1)
        tc = targets.detach().to(device=p.device, dtype=torch.int64).contiguous().view(-1)
        per = torch.empty(n, dtype=torch.float32, device=p.device)
        loss = torch.zeros(1, dtype=torch.float32, device=p.device)
        L.call('tmpnn_loss_focal_fwd', n, L.ptr(pc), L.ptr(tc), L.ptr(per), L.ptr(loss), L.stream())
        ctx.pc, ctx.tc, ctx.n, ctx.mean, ctx.shape = pc, tc, n, mean, p.shape
        return loss[0] if mean else loss[0] * n

    @staticmethod
    def backward(ctx, gout):
        g = gout.detach().to(torch.float32).contiguous().view(1)
        if not ctx.mean:
            g = g * ctx.n
        dp = torch.empty(ctx.n, dtype=torch.float32, device=ctx.pc.device)
        L.call('tmpnn_loss_focal_bwd', ctx.n, L.ptr(ctx.pc), L.ptr(ctx.tc), L.ptr(g), L.ptr(dp), L.stream())
        return dp.view(ctx.shape), None, None


class FocalLoss(nn.Module):
    """Reference ``models/loss.py:47-74`` as the drivers configure it (``train.py:333-334``: gamma = 0,
    alpha = None): ``mean(-log(p_t + 1e-10))`` with p_t = p for target 1 and 1 - p for target 0."""

    def __init__(self, gamma=0, alpha=None, size_average=True):
        super().__init__()
        if gamma != 0 or alpha is not None:
            raise NotImplementedError('only gamma=0, alpha=None (the configuration train.py uses) is built')
        self.gamma, self.alpha, self.size_average, self.eps = gamma, alpha, size_average, 1e-10

    def forward(self, outputs, targets):
        if outputs.numel() == 0:  # reference: mean of an empty tensor
            return outputs.sum() * float('nan') if self.size_average else outputs.sum()
        if not outputs.is_cuda:
            raise RuntimeError('trackmpnn_b200.FocalLoss runs on CUDA only; no CPU path exists')
        return _FocalFn.apply(outputs, targets, bool(self.size_average))
