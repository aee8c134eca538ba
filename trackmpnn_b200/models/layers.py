"""Parameter containers mirroring reference ``models/layers.py``.

``FactorGraphGRU`` keeps the reference's sub-module names (``edge_gru``, ``node_gru``) and
initialisation (``models/layers.py:52-82``) so state_dicts are interchangeable, but its math
is not run here: ``TrackMPNN.forward`` hands the parameters to the fused CUDA step
(``tmpnn_mp_step_fwd``).  Calling ``FactorGraphGRU.forward`` directly runs that same CUDA
step for a single feature group.
"""
import torch
import torch.nn as nn


class GraphAttentionLayer(nn.Module):
    """Parameter container of one attention head (reference ``models/layers.py:7-24``): ``W_att``
    [in, out] and ``a`` [out, 1], xavier-uniform with gain 1.414.  The attention itself runs in
    ``tmpnn_gat_aggregate_dets`` (``csrc/gat.cu``) on the edge list -- no dense N x N matrix."""

    def __init__(self, in_features, out_features, alpha=0.2, concat=False):
        super().__init__()
        if alpha != 0.2 or concat:
            raise NotImplementedError('the kernels implement the reference defaults (alpha=0.2, concat=False)')
        self.in_features, self.out_features, self.alpha, self.concat = in_features, out_features, alpha, concat
        self.W_att = nn.Parameter(torch.zeros(size=(in_features, out_features)))
        nn.init.xavier_uniform_(self.W_att.data, gain=1.414)
        self.a = nn.Parameter(torch.zeros(size=(out_features, 1)))
        nn.init.xavier_uniform_(self.a.data, gain=1.414)

    def __repr__(self):
        return self.__class__.__name__ + ' (' + str(self.in_features) + ' -> ' + str(self.out_features) + ')'


class FactorGraphGRU(nn.Module):
    """Different GRU cells for association ("edge") rows and detection ("node") rows."""

    def __init__(self, nhidden, nattheads=0, msg_type='diff', bias=True):
        super().__init__()
        assert msg_type in ('concat', 'diff'), 'Incorrect message type for model!'
        if nhidden != 64:
            raise NotImplementedError('the sm_100a kernels are specialised for --num-hidden-feats 64')
        self.nhidden, self.msg_type, self.nattheads, self.bias = nhidden, msg_type, nattheads, bias
        # construction order follows the reference (edge cell, heads, node cell) so that a seed gives the same init
        self.edge_gru = nn.GRUCell((2 if msg_type == 'concat' else 1) * nhidden, nhidden, bias=bias)
        self.gat = None if nattheads <= 0 else nn.ModuleList([GraphAttentionLayer(nhidden, nhidden) for _ in range(nattheads)])
        self.node_gru = nn.GRUCell(nhidden, nhidden, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        for cell in (self.edge_gru, self.node_gru):
            cell.weight_ih.data.normal_(mean=0.0, std=0.01)
            cell.weight_hh.data.normal_(mean=0.0, std=0.01)
            if self.bias is not None:
                cell.bias_ih.data.uniform_(0, 0)
                cell.bias_hh.data.uniform_(0, 0)

    def forward(self, h, node_adj, edge_adj):
        from ..functional import mp_step_single_group
        return mp_step_single_group(self, h, node_adj)

    def __repr__(self):
        return f'{self.__class__.__name__} ({self.nhidden} -> {self.nhidden})'
