"""``TrackMPNN`` with the reference's constructor, ``forward`` signature, return tuple and
``state_dict`` layout (reference ``models/track_mpnn.py:9-75``), computed by the sm_100a
library instead of dense N x N PyTorch ops.

    model = TrackMPNN(features='2d', ncategories=3, nhidden=64, nattheads=0, msg_type='diff')
    scores, logits, h_out, attention = model(x, h_in, node_adj, edge_adj)

Reference snapshots load with ``strict=True`` and ours load into the reference.  There is
no CPU path: parameters and the graph must live on a CUDA device.
"""
import torch
import torch.nn as nn

from .layers import FactorGraphGRU
from .. import functional as F_


class TrackMPNN(nn.Module):
    def __init__(self, features, ncategories, nhidden, nattheads, msg_type, return_attention=True,
                 use_tensor_cores='auto'):
        super().__init__()
        # 'auto': tcgen05 kernel for large graphs with msg_type 'diff', fp32 FMA kernel otherwise
        self.use_tensor_cores = use_tensor_cores
        self.input_transforms = nn.ModuleList([])
        self.factor_grus = nn.ModuleList([])
        self.feature_idx = []
        self.nhidden = nhidden
        # the reference's drivers unpack 3 values although forward returns 4;
        # return_attention=False gives the 3-tuple those drivers expect
        self.return_attention = return_attention
        nfeatures = 0
        # group order and the substring tests follow reference models/track_mpnn.py:17-33
        for tag, width in (('2d', ncategories + 5), ('temp', 2), ('vis', 128)):
            if tag in features:
                self.input_transforms.append(self.get_input_transform(width, nhidden))
                self.factor_grus.append(FactorGraphGRU(nhidden, nattheads, msg_type, True))
                self.feature_idx.append(list(range(nfeatures, nfeatures + width)))
                nfeatures += width
        ngroups = len(self.feature_idx)
        self.output_transform_node = self._head(ngroups * nhidden, +4.595)
        self.output_transform_edge = self._head(ngroups * nhidden, -4.595)
        self.output_activation = nn.Sigmoid()

    @staticmethod
    def _head(n_in, bias):
        lin = nn.Linear(n_in, 1, bias=True)
        lin.weight.data.normal_(mean=0.0, std=0.01)
        lin.bias.data.uniform_(bias, bias)
        return lin

    def get_input_transform(self, n_in, n_out):
        layers = []
        for a, b in ((n_in, n_out), (n_out, n_out)):
            lin = nn.Linear(a, b, bias=True)
            lin.weight.data.normal_(mean=0.0, std=0.01)
            lin.bias.data.uniform_(0, 0)
            layers.append(lin)
        return nn.Sequential(layers[0], nn.BatchNorm1d(n_out), nn.ReLU(), layers[1])

    def forward(self, x, h_in, node_adj, edge_adj):
        if not next(self.parameters()).is_cuda:
            raise RuntimeError('trackmpnn_b200.TrackMPNN runs on CUDA only (call model.cuda()); no CPU path exists')
        out = F_.track_mpnn_forward(self, x, h_in, node_adj, edge_adj)
        return out if self.return_attention else out[:3]
