"""Helpers shared by the CPU (oracle) and GPU (CUDA path) replays of tests/golden/*.npz."""
import ast
import glob
import os

import numpy as np

from oracle import trackmpnn_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def attention_keep_matrix(seed, step, group, head, n):
    """The dropout decision for every entry of the dense [n, n] attention of (step, group, head): the reference's
    nn.Dropout draws from torch's CPU generator, which no other implementation can replay, so the train fixtures with
    attention heads pin the mask instead -- the generator script replaces ``GraphAttentionLayer.dropout`` by a
    multiplication with ``keep / (1 - p)`` (p = 0.5), and every replay uses the same matrix."""
    return np.random.RandomState(seed * 100000 + step * 100 + group * 10 + head).rand(n, n) >= 0.5


def golden_names(kind=None):
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))
    if kind is not None:
        names = [n for n in names if n.startswith(kind)]
    return names


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
        self.meta = ast.literal_eval(str(self.z['meta']))
        self.n_steps = int(self.z['n_steps'])
        self.X = self.z['X']
        self.y = self.z['y']

    def params(self, prefix='w/'):
        return {k[len(prefix):]: self.z[k].copy() for k in self.z.files if k.startswith(prefix)}

    def has(self, s, key):
        return f's{s}/{key}' in self.z.files

    def get(self, s, key):
        return self.z[f's{s}/{key}']

    def graph(self, s, prefix=''):
        """Edge-list Graph for the reference's (y_pred, node_adj[, labels]) at step s."""
        yp = self.get(s, prefix + 'y_pred').astype(np.int64)
        r, c, v = self.get(s, prefix + 'adj').astype(np.int64)
        n = yp.shape[0]
        src = -np.ones(n, np.int64); dst = -np.ones(n, np.int64)
        off = r != c
        pos = off & (v > 0); neg = off & (v < 0)
        src[r[pos]] = c[pos]; dst[r[neg]] = c[neg]
        lab = self.get(s, prefix + 'labels').astype(np.int64) if self.has(s, prefix + 'labels') else None
        return O.Graph(yp[:, 0].copy(), yp[:, 1].copy(), yp[:, 2].copy(), src, dst, lab)


def assert_graph_equal(a, b, what=''):
    for f in ('ts', 'det', 'ass', 'src', 'dst'):
        np.testing.assert_array_equal(getattr(a, f), getattr(b, f), err_msg=f'{what}: {f}')
    if a.label is not None and b.label is not None:
        np.testing.assert_array_equal(a.label, b.label, err_msg=f'{what}: label')


def assert_grads_close(got, gold, rel=2e-3):
    """Parameter gradients of a BPTT chunk against the reference's (``g/<name>`` in a train fixture).
    Per parameter: max-abs error <= rel * max|want| + 1e-6 * (largest gradient entry of the model); the
    second term covers gradients that are analytically zero (the Linear bias in front of a train-mode
    BatchNorm) and hold nothing but round-off in the reference too."""
    names = [k[2:] for k in gold.z.files if k.startswith('g/')]
    gmax = max(float(np.abs(gold.z['g/' + k]).max()) for k in names)
    for k in names:
        want = gold.z['g/' + k]
        atol = rel * float(np.abs(want).max()) + 1e-6 * gmax + 1e-9
        np.testing.assert_allclose(np.asarray(got[k]).reshape(want.shape), want, atol=atol, rtol=0, err_msg=k)


def pin_attention_dropout(model, gold, state):
    """Makes the CUDA model's attention heads use the fixture's dropout decisions: ``state['s']`` (MP step) and
    ``state['n']`` (rows of the graph at that step) are set by the test before each forward; the dense decision matrix
    is read at (detection row, incident edge row) for every entry of the incidence index."""
    import torch
    seed = gold.meta.get('seed')
    for g, gru in enumerate(model.factor_grus):
        if gru.gat is None:
            continue

        def fn(head, index, device, g=g):
            M = attention_keep_matrix(seed, state['s'], g, head, state['n'])
            nd = int(index.n_dets.item())
            seg = index.seg_ptr[:2 * nd + 1].long().cpu()
            lens = seg[2::2] - seg[0:-1:2]
            tot = int(seg[-1])
            owner = torch.repeat_interleave(index.det_rows[:nd].long().cpu(), lens).numpy()
            cols = index.inc[:tot].long().cpu().numpy()
            keep = torch.zeros(index.cap_inc, dtype=torch.uint8)
            keep[:tot] = torch.from_numpy(M[owner, cols].astype(np.uint8))
            return keep.to(device)
        gru.attention_keep_fn = fn
