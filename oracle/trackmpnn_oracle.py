"""CPU oracle for the TrackMPNN message-passing hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *checker*, not the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  Nothing under ``trackmpnn_b200/`` imports it, and the product
path raises when its CUDA library is missing instead of falling back to this.

It restates, in edge-list (never dense N x N) numpy form, the behaviour of the
reference at ``/root/reference``:

* ``forward``            <- ``models/track_mpnn.py:54-75`` + ``models/layers.py:84-116``
* ``initialize_graph``   <- ``utils/graph.py:96-186``
* ``update_graph``       <- ``utils/graph.py:189-334``
* ``prune_graph``        <- ``utils/graph.py:337-389``
* ``decode_tracks``      <- ``utils/graph.py:392-539``
* ``hungarian``          <- ``utils/graph.py:33-93`` (calls scipy's
  ``linear_sum_assignment`` exactly like the reference does; scipy is the
  reference's own un-pinned third-party dependency, ``Pipfile:14``)
* ``create_targets`` / ``ce_loss`` / ``focal_loss`` <- ``models/loss.py:8-115``

Parity pin: the reference holds no golden vectors or tests of its own
(SURVEY.md section 4), so this oracle is pinned against outputs of the reference
itself, generated in the build container by ``tests/golden/make_golden.py``
(which imports ``/root/reference`` read-only) and committed under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them.

Graph state is a ``Graph`` of int64 numpy arrays, one entry per row of the
reference's ``y_pred`` / adjacency:
  ts[i]   timestamp of a detection row, -1 for an association ("edge") row
  det[i]  detection id (index into the sequence's X / y), -1 for edge rows
  ass[i]  detection id this detection is associated to, -1 if none
  src[i], dst[i]  the two endpoint rows of an edge row (src < i < dst), -1 for detections
  label[i] binary ground-truth class (None when the caller has no labels)
``node_adj[e, src] = +1, node_adj[e, dst] = -1, node_adj[d, d] = 1`` and
``edge_adj = node_adj^T`` off the diagonal with ``edge_adj[e, e] = 1``
(reference ``utils/graph.py:152-163, 299-308``).
"""
from dataclasses import dataclass
import numpy as np

try:  # the reference's own dependency for --hungarian (utils/graph.py:18)
    from scipy.optimize import linear_sum_assignment
except Exception:  # pragma: no cover
    linear_sum_assignment = None


# --------------------------------------------------------------------------------------
# graph container
# --------------------------------------------------------------------------------------
@dataclass
class Graph:
    ts: np.ndarray
    det: np.ndarray
    ass: np.ndarray
    src: np.ndarray
    dst: np.ndarray
    label: np.ndarray = None

    @property
    def n(self):
        return int(self.ts.shape[0])

    def copy(self):
        return Graph(self.ts.copy(), self.det.copy(), self.ass.copy(), self.src.copy(), self.dst.copy(),
                     None if self.label is None else self.label.copy())

    def y_pred(self):
        """[N, 3] int64 rows [ts, det_id, ass_id] (reference ``utils/graph.py:137-141``)."""
        return np.stack((self.ts, self.det, self.ass), axis=1).astype(np.int64)

    def dense_node_adj(self):
        """Dense node_adj incl. the I_node diagonal -- small cases only (tests)."""
        n = self.n
        a = np.zeros((n, n), dtype=np.float32)
        e = np.nonzero(self.ts < 0)[0]
        a[e, self.src[e]] = 1.0
        a[e, self.dst[e]] = -1.0
        d = np.nonzero(self.ts >= 0)[0]
        a[d, d] = 1.0
        return a

    def dense_edge_adj(self):
        n = self.n
        a = np.zeros((n, n), dtype=np.float32)
        e = np.nonzero(self.ts < 0)[0]
        a[self.src[e], e] = 1.0
        a[self.dst[e], e] = -1.0
        a[e, e] = 1.0
        return a

    @staticmethod
    def from_dense(y_pred, node_adj, labels=None):
        """Recover the edge list from a dense node_adj (with or without diagonal)."""
        y_pred = np.asarray(y_pred).astype(np.int64)
        a = np.array(node_adj, dtype=np.float32, copy=True)
        n = a.shape[0]
        a[np.arange(n), np.arange(n)] = 0
        src = -np.ones(n, dtype=np.int64)
        dst = -np.ones(n, dtype=np.int64)
        for e in np.nonzero(y_pred[:, 0] < 0)[0]:
            p = np.nonzero(a[e] > 0)[0]
            m = np.nonzero(a[e] < 0)[0]
            assert p.size == 1 and m.size == 1
            src[e], dst[e] = p[0], m[0]
        return Graph(y_pred[:, 0].copy(), y_pred[:, 1].copy(), y_pred[:, 2].copy(), src, dst,
                     None if labels is None else np.asarray(labels).astype(np.int64))


# --------------------------------------------------------------------------------------
# model forward
# --------------------------------------------------------------------------------------
def feature_groups(features, ncategories):
    """Column slices per feature group, reference ``models/track_mpnn.py:17-33``
    (substring tests, order 2d / temp / vis)."""
    groups, n = [], 0
    if '2d' in features:
        groups.append((n, n + ncategories + 5)); n += ncategories + 5
    if 'temp' in features:
        groups.append((n, n + 2)); n += 2
    if 'vis' in features:
        groups.append((n, n + 128)); n += 128
    return groups


def init_params(features, ncategories, nhidden, msg_type, seed=5, scale=1.0, edge_bias=None):
    """Random parameters with the reference's state_dict names and init statistics
    (``models/track_mpnn.py:35-52``, ``models/layers.py:70-82``).  Not bit-identical to
    torch's RNG -- tests that need the reference's exact weights export its state_dict."""
    rs = np.random.RandomState(seed)
    H = nhidden
    groups = feature_groups(features, ncategories)
    p = {}
    for g, (a, b) in enumerate(groups):
        p[f'input_transforms.{g}.0.weight'] = rs.normal(0, 0.01, (H, b - a)).astype(np.float32) * scale
        p[f'input_transforms.{g}.0.bias'] = np.zeros(H, np.float32)
        p[f'input_transforms.{g}.1.weight'] = np.ones(H, np.float32)
        p[f'input_transforms.{g}.1.bias'] = np.zeros(H, np.float32)
        p[f'input_transforms.{g}.1.running_mean'] = np.zeros(H, np.float32)
        p[f'input_transforms.{g}.1.running_var'] = np.ones(H, np.float32)
        p[f'input_transforms.{g}.1.num_batches_tracked'] = np.zeros((), np.int64)
        p[f'input_transforms.{g}.3.weight'] = rs.normal(0, 0.01, (H, H)).astype(np.float32) * scale
        p[f'input_transforms.{g}.3.bias'] = np.zeros(H, np.float32)
        ein = 2 * H if msg_type == 'concat' else H
        for cell, cin in (('edge_gru', ein), ('node_gru', H)):
            p[f'factor_grus.{g}.{cell}.weight_ih'] = rs.normal(0, 0.01, (3 * H, cin)).astype(np.float32) * scale
            p[f'factor_grus.{g}.{cell}.weight_hh'] = rs.normal(0, 0.01, (3 * H, H)).astype(np.float32) * scale
            p[f'factor_grus.{g}.{cell}.bias_ih'] = np.zeros(3 * H, np.float32)
            p[f'factor_grus.{g}.{cell}.bias_hh'] = np.zeros(3 * H, np.float32)
    G = len(groups)
    p['output_transform_node.weight'] = rs.normal(0, 0.01, (1, G * H)).astype(np.float32) * scale
    p['output_transform_node.bias'] = np.full(1, 4.595, np.float32)
    p['output_transform_edge.weight'] = rs.normal(0, 0.01, (1, G * H)).astype(np.float32) * scale
    p['output_transform_edge.bias'] = np.full(1, -4.595 if edge_bias is None else edge_bias, np.float32)
    return p


def _sigmoid(x):
    x = x.astype(np.float32)
    return (1.0 / (1.0 + np.exp(-x, dtype=np.float32))).astype(np.float32)


def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """torch.nn.GRUCell, gate order r, z, n (used at ``models/layers.py:97,114``)."""
    H = h.shape[1]
    gi = x @ w_ih.T + b_ih
    gh = h @ w_hh.T + b_hh
    r = _sigmoid(gi[:, :H] + gh[:, :H])
    z = _sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:]).astype(np.float32)
    return ((1.0 - z) * n + z * h).astype(np.float32)


def input_transform(params, g, x_det, n_edge_rows, training=False, momentum=0.1, eps=1e-5):
    """Linear -> BatchNorm1d -> ReLU -> Linear on the detection rows of the new block
    (``models/track_mpnn.py:45-52,59``).  The reference also pushes the all-zero edge rows
    through it and masks them to zero afterwards (``:61``); in train mode those
    ``n_edge_rows`` identical rows (value b1 after Linear1) take part in the batch
    statistics, which is reproduced here.  Returns (out_det, new_running_stats | None)."""
    w1 = params[f'input_transforms.{g}.0.weight']; b1 = params[f'input_transforms.{g}.0.bias']
    gam = params[f'input_transforms.{g}.1.weight']; bet = params[f'input_transforms.{g}.1.bias']
    rm = params[f'input_transforms.{g}.1.running_mean']; rv = params[f'input_transforms.{g}.1.running_var']
    w2 = params[f'input_transforms.{g}.3.weight']; b2 = params[f'input_transforms.{g}.3.bias']
    a = (x_det @ w1.T + b1).astype(np.float32)
    new_stats = None
    if training:
        n = a.shape[0] + n_edge_rows
        if n <= 1:
            raise ValueError("Expected more than 1 value per channel when training")
        mu = (a.sum(0, dtype=np.float64) + n_edge_rows * b1.astype(np.float64)) / n
        var = (((a.astype(np.float64) - mu) ** 2).sum(0) + n_edge_rows * (b1.astype(np.float64) - mu) ** 2) / n
        new_stats = ((1 - momentum) * rm + momentum * mu.astype(np.float32),
                     (1 - momentum) * rv + momentum * (var * n / (n - 1)).astype(np.float32))
        mu = mu.astype(np.float32); var = var.astype(np.float32)
    else:
        mu, var = rm, rv
    bn = (a - mu) / np.sqrt(var + np.float32(eps)) * gam + bet
    out = np.maximum(bn, 0).astype(np.float32) @ w2.T + b2
    return out.astype(np.float32), new_stats


def aggregate(h, g, msg_type='diff'):
    """The two sparse products of ``models/layers.py:90-95,103`` in edge-list form.
    Returns (node_support for edge rows [n_e, H or 2H], edge_support for det rows [n_d, H],
    edge_rows, det_rows)."""
    e = np.nonzero(g.ts < 0)[0]
    d = np.nonzero(g.ts >= 0)[0]
    if msg_type == 'concat':
        xs = np.concatenate((h[g.src[e]], h[g.dst[e]]), axis=1)
    else:
        xs = h[g.src[e]] - h[g.dst[e]]
    agg = np.zeros_like(h)
    # ascending edge order within each detection, like the row-major sparse product
    np.add.at(agg, g.src[e], h[e])
    np.add.at(agg, g.dst[e], -h[e])
    return xs.astype(np.float32), agg[d].astype(np.float32), e, d


def gat_heads(params, gi):
    """Attention heads of feature group gi present in a state_dict (``factor_grus.gi.gat.k.{W_att,a}``)."""
    k = 0
    while f'factor_grus.{gi}.gat.{k}.W_att' in params:
        k += 1
    return k


def gat_support(h, g, w_att, a, alpha=0.2, keep=None):
    """One ``GraphAttentionLayer.forward`` (``models/layers.py:26-43``) on the edge list; eval mode, or -- with
    ``keep`` (bool [N, N], the dropout decision of every attention entry) -- train mode, where the softmax output is
    multiplied by ``keep / (1 - 0.5)`` before the weighted sum (``:37``; the returned attention is after dropout):
    e_j = LeakyReLU(a . |W h[src_j] - W h[dst_j]|) per edge row, softmax over the edges incident to each
    detection, h'_d = sum_j alpha_dj (+1 if d is the source of j else -1) h_j.
    Returns (support for every row [N, H] -- zero off the detection rows --, e [N], per-detection
    dict {row: (incident edge rows, alpha)})."""
    n = g.n
    h_att = (h @ w_att).astype(np.float32)
    e_rows = np.nonzero(g.ts < 0)[0]
    ev = np.zeros(n, np.float32)
    if e_rows.size:
        d = np.abs(h_att[g.src[e_rows]] - h_att[g.dst[e_rows]]).astype(np.float32)
        v = (d @ a.reshape(-1)).astype(np.float32)
        ev[e_rows] = np.where(v > 0, v, np.float32(alpha) * v)
    past, fut = _segments(g)
    out = np.zeros_like(h)
    att = {}
    for r in np.nonzero(g.ts >= 0)[0]:
        inc = np.asarray(sorted(past.get(int(r), []) + fut.get(int(r), [])), np.int64)
        if inc.size == 0:
            continue
        x = ev[inc]
        w = np.exp(x - x.max(), dtype=np.float32)
        w = (w / w.sum(dtype=np.float32)).astype(np.float32)
        if keep is not None:
            w = (w * keep[r, inc].astype(np.float32) * np.float32(2.0)).astype(np.float32)
        sign = np.where(g.src[inc] == r, 1.0, -1.0).astype(np.float32)
        out[r] = ((w * sign)[:, None] * h[inc]).sum(0, dtype=np.float32)
        att[int(r)] = (inc, w)
    return out, ev, att


def dense_attention(att, n):
    """The reference's dense [N, N] attention of one head: softmax rows of the detections, and the
    uniform 1/N rows its masked softmax produces for rows without incident edges."""
    A = np.full((n, n), np.float32(1.0) / np.float32(n), np.float32)
    for r, (inc, w) in att.items():
        A[r] = 0
        A[r, inc] = w
    return A


def forward(params, x_new, h_in, g, features='2d', ncategories=3, nhidden=64, msg_type='diff',
            training=False, update_running_stats=True, return_attention=False, attention_keep=None):
    """``TrackMPNN.forward`` (``models/track_mpnn.py:54-75``).  With attention heads in training mode
    ``attention_keep(group, head, n) -> bool [n, n]`` supplies the dropout decisions (the reference draws them from
    torch's generator; fixtures pin them instead).

    x_new [N'-N, F] float32 (edge rows all-zero), h_in [N, G*H] or None, g has N' rows.
    Returns scores [N',1], logits [N',1], h_out [N', G*H] (attention is a tuple of None)."""
    H = nhidden
    groups = feature_groups(features, ncategories)
    G = len(groups)
    n_tot = g.n
    n_new = x_new.shape[0]
    n_old = 0 if h_in is None else h_in.shape[0]
    assert n_old + n_new == n_tot, (n_old, n_new, n_tot)
    is_det = g.ts >= 0
    h_out = np.zeros((n_tot, G * H), np.float32)
    attention = []
    for gi, (a, b) in enumerate(groups):
        if n_new > 0:
            new_det = is_det[n_old:]
            n_edge_new = int((~new_det).sum())
            upd = np.zeros((n_new, H), np.float32)
            out_det, stats = input_transform(params, gi, x_new[new_det][:, a:b], n_edge_new, training)
            upd[new_det] = out_det
            if stats is not None and update_running_stats:
                params[f'input_transforms.{gi}.1.running_mean'] = stats[0]
                params[f'input_transforms.{gi}.1.running_var'] = stats[1]
                params[f'input_transforms.{gi}.1.num_batches_tracked'] = \
                    params[f'input_transforms.{gi}.1.num_batches_tracked'] + 1
            h = upd if h_in is None else np.concatenate((h_in[:, gi * H:(gi + 1) * H], upd), axis=0)
        else:
            h = h_in[:, gi * H:(gi + 1) * H]
        h = np.ascontiguousarray(h, dtype=np.float32)
        xs, agg, e, d = aggregate(h, g, msg_type)
        nheads = gat_heads(params, gi)
        if nheads:  # attention-weighted edge_support instead of the plain signed sum (models/layers.py:105-112)
            assert not training or attention_keep is not None, 'train-mode attention needs the dropout decisions'
            acc = np.zeros_like(h)
            heads = []
            for k in range(nheads):
                sup, _, att = gat_support(h, g, params[f'factor_grus.{gi}.gat.{k}.W_att'], params[f'factor_grus.{gi}.gat.{k}.a'],
                                          keep=attention_keep(gi, k, n_tot) if training else None)
                acc = acc + sup
                heads.append(att)
            agg = (acc / np.float32(nheads)).astype(np.float32)[d]
            attention.append(heads)
        else:
            attention.append(None)
        pre = f'factor_grus.{gi}.'
        hn = np.empty_like(h)
        hn[e] = gru_cell(xs, h[e], params[pre + 'edge_gru.weight_ih'], params[pre + 'edge_gru.weight_hh'],
                         params[pre + 'edge_gru.bias_ih'], params[pre + 'edge_gru.bias_hh'])
        hn[d] = gru_cell(agg, h[d], params[pre + 'node_gru.weight_ih'], params[pre + 'node_gru.weight_hh'],
                         params[pre + 'node_gru.bias_ih'], params[pre + 'node_gru.bias_hh'])
        h_out[:, gi * H:(gi + 1) * H] = hn
    wn = params['output_transform_node.weight'][0]; bn = params['output_transform_node.bias'][0]
    we = params['output_transform_edge.weight'][0]; be = params['output_transform_edge.bias'][0]
    logits = np.where(is_det, h_out @ wn + bn, h_out @ we + be).astype(np.float32)[:, None]
    if return_attention:
        return _sigmoid(logits), logits, h_out, attention
    return _sigmoid(logits), logits, h_out


# --------------------------------------------------------------------------------------
# graph bookkeeping
# --------------------------------------------------------------------------------------
def _times(y):
    return np.asarray(y)[0, :, 0]


def initialize_graph(X, y, t_st=0, mode='test'):
    """``utils/graph.py:96-186``.  X [1, ND, F], y [1, ND, 2] float ([ts, track_id]).
    Returns (Graph, feats [N, F], t1+1, tN+1) or None (the reference's 7 x None)."""
    X = np.asarray(X); y = np.asarray(y)
    assert X.shape[0] == y.shape[0] == 1, "Only batch size 1 supported!"
    assert X.shape[1] == y.shape[1], "Input dimension mismatch!"
    times = np.sort(y[0, :, 0])
    later = times[times >= t_st]
    t0 = later[0]
    tN = times[-1]
    after = times[times > t0]
    t1 = after[0] if after.size else t0
    t0, t1, tN = int(t0), int(t1), int(tN)
    if t0 == t1 or ((y[0, :, 1] == -1).all() and mode == 'train'):
        return None
    ids0 = np.nonzero(y[0, :, 0] == t0)[0]
    ids1 = np.nonzero(y[0, :, 0] == t1)[0]
    n0, n1 = ids0.size, ids1.size
    n = n0 + n0 * n1 + n1
    ts = -np.ones(n, np.int64); det = -np.ones(n, np.int64); ass = -np.ones(n, np.int64)
    src = -np.ones(n, np.int64); dst = -np.ones(n, np.int64)
    ts[:n0] = t0; ts[n0 + n0 * n1:] = t1
    det[:n0] = ids0; det[n0 + n0 * n1:] = ids1
    k = np.arange(n0 * n1)
    src[n0 + k] = k // n1                       # edge (i, j) sits at n0 + i*n1 + j  (:152-156)
    dst[n0 + k] = n0 + n0 * n1 + (k % n1)
    feats = np.concatenate((X[0, ids0], np.zeros((n0 * n1, X.shape[2]), X.dtype), X[0, ids1]), 0)
    label = np.zeros(n, np.int64)
    tr0 = y[0, ids0, 1]; tr1 = y[0, ids1, 1]
    label[:n0] = tr0 >= 0
    label[n0 + n0 * n1:] = tr1 >= 0
    for j in range(n1):
        if tr1[j] == -1:
            continue
        idx = np.nonzero(tr0 == tr1[j])[0]
        if idx.size == 1:
            label[n0 + idx[0] * n1 + j] = 1
        elif idx.size > 1:
            assert False, "More than one detection from same timestep assinged to same track!"
    return Graph(ts, det, ass, src, dst, label), feats.astype(np.float32), t1 + 1, tN + 1


def _out_edges(g):
    """Ascending out-edge rows per source row (dict row -> ndarray)."""
    e = np.nonzero(g.ts < 0)[0]
    order = np.argsort(g.src[e], kind='stable')
    es = e[order]; ss = g.src[es]
    out = {}
    if es.size:
        cuts = np.nonzero(np.diff(ss))[0] + 1
        for seg in np.split(np.arange(es.size), cuts):
            out[int(ss[seg[0]])] = es[seg]
    return out


def associate_greedy(g, p):
    """``utils/graph.py:251-268`` == ``:437-454``.  p = scores[:, 1]."""
    ass = -np.ones(g.n, np.int64)
    out = _out_edges(g)
    det_rows = np.nonzero(g.ts >= 0)[0]
    for i in det_rows:
        if not (p[i] >= 0.5):
            continue
        ids = out.get(int(i))
        if ids is None:
            continue
        idx = ids[p[ids] >= 0.5]
        idx = idx[p[g.dst[idx]] >= 0.5] if idx.size else idx
        if idx.size > 0:
            nxt = det_rows[det_rows > idx[0]][0]      # first detection row after the first positive edge
            idx = idx[idx < nxt]                      # edges of the nearest timestep only
            best = idx[np.argmax(p[idx])]             # first arg-max
            ass[i] = g.det[g.dst[best]]
    return ass


def hungarian(g, scores, ass, t, threshold=0.5):
    """``utils/graph.py:33-93`` for one timestep t; updates ``ass`` in place."""
    cur = np.nonzero(g.ts == t)[0]
    if cur.size == 0:
        return ass
    e_all = np.nonzero(g.ts < 0)[0]
    inc = e_all[np.isin(g.dst[e_all], cur)]
    if inc.size == 0:
        return ass
    prev = np.unique(g.src[inc])
    prev = prev[ass[prev] == -1]
    C = np.full((prev.size, cur.size), 100.0, dtype=np.float32)
    pi = {int(r): k for k, r in enumerate(prev)}
    ci = {int(r): k for k, r in enumerate(cur)}
    seen = set()
    for e in inc:
        a = pi.get(int(g.src[e]))
        if a is None:
            continue
        key = (a, ci[int(g.dst[e])])
        assert key not in seen, "Two detection nodes connected through more than one edge!"
        seen.add(key)
        C[key] = scores[e, 0]
    r, c = linear_sum_assignment(C)
    for i, j in zip(r, c):
        if C[i, j] > threshold:
            continue
        ass[prev[i]] = g.det[cur[j]]
    return ass


def associate_hungarian(g, scores):
    """The driver loop ``utils/graph.py:247-249`` == ``:433-435``."""
    ass = -np.ones(g.n, np.int64)
    for t in range(int(g.ts[0]), int(g.ts[-1]) + 1):
        ass = hungarian(g, scores, ass, t)
    return ass


def associate_teacher(g):
    """Train-mode teacher forcing, ``utils/graph.py:229-245``."""
    ass = -np.ones(g.n, np.int64)
    out = _out_edges(g)
    for i in np.nonzero(g.ts >= 0)[0]:
        if g.label[i] == 1:
            ids = out.get(int(i))
            if ids is None:
                continue
            idx = ids[g.label[ids] != 0]
            if idx.size == 0:
                continue
            elif idx.size == 1:
                ass[i] = g.det[g.dst[idx[0]]]
            else:
                assert False, "More than one GT edge from same node!"
        else:
            ass[i] = g.det[i]
    return ass


def update_graph(g, scores, X, y, t, use_hungarian=False, mode='test'):
    """``utils/graph.py:189-334``.  scores [N, 2] = [1-p, p].  Returns (Graph', feats_new)."""
    X = np.asarray(X); y_i = np.asarray(y)[0].astype(np.int64)
    assert X.shape[0] == 1, "Only batch size 1 supported!"
    g = g.copy()
    p = np.asarray(scores, dtype=np.float32)[:, 1]
    if mode == 'train':
        g.ass = associate_teacher(g)
    elif use_hungarian:
        g.ass = associate_hungarian(g, np.asarray(scores, dtype=np.float32))
    else:
        g.ass = associate_greedy(g, p)
    n_past = g.n
    if mode == 'train':
        t_prev = np.amax(g.ts[g.ts < t])     # edge rows have ts=-1 < t, like the reference's y_pred[:,0] < t
        active = np.nonzero(((g.ts != -1) & (g.ass == -1)) | (g.ts == t_prev))[0]
    else:
        active = np.nonzero((g.ts != -1) & (g.ass == -1) & (p >= 0.5))[0]
    ids_t = np.nonzero(y_i[:, 0] == t)[0]
    A, nt = active.size, ids_t.size
    feats = np.concatenate((np.zeros((A * nt, X.shape[2]), X.dtype), X[0, ids_t]), 0).astype(np.float32)
    if nt != 0:
        pad = A * nt + nt
        k = np.arange(A * nt)
        g.ts = np.concatenate((g.ts, -np.ones(A * nt, np.int64), np.full(nt, t, np.int64)))
        g.det = np.concatenate((g.det, -np.ones(A * nt, np.int64), ids_t.astype(np.int64)))
        g.ass = np.concatenate((g.ass, -np.ones(pad, np.int64)))
        g.src = np.concatenate((g.src, active[k // nt] if A else np.zeros(0, np.int64), -np.ones(nt, np.int64)))
        g.dst = np.concatenate((g.dst, n_past + A * nt + (k % nt), -np.ones(nt, np.int64)))
        if g.label is not None:
            lab = np.zeros(pad, np.int64)
            tr_act = y_i[g.det[active], 1] if A else np.zeros(0, np.int64)
            tr_t = y_i[ids_t, 1]
            lab[A * nt:] = tr_t >= 0
            for j in range(nt):
                if tr_t[j] == -1:
                    continue
                for a in np.nonzero(tr_act == tr_t[j])[0]:
                    lab[a * nt + j] = 1
            g.label = np.concatenate((g.label, lab))
    return g, feats


def prune_graph(g, states, scores, t_st, t_ed, threshold=0.5):
    """``utils/graph.py:337-389``.  Returns (Graph', states', scores', keep_idx)."""
    assert t_st <= t_ed, "t_st must be lesser than or equal to t_ed!"
    idx = np.nonzero((g.ts >= t_st) & (g.ts <= t_ed))[0]
    if idx.size == 0:
        return g, states, scores, np.arange(g.n)
    i_st, i_ed = idx[0], idx[-1]
    ind = np.arange(g.n)
    keep = np.nonzero((scores[:, 1] >= threshold) | (g.ts != -1) | (ind < i_st) | (ind > i_ed))[0]
    return _gather_rows(g, keep), states[keep], scores[keep], keep


def _gather_rows(g, keep):
    new_of_old = -np.ones(g.n, np.int64)
    new_of_old[keep] = np.arange(keep.size)
    src = g.src[keep].copy(); dst = g.dst[keep].copy()
    e = src >= 0
    src[e] = new_of_old[src[e]]; dst[e] = new_of_old[dst[e]]
    assert (src[e] >= 0).all() and (dst[e] >= 0).all(), "dangling edge after row deletion"
    return Graph(g.ts[keep], g.det[keep], g.ass[keep], src, dst, None if g.label is None else g.label[keep])


def decode_tracks(g, states, scores, y_out, t_upto, ret_win_size, use_hungarian=False):
    """``utils/graph.py:392-539``.  Mutates and returns y_out [ND, 2] int64 like the reference.
    Returns (Graph', y_out, states', scores', keep_idx)."""
    g = g.copy()
    scores = np.asarray(scores, dtype=np.float32)
    p = scores[:, 1]
    if use_hungarian:
        g.ass = associate_hungarian(g, scores)
    else:
        g.ass = associate_greedy(g, p)

    # walk chains in detection-id order (:456-490)
    next_id = int(np.amax(y_out[:, 1])) + 1
    visited = np.zeros(y_out.shape[0], np.int64)
    row_of_det = {int(d): int(r) for r, d in enumerate(g.det) if d >= 0}
    for i in range(y_out.shape[0]):
        d = i
        n = row_of_det.get(d)
        if n is None:
            visited[d] = 1
            continue
        if g.ts[n] >= t_upto or p[n] < 0.5:
            visited[d] = 1
            continue
        if visited[d]:
            continue
        if y_out[d, 1] == -1:
            cur = next_id; next_id += 1
        else:
            cur = y_out[d, 1]
        while True:
            visited[d] = 1
            y_out[d, 1] = cur
            if g.ass[n] == -1:
                break
            if y_out[d, 0] >= t_upto and y_out[g.ass[n], 0] >= t_upto:
                break
            d = int(g.ass[n])
            n = row_of_det[d]

    # delete everything before t_upto (:492-512)
    cand = np.nonzero((g.ts < t_upto) & (g.ts != -1))[0]
    max_id = int(cand[-1]) + 1 if cand.size else 0
    delete = np.zeros(g.n, bool)
    delete[:max_id] = True
    for idx in range(max_id):
        if g.ts[idx] == -1:
            continue
        if g.ass[idx] == -1 and p[idx] >= 0.5 and g.ts[idx] >= t_upto - ret_win_size:
            delete[idx] = False
        else:
            e = np.nonzero(g.src == idx)[0]
            delete[e[e >= max_id]] = True
    keep = np.nonzero(~delete)[0]
    return _gather_rows(g, keep), y_out, states[keep], scores[keep], keep


# --------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------
def _segments(g):
    """past[d] = ascending edge rows with dst == d, future[d] = ascending edge rows with src == d."""
    e = np.nonzero(g.ts < 0)[0]
    past, fut = {}, {}
    for x in e:
        past.setdefault(int(g.dst[x]), []).append(int(x))
        fut.setdefault(int(g.src[x]), []).append(int(x))
    return past, fut


def create_targets(g):
    """``models/loss.py:8-44`` with labels = g.label and idx_node = all detection rows."""
    lab = g.label
    tgt = np.zeros_like(lab)
    dets = np.nonzero(g.ts >= 0)[0]
    tgt[dets] = lab[dets]
    past, fut = _segments(g)
    for d in dets:
        seg = np.asarray(past.get(int(d), []), np.int64)
        if seg.size:
            pos = np.nonzero(lab[seg])[0]
            if pos.size:
                tgt[seg[pos[-1]]] = 1           # latest positive past edge
        seg = np.asarray(fut.get(int(d), []), np.int64)
        if seg.size:
            pos = np.nonzero(lab[seg])[0]
            if pos.size:
                tgt[seg[pos[0]]] = 1            # earliest positive future edge
    return tgt


def _logsumexp(v):
    m = np.max(v)
    return m + np.log(np.sum(np.exp(v - m)))


def ce_loss(logits, targets, g):
    """``models/loss.py:81-115``: per detection, softmax-CE over its past and over its future
    incident edges (when the segment holds a positive target), divided by the segment size."""
    lg = np.asarray(logits, dtype=np.float64).reshape(-1)
    dets = np.nonzero(g.ts >= 0)[0]
    past, fut = _segments(g)
    loss = 0.0
    for d in dets:
        for seg, pick in ((past.get(int(d)), -1), (fut.get(int(d)), 0)):
            if not seg:
                continue
            seg = np.asarray(seg, np.int64)
            pos = np.nonzero(targets[seg])[0]
            if pos.size == 0:
                continue
            k = pos[pick]
            loss += (_logsumexp(lg[seg]) - lg[seg[k]]) / seg.size
    return np.float32(loss)


def focal_loss(p, targets, gamma=0, eps=1e-10):
    """``models/loss.py:57-74`` with alpha=None, size_average=True (``train.py:333-334``).
    p = scores[idx, 0] as passed by the driver, targets in {0, 1}."""
    p = np.asarray(p, dtype=np.float32); targets = np.asarray(targets).astype(np.int64)
    if p.size == 0:
        return np.float32(np.nan)
    pt = np.where(targets == 1, p, np.float32(1.0) - p).astype(np.float32)
    logpt = np.log(pt + np.float32(eps))
    ptv = np.exp(logpt)
    return np.float32(np.mean(-1.0 * (1 - ptv) ** gamma * logpt))
