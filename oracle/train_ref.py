"""CPU oracle for the TRAINING side of the hot path: one BPTT chunk, loss and every parameter
gradient.  TEST INFRASTRUCTURE ONLY (same rules as trackmpnn_oracle.py: imported by tests/,
smoke() and bench.py's cpu legs, never by the product).

The forward is the edge-list restatement of ``models/track_mpnn.py:54-75`` +
``models/layers.py:84-116`` written with plain torch fp32 ops so that torch autograd supplies
the backward pass the reference gets from ``loss.backward()`` (``train.py:65-134``); the losses
restate ``models/loss.py:8-115``; graph growth comes from ``trackmpnn_oracle.update_graph``
(teacher forcing, ``utils/graph.py:229-245, 271-327``).

Parity pin: ``tests/test_oracle_golden.py::test_train_gradients`` checks loss and all
gradients against ``tests/golden/train_*.npz`` (``g/<parameter name>``), which
``tests/golden/make_golden.py`` produced by running the unmodified reference.
"""
import numpy as np
import torch

from . import trackmpnn_oracle as O

H = 64


def _gru(x, h, w_ih, w_hh, b_ih, b_hh):
    gi = x @ w_ih.t() + b_ih
    gh = h @ w_hh.t() + b_hh
    r = torch.sigmoid(gi[:, :H] + gh[:, :H])
    z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    return (1 - z) * n + z * h


def input_transform_train(p, gi, x_det, n_edge_rows, eps=1e-5):
    """Linear -> train-mode BatchNorm over (detection rows + n_edge_rows copies of b1) -> ReLU ->
    Linear, detection rows only.  Differentiable; returns (out, mu, var)."""
    w1, b1 = p[f'input_transforms.{gi}.0.weight'], p[f'input_transforms.{gi}.0.bias']
    gam, bet = p[f'input_transforms.{gi}.1.weight'], p[f'input_transforms.{gi}.1.bias']
    w2, b2 = p[f'input_transforms.{gi}.3.weight'], p[f'input_transforms.{gi}.3.bias']
    a = x_det @ w1.t() + b1
    n = a.shape[0] + n_edge_rows
    if n <= 1:
        raise ValueError('Expected more than 1 value per channel when training')
    mu = (a.sum(0) + n_edge_rows * b1) / n
    var = (((a - mu) ** 2).sum(0) + n_edge_rows * (b1 - mu) ** 2) / n
    bn = (a - mu) / torch.sqrt(var + eps) * gam + bet
    return torch.relu(bn) @ w2.t() + b2, mu, var


def gat_support_train(p, gi, k, h, g, keep):
    """One attention head in train mode (``models/layers.py:26-43``) in differentiable torch ops on the edge list:
    per-edge score, softmax over each detection's incident edges, dropout through the pinned ``keep`` [N, N]
    decisions (x 2), signed weighted sum.  Returns the support of every row ([N, H], zero off the detection rows)."""
    w_att, a = p[f'factor_grus.{gi}.gat.{k}.W_att'], p[f'factor_grus.{gi}.gat.{k}.a']
    e_rows = np.nonzero(g.ts < 0)[0]
    hatt = h @ w_att
    pre = (torch.abs(hatt[torch.from_numpy(g.src[e_rows])] - hatt[torch.from_numpy(g.dst[e_rows])]) @ a).reshape(-1)
    ev = torch.zeros(g.n, dtype=torch.float32).index_put((torch.from_numpy(e_rows),), torch.nn.functional.leaky_relu(pre, 0.2))
    past, fut = O._segments(g)
    rows, sup = [], []
    for r in np.nonzero(g.ts >= 0)[0]:
        inc = np.asarray(sorted(past.get(int(r), []) + fut.get(int(r), [])), np.int64)
        if inc.size == 0:
            continue
        w = torch.softmax(ev[torch.from_numpy(inc)], 0) * torch.from_numpy(keep[r, inc].astype(np.float32) * 2.0)
        sign = torch.from_numpy(np.where(g.src[inc] == r, 1.0, -1.0).astype(np.float32))
        rows.append(int(r))
        sup.append(((w * sign)[:, None] * h[torch.from_numpy(inc)]).sum(0))
    out = torch.zeros_like(h)
    return out.index_put((torch.tensor(rows, dtype=torch.int64),), torch.stack(sup)) if rows else out


def forward_train(p, x_new, h_in, g, groups, msg_type, attention_keep=None):
    """One train-mode ``TrackMPNN.forward`` on graph ``g`` (oracle Graph).  p: dict of torch
    parameters (requires_grad).  ``attention_keep(group, head, n)``: pinned dropout decisions when the parameters hold
    attention heads.  Returns scores[N,1], logits[N,1], h_out[N, G*H]."""
    n_tot = g.n
    n_new = x_new.shape[0]
    n_old = n_tot - n_new
    is_det = torch.from_numpy(g.ts >= 0)
    e = torch.from_numpy(np.nonzero(g.ts < 0)[0])
    d = torch.from_numpy(np.nonzero(g.ts >= 0)[0])
    src = torch.from_numpy(g.src[g.ts < 0]); dst = torch.from_numpy(g.dst[g.ts < 0])
    outs = []
    for gi, (a, b) in enumerate(groups):
        if n_new > 0:
            new_det = is_det[n_old:]
            n_edge_new = int((~new_det).sum())
            out_det, _, _ = input_transform_train(p, gi, x_new[new_det][:, a:b], n_edge_new)
            upd = torch.zeros((n_new, H), dtype=torch.float32)
            upd = upd.index_put((torch.nonzero(new_det)[:, 0],), out_det)
            h = upd if h_in is None else torch.cat((h_in[:, gi * H:(gi + 1) * H], upd), 0)
        else:
            h = h_in[:, gi * H:(gi + 1) * H]
        if msg_type == 'concat':
            xs = torch.cat((h[src], h[dst]), 1)
        else:
            xs = h[src] - h[dst]
        nheads = O.gat_heads(p, gi)
        if nheads:   # models/layers.py:105-112: heads summed, divided by their number
            agg = sum(gat_support_train(p, gi, k, h, g, attention_keep(gi, k, n_tot)) for k in range(nheads))[d] / float(nheads)
        else:
            agg = torch.zeros_like(h).index_add(0, src, h[e]).index_add(0, dst, -h[e])[d]
        pre = f'factor_grus.{gi}.'
        he = _gru(xs, h[e], p[pre + 'edge_gru.weight_ih'], p[pre + 'edge_gru.weight_hh'],
                  p[pre + 'edge_gru.bias_ih'], p[pre + 'edge_gru.bias_hh'])
        hd = _gru(agg, h[d], p[pre + 'node_gru.weight_ih'], p[pre + 'node_gru.weight_hh'],
                  p[pre + 'node_gru.bias_ih'], p[pre + 'node_gru.bias_hh'])
        hn = torch.zeros_like(h).index_put((e,), he).index_put((d,), hd)
        outs.append(hn)
    h_out = torch.cat(outs, 1)
    ln = h_out @ p['output_transform_node.weight'][0] + p['output_transform_node.bias'][0]
    le = h_out @ p['output_transform_edge.weight'][0] + p['output_transform_edge.bias'][0]
    logits = torch.where(is_det, ln, le)[:, None]
    return torch.sigmoid(logits), logits, h_out


def ce_loss(logits, targets, g):
    """``models/loss.py:81-115`` (differentiable)."""
    lg = logits.reshape(-1)
    past, fut = O._segments(g)
    loss = torch.zeros((), dtype=torch.float32)
    for d in np.nonzero(g.ts >= 0)[0]:
        for seg, pick in ((past.get(int(d)), -1), (fut.get(int(d)), 0)):
            if not seg:
                continue
            seg = np.asarray(seg, np.int64)
            pos = np.nonzero(targets[seg])[0]
            if pos.size == 0:
                continue
            s = lg[torch.from_numpy(seg)]
            loss = loss + (torch.logsumexp(s, 0) - s[int(pos[pick])]) / float(seg.size)
    return loss


def focal_loss(p1, targets, eps=1e-10):
    """``models/loss.py:57-74`` with gamma=0, alpha=None: mean(-log(p_t + eps))."""
    t = torch.from_numpy(np.asarray(targets).astype(np.int64))
    pt = torch.where(t == 1, p1, 1 - p1)
    return (-torch.log(pt + eps)).mean()


def step_losses(scores, logits, g, tp_classifier):
    """The per-step loss terms of ``train.py:70-81`` -> (loss_c, loss_f, targets)."""
    tg = O.create_targets(g)
    e = np.nonzero(g.ts < 0)[0]; d = np.nonzero(g.ts >= 0)[0]
    lc = ce_loss(logits, tg, g)
    lf = focal_loss(scores[torch.from_numpy(e), 0], tg[e])
    if tp_classifier:
        lf = focal_loss(scores[torch.from_numpy(d), 0], tg[d]) + lf
    return lc, lf, tg


def train_chunk(params, X, y, features='2d', ncategories=3, msg_type='diff', tp_classifier=True, attention_keep=None):
    """One chunk of ``train.py:65-134``.  params: dict name -> numpy array (state_dict layout).
    ``attention_keep(step, group, head, n)``: pinned dropout decisions of the attention heads, if any.
    Returns dict(loss, loss_c, loss_f, grads {name: ndarray}, graphs [Graph per step], logits, h)."""
    groups = O.feature_groups(features, ncategories)
    p = {k: torch.tensor(np.asarray(v), dtype=torch.float32, requires_grad=True) for k, v in params.items()
         if 'running' not in k and 'num_batches' not in k}
    g, feats, t_st, t_end = O.initialize_graph(X, y, 0, 'train')
    h = None
    loss_c = loss_f = 0.0
    graphs, all_logits, all_h = [], [], []
    for step, t_cur in enumerate([None] + list(range(t_st, t_end))):
        if t_cur is not None:
            # teacher forcing: the graph growth does not read the scores (utils/graph.py:229-245, 271-274)
            g, feats = O.update_graph(g, np.zeros((g.n, 2), np.float32), X, y, t_cur, mode='train')
        keep = None if attention_keep is None else (lambda gi, k, n, step=step: attention_keep(step, gi, k, n))
        scores, logits, h = forward_train(p, torch.from_numpy(np.ascontiguousarray(feats, dtype=np.float32)), h, g,
                                          groups, msg_type, keep)
        lc, lf, _ = step_losses(scores, logits, g, tp_classifier)
        loss_c = loss_c + lc; loss_f = loss_f + lf
        graphs.append(g.copy()); all_logits.append(logits.detach().numpy()); all_h.append(h.detach().numpy())
    loss = loss_c + loss_f
    loss.backward()
    grads = {k: (np.zeros(tuple(v.shape), np.float32) if v.grad is None else v.grad.numpy()) for k, v in p.items()}
    return dict(loss=float(loss), loss_c=float(loss_c), loss_f=float(loss_f), grads=grads, graphs=graphs,
                logits=all_logits, h=all_h)
