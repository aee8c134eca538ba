"""Pins oracle/trackmpnn_oracle.py against outputs of the unmodified reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only.

Integer results (graph rows, edge lists, labels, targets, decoded track ids) must be
bit-exact.  Floating point: max-abs 1e-4 on logits / hidden states (BASELINE.json
north_star tolerance); observed error is ~1e-6.
"""
import numpy as np
import pytest

from oracle import trackmpnn_oracle as O
from golden_util import Golden, golden_names, assert_graph_equal, assert_grads_close, attention_keep_matrix

TOL = 1e-4


def _fwd(gold, params, x, h_in, g, training=False, step=0):
    m = gold.meta
    keep = (lambda gi, k, n: attention_keep_matrix(m['seed'], step, gi, k, n)) if m.get('nattheads', 0) and training else None
    return O.forward(params, x, h_in, g, features=m['features'], ncategories=m['ncategories'], nhidden=64,
                     msg_type=m['msg_type'], training=training, attention_keep=keep)


@pytest.mark.parametrize('name', golden_names('infer'))
def test_infer_step_locked(name):
    """Every op of the inference loop, each fed the reference's own inputs for that step."""
    gold = Golden(name)
    m = gold.meta
    params = gold.params()
    y_out = gold.y[0].astype(np.int64); y_out[:, 1] = -1
    h_prev = None
    for s in range(gold.n_steps):
        op = int(gold.get(s, 'op'))
        g_ref = gold.graph(s)
        if op in (0, 2):
            t_st = 0 if op == 0 else int(gold.get(s, 't'))
            g, feats, t1, tN = O.initialize_graph(gold.X, gold.y, t_st=t_st, mode='test')
            assert tN == int(gold.z['t_end'])
            h_prev = None
        else:
            g_prev = gold.graph(s - 1, 'dec_') if gold.has(s - 1, 'dec_y_pred') else gold.graph(s - 1)
            sc_prev = gold.get(s - 1, 'dec_scores') if gold.has(s - 1, 'dec_scores') else gold.get(s - 1, 'scores')
            g, feats = O.update_graph(g_prev, sc_prev, gold.X, gold.y, int(gold.get(s, 't')),
                                      use_hungarian=m['hungarian'], mode='test')
        assert_graph_equal(g, g_ref, f'{name} step {s} update')
        np.testing.assert_array_equal(feats, gold.get(s, 'feats'))
        scores, logits, h = _fwd(gold, params, feats, h_prev, g)
        np.testing.assert_allclose(logits, gold.get(s, 'logits'), atol=TOL, rtol=0)
        np.testing.assert_allclose(h, gold.get(s, 'h'), atol=TOL, rtol=0)
        ref_scores = gold.get(s, 'scores')
        if m['tp_classifier']:
            np.testing.assert_allclose(scores[:, 0], ref_scores[:, 1], atol=TOL, rtol=0)
        sc, hh = ref_scores, gold.get(s, 'h')
        if gold.has(s, 'prune_y_pred'):
            t_lo, t_hi = gold.get(s, 'prune_t')
            g, hh, sc, keep = O.prune_graph(g_ref, hh, sc, int(t_lo), int(t_hi), 0.5)
            assert_graph_equal(g, gold.graph(s, 'prune_'), f'{name} step {s} prune')
            np.testing.assert_array_equal(hh, gold.get(s, 'prune_h'))
            np.testing.assert_array_equal(sc, gold.get(s, 'prune_scores'))
            g_ref = g
        if gold.has(s, 'dec_y_pred'):
            g2, y_out, h2, sc2, keep = O.decode_tracks(g_ref, hh, sc, y_out, int(gold.get(s, 't_upto')),
                                                       m['ret_win_size'], use_hungarian=m['hungarian'])
            assert_graph_equal(g2, gold.graph(s, 'dec_'), f'{name} step {s} decode')
            np.testing.assert_array_equal(y_out, gold.get(s, 'y_out'))
            np.testing.assert_array_equal(h2, gold.get(s, 'dec_h'))
            np.testing.assert_array_equal(sc2, gold.get(s, 'dec_scores'))
            h_prev = gold.get(s, 'dec_h')
        else:
            h_prev = gold.get(s, 'h')


@pytest.mark.parametrize('name', golden_names('infer'))
def test_infer_free_running(name):
    """The whole loop on the oracle's own state: final tracks bit-exact, states within tolerance."""
    gold = Golden(name)
    m = gold.meta
    params = gold.params()
    y_out = gold.y[0].astype(np.int64); y_out[:, 1] = -1
    g, feats, t_st, t_end = O.initialize_graph(gold.X, gold.y, 0, 'test')

    def fix(scores, g):
        sc = np.concatenate((1 - scores, scores), 1).astype(np.float32)
        if not m['tp_classifier']:
            sc[g.ts >= 0] = (0.0, 1.0)
        return sc

    scores, logits, h = _fwd(gold, params, feats, None, g)
    scores = fix(scores, g)
    s = 0
    t_skip = t_st
    for t_cur in range(t_st, t_end):
        if t_cur < t_skip:
            continue
        s += 1
        if feats.shape[0] == 0 and h.shape[0] == 0:
            r = O.initialize_graph(gold.X, gold.y, t_cur, 'test')
            if r is None:
                break
            g, feats, t_skip, _ = r
            h = None
        else:
            g, feats = O.update_graph(g, scores, gold.X, gold.y, t_cur, use_hungarian=m['hungarian'], mode='test')
        h_before = h
        scores, logits, h = _fwd(gold, params, feats, h, g)
        scores = fix(scores, g)
        np.testing.assert_allclose(h, gold.get(s, 'h'), atol=TOL, rtol=0)
        if gold.has(s, 'att'):  # dense attention of every group / head at this step
            _, _, _, att = O.forward(params, feats, h_before, g, features=m['features'], ncategories=m['ncategories'],
                                     msg_type=m['msg_type'], return_attention=True)
            want = gold.get(s, 'att')
            for gi in range(want.shape[0]):
                for k in range(want.shape[1]):
                    np.testing.assert_allclose(O.dense_attention(att[gi][k], g.n), want[gi, k], atol=1e-5, rtol=0)
        if gold.has(s, 'prune_y_pred'):
            t_lo, t_hi = gold.get(s, 'prune_t')
            g, h, scores, _ = O.prune_graph(g, h, scores, int(t_lo), int(t_hi), 0.5)
        t_upto = t_end if t_cur == t_end - 1 else t_cur - m['cur_win_size'] + 2
        g, y_out, h, scores, _ = O.decode_tracks(g, h, scores, y_out, t_upto, m['ret_win_size'],
                                                 use_hungarian=m['hungarian'])
        np.testing.assert_array_equal(y_out, gold.get(s, 'y_out'))
    assert s + 1 == gold.n_steps


@pytest.mark.parametrize('name', golden_names('train'))
def test_train_chunk(name):
    """Teacher-forced graph growth, train-mode BatchNorm, targets, CE and BCE per MP step."""
    gold = Golden(name)
    m = gold.meta
    params = gold.params()
    g, feats, t_st, t_end = O.initialize_graph(gold.X, gold.y, 0, 'train')
    h = None
    loss_c = loss_f = 0.0
    s = 0
    for t_cur in [None] + list(range(t_st, t_end)):
        if t_cur is not None:
            s += 1
            g, feats = O.update_graph(g, gold.get(s - 1, 'scores'), gold.X, gold.y, t_cur, mode='train')
        assert_graph_equal(g, gold.graph(s), f'{name} step {s}')
        np.testing.assert_array_equal(feats, gold.get(s, 'feats'))
        scores, logits, h = _fwd(gold, params, feats, h, g, training=True, step=s)
        np.testing.assert_allclose(logits, gold.get(s, 'logits'), atol=TOL, rtol=0)
        np.testing.assert_allclose(h, gold.get(s, 'h'), atol=TOL, rtol=0)
        tg = O.create_targets(g)
        np.testing.assert_array_equal(tg, gold.get(s, 'targets'))
        lc = O.ce_loss(logits, tg, g)
        np.testing.assert_allclose(lc, gold.get(s, 'loss_c'), rtol=1e-4, atol=1e-5)
        e = g.ts < 0; d = g.ts >= 0
        lf = O.focal_loss(scores[e, 0], tg[e])
        if m['tp_classifier']:
            lf = lf + O.focal_loss(scores[d, 0], tg[d])
        np.testing.assert_allclose(lf, gold.get(s, 'loss_f'), rtol=1e-4, atol=1e-5)
        loss_c += float(lc); loss_f += float(lf)
    assert s + 1 == gold.n_steps
    np.testing.assert_allclose(loss_c + loss_f, float(gold.z['loss']), rtol=1e-4)
    for k in gold.z.files:
        if k.startswith('w_after/') and 'num_batches' not in k:
            np.testing.assert_allclose(params[k[len('w_after/'):]], gold.z[k], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize('name', golden_names('train'))
def test_train_gradients(name):
    """oracle/train_ref.py (edge-list forward in torch fp32 + autograd) against the reference's own
    loss.backward(): loss and every parameter gradient of a BPTT chunk."""
    from oracle import train_ref as T
    gold = Golden(name)
    m = gold.meta
    out = T.train_chunk(gold.params(), gold.X, gold.y, features=m['features'], ncategories=m['ncategories'],
                        msg_type=m['msg_type'], tp_classifier=m['tp_classifier'],
                        attention_keep=(lambda s, gi, k, n: attention_keep_matrix(m['seed'], s, gi, k, n))
                        if m.get('nattheads', 0) else None)
    np.testing.assert_allclose(out['loss'], float(gold.z['loss']), rtol=1e-4)
    assert len(out['graphs']) == gold.n_steps
    assert_grads_close(out['grads'], gold)
