"""The reference's inference driver loop (``infer.py:48-87``) on the oracle.  TEST INFRASTRUCTURE:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.

``run_infer`` returns the decoded tracks plus the work counters the bench metric is built
from: edge_updates = sum over forward calls of the number of edge rows in the graph
(SURVEY.md section 8d), frames = iterations of the frame loop that ran a forward."""
import numpy as np

from . import trackmpnn_oracle as O


def run_infer(params, X, y, features='2d', ncategories=3, msg_type='diff', cur_win_size=5, ret_win_size=0,
              use_hungarian=False, tp_classifier=True, max_frames=None, record_margin=False, keep_state=False):
    X = np.asarray(X, np.float32); y = np.asarray(y, np.float32)
    if X.ndim == 2:
        X = X[None]; y = y[None]
    kw = dict(features=features, ncategories=ncategories, nhidden=64, msg_type=msg_type)
    y_out = y[0].astype(np.int64); y_out[:, 1] = -1
    stats = dict(edge_updates=0, det_updates=0, frames=0, forwards=0, margin=np.inf)

    def fix(scores, g):
        sc = np.concatenate((1 - scores, scores), 1).astype(np.float32)
        if not tp_classifier:
            sc[g.ts >= 0] = (0.0, 1.0)
        if record_margin and sc.shape[0]:
            stats['margin'] = min(stats['margin'], float(np.min(np.abs(sc[:, 1] - 0.5))))
        return sc

    def fwd(feats, h, g):
        scores, logits, h = O.forward(params, feats, h, g, **kw)
        ne = int((g.ts < 0).sum())
        stats['edge_updates'] += ne
        stats['det_updates'] += g.n - ne
        stats['forwards'] += 1
        return fix(scores, g), h

    r = O.initialize_graph(X, y, 0, 'test')
    if r is None:
        return y_out, stats
    g, feats, t_st, t_end = r
    scores, h = fwd(feats, None, g)
    t_skip = t_st
    for t_cur in range(t_st, t_end):
        if t_cur < t_skip:
            continue
        if max_frames is not None and stats['frames'] >= max_frames:
            break
        if feats.shape[0] == 0 and h.shape[0] == 0:
            r = O.initialize_graph(X, y, t_cur, 'test')
            if r is None:
                break
            g, feats, t_skip, _ = r
            h = None
        else:
            g, feats = O.update_graph(g, scores, X, y, t_cur, use_hungarian=use_hungarian, mode='test')
        scores, h = fwd(feats, h, g)
        t_upto = t_end if t_cur == t_end - 1 else t_cur - cur_win_size + 2
        g, y_out, h, scores, _ = O.decode_tracks(g, h, scores, y_out, t_upto, ret_win_size, use_hungarian=use_hungarian)
        stats['frames'] += 1
    if keep_state:  # the window graph, states and scores the loop stopped with (tests compare them with the engine's)
        stats['state'] = dict(g=g, h=h, scores=scores)
    return y_out, stats
