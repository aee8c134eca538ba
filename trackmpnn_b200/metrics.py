"""CLEAR-MOT / identity metrics of decoded tracks: the step behind the path (reference ``utils/metrics.py:7-61``).

The reference hands its decoded tracks to the third-party package **motmetrics** (py-motmetrics, un-pinned in the
reference's ``Pipfile``; ``mm.MOTAccumulator``, ``mm.distances.iou_matrix``, ``mm.metrics.motchallenge_metrics``), which
is not installed here and not vendored by the reference.  This module restates that package's published algorithm for
exactly the calls the reference makes, with the reference's own function names and argument meaning:

* ``create_mot_accumulator(bbox_pred, bbox_gt, y_out, y_gt)``  (``utils/metrics.py:7-44``): per frame the 1 - IoU
  distance matrix between ground-truth and predicted boxes (pairs with IoU < 0.5 are not matchable) and one CLEAR-MOT
  update: correspondences of the previous frame are kept while they stay valid, the rest is matched at minimum total
  distance; a matched object whose last hypothesis was a different one counts as an identity SWITCH; unmatched objects
  are MISSes, unmatched hypotheses FALSE POSITIVES (Bernardin & Stiefelhagen 2008, as implemented by motmetrics).
* ``calc_mot_metrics(accs)``  (``utils/metrics.py:47-61``): the 15 MOT-challenge figures over all accumulators together
  ("OVERALL" row of ``compute_many``): idf1 idp idr recall precision num_unique_objects mostly_tracked partially_tracked
  mostly_lost num_false_positives num_misses num_switches num_fragmentations mota motp.

Host code on purpose: it runs once per validation pass on a few thousand boxes; nothing here is on the hot path.
Parity: motmetrics cannot be run in this image, so the tests pin this module to hand-computed cases (the reference's
own outputs cannot be generated here).  Where several assignments have the same total distance the choice between them
is the solver's (motmetrics itself delegates to whichever LAP solver is installed).
"""
import numpy as np

MOTCHALLENGE_METRICS = ['idf1', 'idp', 'idr', 'recall', 'precision', 'num_unique_objects', 'mostly_tracked',
                        'partially_tracked', 'mostly_lost', 'num_false_positives', 'num_misses', 'num_switches',
                        'num_fragmentations', 'mota', 'motp']


def iou_distance_matrix(objs, hyps, max_iou=0.5):
    """``mm.distances.iou_matrix``: boxes as [x, y, w, h]; distance = 1 - IoU, NaN where it exceeds ``max_iou``."""
    objs = np.asarray(objs, np.float64).reshape(-1, 4)
    hyps = np.asarray(hyps, np.float64).reshape(-1, 4)
    if objs.shape[0] == 0 or hyps.shape[0] == 0:
        return np.empty((objs.shape[0], hyps.shape[0]))
    ox0, oy0, ox1, oy1 = objs[:, 0:1], objs[:, 1:2], objs[:, 0:1] + objs[:, 2:3], objs[:, 1:2] + objs[:, 3:4]
    hx0, hy0, hx1, hy1 = hyps[:, 0], hyps[:, 1], hyps[:, 0] + hyps[:, 2], hyps[:, 1] + hyps[:, 3]
    iw = np.clip(np.minimum(ox1, hx1) - np.maximum(ox0, hx0), 0, None)
    ih = np.clip(np.minimum(oy1, hy1) - np.maximum(oy0, hy0), 0, None)
    inter = iw * ih
    union = objs[:, 2:3] * objs[:, 3:4] + hyps[:, 2] * hyps[:, 3] - inter
    with np.errstate(divide='ignore', invalid='ignore'):
        d = 1.0 - np.where(union > 0, inter / union, 0.0)
    d[d > max_iou] = np.nan
    return d


def linear_sum_assignment(cost):
    """Minimum-cost assignment of a rectangular matrix (shortest augmenting paths with dual variables; every entry finite).
    Returns (row indices, column indices) of min(nr, nc) pairs, rows ascending."""
    cost = np.asarray(cost, np.float64)
    transposed = cost.shape[0] > cost.shape[1]
    if transposed:
        cost = cost.T
    nr, nc = cost.shape
    u, v = np.zeros(nr), np.zeros(nc)
    col4row = -np.ones(nr, np.int64)
    row4col = -np.ones(nc, np.int64)
    for cur in range(nr):
        shortest = np.full(nc, np.inf)
        path = -np.ones(nc, np.int64)
        done_cols = np.zeros(nc, bool)
        done_rows = []
        i, min_val, sink = cur, 0.0, -1
        while sink < 0:
            done_rows.append(i)
            red = min_val + cost[i] - u[i] - v
            better = (~done_cols) & (red < shortest)
            shortest[better] = red[better]
            path[better] = i
            cand = np.where(done_cols, np.inf, shortest)
            j = int(np.argmin(cand))
            # among equally short columns prefer an unassigned one (ends the search)
            ties = np.nonzero((cand == cand[j]) & (row4col < 0) & ~done_cols)[0]
            if ties.size:
                j = int(ties[0])
            min_val = cand[j]
            if not np.isfinite(min_val):
                raise ValueError('cost matrix is infeasible')
            done_cols[j] = True
            if row4col[j] < 0:
                sink = j
            else:
                i = int(row4col[j])
        u[cur] += min_val
        for r in done_rows[1:]:
            u[r] += min_val - shortest[col4row[r]]
        v[done_cols] -= min_val - shortest[done_cols]
        j = sink
        while True:
            i = int(path[j])
            row4col[j] = i
            col4row[i], j = j, col4row[i]
            if i == cur:
                break
    rows = np.arange(nr)
    if transposed:
        order = np.argsort(col4row)
        return col4row[order], rows[order]
    return rows, col4row


class MOTAccumulator:
    """Event log of one sequence (``mm.MOTAccumulator`` with auto-incremented or explicit frame ids)."""

    def __init__(self, max_switch_time=float('inf')):
        self.max_switch_time = max_switch_time
        self.events = []        # (frame, type, object id or None, hypothesis id or None, distance or nan)
        self.raw = []           # (frame, object id, hypothesis id, distance): every valid pair, for the identity metrics
        self.obj_frames = {}    # object id -> frames it was present in
        self.hyp_frames = {}
        self.m = {}             # object id -> hypothesis id of its last match
        self.last_occurrence = {}
        self._frame = -1

    def update(self, oids, hids, dists, frameid=None):
        oids, hids = list(oids), list(hids)
        dists = np.asarray(dists, np.float64).reshape(len(oids), len(hids))
        self._frame = self._frame + 1 if frameid is None else frameid
        f = self._frame
        for o in oids:
            self.obj_frames.setdefault(o, []).append(f)
        for h in hids:
            self.hyp_frames.setdefault(h, []).append(f)
        valid = np.isfinite(dists)
        for i, j in zip(*np.nonzero(valid)):
            self.raw.append((f, oids[i], hids[j], float(dists[i, j])))
        o_left, h_left = set(range(len(oids))), set(range(len(hids)))
        # 1. correspondences of the previous frames that are still valid stay (no re-assignment, no switch)
        for i, o in enumerate(oids):
            if o in self.m and self.m[o] in hids:
                j = hids.index(self.m[o])
                if j in h_left and valid[i, j]:
                    self._event(f, 'MATCH', o, hids[j], dists[i, j])
                    o_left.discard(i); h_left.discard(j)
                    self.last_occurrence[o] = f
        # 2. minimum total distance over the rest; pairs that are not matchable cost more than any set of valid pairs
        oi, hj = sorted(o_left), sorted(h_left)
        if oi and hj:
            sub = dists[np.ix_(oi, hj)]
            big = 1e6
            rows, cols = linear_sum_assignment(np.where(np.isfinite(sub), sub, big))
            for a, b in zip(rows, cols):
                if not np.isfinite(sub[a, b]):
                    continue
                i, j = oi[a], hj[b]
                o, h = oids[i], hids[j]
                switch = (o in self.m and self.m[o] != h
                          and abs(f - self.last_occurrence.get(o, f)) <= self.max_switch_time)
                self._event(f, 'SWITCH' if switch else 'MATCH', o, h, sub[a, b])
                o_left.discard(i); h_left.discard(j)
                self.m[o] = h
                self.last_occurrence[o] = f
        # 3. what is left: misses and false positives
        for i in sorted(o_left):
            self._event(f, 'MISS', oids[i], None, np.nan)
        for j in sorted(h_left):
            self._event(f, 'FP', None, hids[j], np.nan)

    def _event(self, f, kind, o, h, d):
        self.events.append((f, kind, o, h, float(d)))
        if kind in ('MATCH', 'SWITCH'):
            self.m[o] = h


def create_mot_accumulator(bbox_pred, bbox_gt, y_out, y_gt):
    """Reference ``utils/metrics.py:7-44``.  bbox_* [N, (cat_id, alpha, x1, y1, x2, y2, ...)], y_* [N, (frame, track id)]
    (rows with track id < 0 are ignored).  Returns the accumulator of the sequence."""
    y_out, y_gt = np.asarray(y_out), np.asarray(y_gt)
    bbox_pred, bbox_gt = np.asarray(bbox_pred, np.float64), np.asarray(bbox_gt, np.float64)
    acc = MOTAccumulator()
    if y_out.shape[0] == 0 and y_gt.shape[0] == 0:
        return acc
    times = np.concatenate((y_out[:, 0], y_gt[:, 0]))
    for t in range(int(times.min()), int(times.max()) + 1):
        oids = np.nonzero((y_gt[:, 0] == t) & (y_gt[:, 1] >= 0))[0]
        hids = np.nonzero((y_out[:, 0] == t) & (y_out[:, 1] >= 0))[0]
        bo = bbox_gt[oids, 2:6].copy(); bo[:, 2:] -= bo[:, :2]      # (x1, y1, x2, y2) -> (x, y, w, h)
        bh = bbox_pred[hids, 2:6].copy(); bh[:, 2:] -= bh[:, :2]
        acc.update(y_gt[oids, 1].astype(np.float32).tolist(), y_out[hids, 1].astype(np.float32).tolist(),
                   iou_distance_matrix(bo, bh, max_iou=0.5), frameid=t)
    return acc


def _identity_counts(accs):
    """IDTP of the global minimum-cost matching between ground-truth and predicted trajectories (Ristani et al. 2016 as
    implemented by motmetrics' ``id_global_assignment``): cost = false negatives + false positives a pairing implies."""
    objs, hyps, pair = [], [], {}
    n_obj = n_hyp = 0
    for k, acc in enumerate(accs):
        for o, fr in acc.obj_frames.items():
            objs.append(((k, o), len(fr))); n_obj += len(fr)
        for h, fr in acc.hyp_frames.items():
            hyps.append(((k, h), len(fr))); n_hyp += len(fr)
        for _, o, h, _ in acc.raw:
            pair[((k, o), (k, h))] = pair.get(((k, o), (k, h)), 0) + 1
    no, nh = len(objs), len(hyps)
    if no == 0 or nh == 0:
        return 0, n_obj, n_hyp
    oi = {key: i for i, (key, _) in enumerate(objs)}
    hi = {key: j for j, (key, _) in enumerate(hyps)}
    big = float(n_obj + n_hyp + 1)
    fn = np.full((no + nh, no + nh), big)
    fp = np.full((no + nh, no + nh), big)
    oc = np.array([c for _, c in objs], np.float64)
    hc = np.array([c for _, c in hyps], np.float64)
    fn[:no, :nh] = oc[:, None]; fp[:no, :nh] = hc[None, :]
    for (ko, kh), c in pair.items():
        fn[oi[ko], hi[kh]] -= c
        fp[oi[ko], hi[kh]] -= c
    # an object left alone loses all its detections, a hypothesis left alone is all false positives
    fn[np.arange(no), nh + np.arange(no)] = oc; fp[np.arange(no), nh + np.arange(no)] = 0.0
    fp[no + np.arange(nh), np.arange(nh)] = hc; fn[no + np.arange(nh), np.arange(nh)] = 0.0
    fn[no:, nh:] = 0.0; fp[no:, nh:] = 0.0
    rows, cols = linear_sum_assignment(fn + fp)
    idfn = fn[rows, cols].sum()
    return int(round(n_obj - idfn)), n_obj, n_hyp


def calc_mot_metrics(accs):
    """Reference ``utils/metrics.py:47-61``: the MOT-challenge metrics over all accumulators together (dict)."""
    accs = list(accs)
    n_obj = sum(len(fr) for a in accs for fr in a.obj_frames.values())
    n_pred = sum(len(fr) for a in accs for fr in a.hyp_frames.values())
    cnt = dict(MATCH=0, SWITCH=0, MISS=0, FP=0)
    dist_sum = 0.0
    mt = pt = ml = frag = uniq = 0
    for a in accs:
        per_obj = {}
        for f, kind, o, h, d in a.events:
            cnt[kind] += 1
            if kind in ('MATCH', 'SWITCH'):
                dist_sum += d
            if o is not None:
                per_obj.setdefault(o, []).append(kind)
        for o, kinds in per_obj.items():
            uniq += 1
            tracked = sum(k != 'MISS' for k in kinds)
            ratio = tracked / len(kinds)
            mt += ratio >= 0.8
            ml += ratio < 0.2
            pt += 0.2 <= ratio < 0.8
            if tracked:   # inside the span first..last tracked frame: transitions tracked -> missed
                first = next(i for i, k in enumerate(kinds) if k != 'MISS')
                last = len(kinds) - 1 - next(i for i, k in enumerate(reversed(kinds)) if k != 'MISS')
                span = [k == 'MISS' for k in kinds[first:last + 1]]
                frag += sum(1 for x, y in zip(span[:-1], span[1:]) if (not x) and y)
    n_det = cnt['MATCH'] + cnt['SWITCH']
    idtp, _, _ = _identity_counts(accs)
    nan = float('nan')
    div = lambda a, b: a / b if b else nan
    return {
        'idf1': div(2.0 * idtp, n_obj + n_pred), 'idp': div(idtp, n_pred), 'idr': div(idtp, n_obj),
        'recall': div(n_det, n_obj), 'precision': div(n_det, n_det + cnt['FP']),
        'num_unique_objects': uniq, 'mostly_tracked': int(mt), 'partially_tracked': int(pt), 'mostly_lost': int(ml),
        'num_false_positives': cnt['FP'], 'num_misses': cnt['MISS'], 'num_switches': cnt['SWITCH'],
        'num_fragmentations': int(frag), 'mota': 1.0 - div(cnt['MISS'] + cnt['SWITCH'] + cnt['FP'], n_obj) if n_obj else nan,
        'motp': div(dist_sum, n_det),
    }


def compute_map(*args, **kwargs):
    """Reference ``utils/metrics.py:64-228`` (VOC-style detection mAP; itself broken on numpy >= 1.24: ``np.str``).  Detection
    quality is a property of the detector in front of the path, not of the tracker: out of scope.
    The name exists so that ``from utils.metrics import create_mot_accumulator, calc_mot_metrics, compute_map``
    (train.py:17) resolves under the overlay."""
    raise NotImplementedError('compute_map (detection mAP) is outside the tracking hot path; use the reference\'s own '
                              'utils/metrics.py for it')
