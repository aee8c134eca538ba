"""MOT metrics (reference utils/metrics.py:7-61 -> trackmpnn_b200/metrics.py) on hand-computed cases.

The reference delegates to the third-party package motmetrics, which is neither vendored by the reference nor installed
here, so its outputs cannot be generated in this container: every expected number below is worked out by hand from the
CLEAR-MOT / identity-metric definitions motmetrics implements (stated next to each case).  The assignment solver is
checked against scipy.optimize.linear_sum_assignment (test-only use of scipy)."""
import numpy as np
import pytest

from trackmpnn_b200 import metrics as M


def _seq(gt, pred):
    """gt / pred: lists of (frame, track id, x1, y1, x2, y2) -> the reference's (bbox [N, 14], y [N, 2]) arrays."""
    def pack(rows):
        rows = np.asarray(rows, np.float64).reshape(-1, 6)
        bbox = np.zeros((rows.shape[0], 14))
        bbox[:, 2:6] = rows[:, 2:6]
        return bbox, rows[:, :2].astype(np.int64)
    bg, yg = pack(gt)
    bp, yp = pack(pred)
    return bp, bg, yp, yg


BOX_A, BOX_B, FAR = (0, 0, 10, 10), (100, 0, 110, 10), (500, 500, 510, 510)


def test_lsap_equals_scipy():
    from scipy.optimize import linear_sum_assignment as ref
    rs = np.random.RandomState(0)
    for nr, nc in ((1, 1), (3, 5), (6, 4), (12, 12), (25, 31), (40, 17)):
        for _ in range(5):
            c = rs.uniform(0, 1, (nr, nc))
            r0, c0 = ref(c)
            r1, c1 = M.linear_sum_assignment(c)
            assert len(r1) == min(nr, nc) and len(set(c1.tolist())) == len(c1)
            assert abs(c[r0, c0].sum() - c[r1, c1].sum()) < 1e-9
    c = np.round(rs.uniform(0, 3, (9, 9)))            # many ties: the optimum value is still unique
    assert abs(c[ref(c)].sum() - c[M.linear_sum_assignment(c)].sum()) < 1e-12


def test_iou_distance():
    d = M.iou_distance_matrix([[0, 0, 10, 10]], [[0, 0, 10, 8], [0, 0, 10, 4], [50, 50, 5, 5]])
    assert d.shape == (1, 3)
    assert abs(d[0, 0] - 0.2) < 1e-12          # IoU 80 / 100
    assert np.isnan(d[0, 1]) and np.isnan(d[0, 2])   # IoU 0.4 and 0: 1 - IoU exceeds max_iou = 0.5 -> not matchable


def test_perfect_tracking():
    gt = [(t, 0) + BOX_A for t in range(3)] + [(t, 1) + BOX_B for t in range(3)]
    pred = [(t, 7) + BOX_A for t in range(3)] + [(t, 9) + BOX_B for t in range(3)]
    m = M.calc_mot_metrics([M.create_mot_accumulator(*_seq(gt, pred))])
    assert m['mota'] == 1.0 and m['motp'] == 0.0 and m['idf1'] == 1.0 and m['idp'] == 1.0 and m['idr'] == 1.0
    assert m['recall'] == 1.0 and m['precision'] == 1.0
    assert (m['num_unique_objects'], m['mostly_tracked'], m['partially_tracked'], m['mostly_lost']) == (2, 2, 0, 0)
    assert (m['num_false_positives'], m['num_misses'], m['num_switches'], m['num_fragmentations']) == (0, 0, 0, 0)
    assert list(m) == M.MOTCHALLENGE_METRICS


def test_switch_miss_false_positive_and_motp():
    """Object 0 over 4 frames: hypothesis 5 (frames 0-1, box 20 % off: distance 0.2), nothing matchable in frame 2 (a far
    hypothesis 6: one MISS + one FP), hypothesis 8 in frame 3 (exact box): the object's last match was 5 -> SWITCH.
    objects 4, matches 2 + switch 1: recall 3/4, precision 3/4, MOTA = 1 - (1 miss + 1 switch + 1 fp) / 4 = 0.25,
    MOTP = (0.2 + 0.2 + 0) / 3, fragmentations: tracked -> missed once inside the span = 1.
    identity: hypothesis 5 shares 2 frames with the object, 8 one -> IDTP 2; predictions 4 -> IDF1 = 2*2 / (4 + 4)."""
    gt = [(t, 0) + BOX_A for t in range(4)]
    pred = [(0, 5, 0, 0, 10, 8), (1, 5, 0, 0, 10, 8), (2, 6) + FAR, (3, 8) + BOX_A]
    m = M.calc_mot_metrics([M.create_mot_accumulator(*_seq(gt, pred))])
    assert (m['num_misses'], m['num_false_positives'], m['num_switches'], m['num_fragmentations']) == (1, 1, 1, 1)
    assert abs(m['mota'] - 0.25) < 1e-12 and abs(m['motp'] - 0.4 / 3) < 1e-12
    assert abs(m['recall'] - 0.75) < 1e-12 and abs(m['precision'] - 0.75) < 1e-12
    assert abs(m['idf1'] - 0.5) < 1e-12 and abs(m['idp'] - 0.5) < 1e-12 and abs(m['idr'] - 0.5) < 1e-12
    assert (m['mostly_tracked'], m['partially_tracked'], m['mostly_lost']) == (0, 1, 0)   # tracked 3 of 4 frames = 0.75


def test_previous_correspondence_is_kept():
    """CLEAR-MOT keeps last frame's pairing while it stays valid even when another hypothesis is closer: frame 1 offers the
    object its old hypothesis 1 at distance 0.2 and a new hypothesis 2 at distance 0 -> MATCH with 1, hypothesis 2 is a FP,
    no switch."""
    gt = [(0, 0) + BOX_A, (1, 0) + BOX_A]
    pred = [(0, 1) + BOX_A, (1, 1, 0, 0, 10, 8), (1, 2) + BOX_A]
    acc = M.create_mot_accumulator(*_seq(gt, pred))
    m = M.calc_mot_metrics([acc])
    assert (m['num_switches'], m['num_false_positives'], m['num_misses']) == (0, 1, 0)
    assert abs(m['motp'] - 0.1) < 1e-12
    assert [e[1:4] for e in acc.events] == [('MATCH', 0.0, 1.0), ('MATCH', 0.0, 1.0), ('FP', None, 2.0)]


def test_track_ratios_and_overall_over_sequences():
    """Sequence 0: object A tracked 5/5 (MT), object B tracked 1/5 = 0.2 (partially: >= 0.2), object C never (ML).
    Sequence 1: one object, one hypothesis, perfect, 2 frames.  The overall row pools the counts:
    objects 15 + 2, misses 4 + 5, no fp / switch -> MOTA = 1 - 9/17; identical ids in different sequences stay apart."""
    box_c = (300, 0, 310, 10)
    gt0 = [(t, 0) + BOX_A for t in range(5)] + [(t, 1) + BOX_B for t in range(5)] + [(t, 2) + box_c for t in range(5)]
    pred0 = [(t, 0) + BOX_A for t in range(5)] + [(0, 1) + BOX_B]
    gt1 = [(t, 0) + BOX_A for t in range(2)]
    pred1 = [(t, 0) + BOX_A for t in range(2)]
    accs = [M.create_mot_accumulator(*_seq(gt0, pred0)), M.create_mot_accumulator(*_seq(gt1, pred1))]
    m = M.calc_mot_metrics(accs)
    assert (m['num_unique_objects'], m['mostly_tracked'], m['partially_tracked'], m['mostly_lost']) == (4, 2, 1, 1)
    assert (m['num_misses'], m['num_false_positives'], m['num_switches'], m['num_fragmentations']) == (9, 0, 0, 0)
    assert abs(m['mota'] - (1 - 9 / 17)) < 1e-12
    assert abs(m['idf1'] - 2 * 8 / (17 + 8)) < 1e-12          # IDTP 5 + 1 + 2, predictions 8
    one = M.calc_mot_metrics(accs[1:])
    assert one['mota'] == 1.0 and one['idf1'] == 1.0


def test_empty_and_negative_ids():
    """Rows with track id -1 (undecoded detections, false-positive ground truth) are ignored (utils/metrics.py:29,33)."""
    gt = [(0, -1) + BOX_A, (0, 3) + BOX_B]
    pred = [(0, -1) + BOX_B, (0, 4) + BOX_B]
    m = M.calc_mot_metrics([M.create_mot_accumulator(*_seq(gt, pred))])
    assert (m['num_unique_objects'], m['num_misses'], m['num_false_positives']) == (1, 0, 0) and m['mota'] == 1.0
    m = M.calc_mot_metrics([M.MOTAccumulator()])
    assert m['num_unique_objects'] == 0 and np.isnan(m['mota'])
