"""Result writers: the step after the hot path.  Decoded tracks -> the on-wire formats of
the reference's ``store_kitti_results`` (``dataset/kitti_mot.py:21-73``) and ``store_bdd100k_results``
(``dataset/bdd100k_mot.py:22-67``): same arguments, same files byte for byte (tests/test_results_writers.py
compares with files the reference's own functions wrote).  Host code (file I/O); the arrays come from
``TrackEngine.results()`` or ``decode_tracks``'s ``y_out``.
"""
import json
import os

import numpy as np


def _frames(y_out):
    """(t_first, t_last, {t: ascending detection indices with a track id})."""
    y_out = np.asarray(y_out)
    ts = y_out[:, 0].astype(np.int64)
    order = np.argsort(ts, kind='stable')          # np.where order inside a frame == ascending detection index
    kept = order[y_out[order, 1] != -1]
    by_t = {}
    for i in kept:
        by_t.setdefault(int(ts[i]), []).append(int(i))
    return int(ts.min()), int(ts.max()), by_t


def _check_unique(tracks):
    assert tracks.size == np.unique(tracks).size, 'Same track ID assigned to two detections from same timestep!'


def store_kitti_results(bbox_pred, y_out, class_dict, output_path):
    """KITTI tracking text format, one line per tracked detection:
    ``frame id class -1 -1 alpha x1 y1 x2 y2 h w l x y z rotation_y score`` (``%.2f``).  Like the reference it first
    drops (in place, ``y_out[:, 1] = -1``) every track whose class is Car and whose best detection score is below 0.7.
    bbox_pred [ND, 14] = (cat_id, alpha, x1, y1, x2, y2, h, w, l, x, y, z, rotation_y, score); y_out [ND, 2] = (ts, id)."""
    names = {v: k for k, v in class_dict.items()}
    out_dir = os.path.dirname(output_path)
    if out_dir and not os.path.exists(out_dir):
        os.makedirs(out_dir)
    bbox_pred = np.asarray(bbox_pred)
    ids = y_out[:, 1]
    valid = ids >= 0
    if valid.any():
        # per track: highest category id and highest score over its detections (dataset/kitti_mot.py:36-45)
        uniq, inv = np.unique(ids[valid], return_inverse=True)
        best_cat = np.full(uniq.size, -np.inf)
        best_score = np.full(uniq.size, -np.inf)
        np.maximum.at(best_cat, inv, bbox_pred[valid, 0])
        np.maximum.at(best_score, inv, bbox_pred[valid, 13])
        drop = np.array([names[int(c)] == 'Car' for c in best_cat]) & (best_score < 0.7)
        if drop.any():
            y_out[np.flatnonzero(valid)[drop[inv]], 1] = -1
    t0, t1, by_t = _frames(y_out)
    lines = []
    for t in range(t0, t1 + 1):
        hids = np.asarray(by_t.get(t, []), dtype=np.int64)
        tracks = y_out[hids, 1].astype('int64')
        _check_unique(tracks)
        for i, d in enumerate(hids):
            b = bbox_pred[d]
            lines.append('%d %d %s -1 -1 %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f\n' %
                         (t, tracks[i], names[int(b[0])], b[1], b[2], b[3], b[4], b[5], b[6], b[7], b[8], b[9], b[10],
                          b[11], b[12], b[13]))
    with open(output_path, 'w') as f:
        f.writelines(lines)


def store_bdd100k_results(bbox_pred, y_out, class_dict, output_path):
    """BDD100K tracking JSON: a list with one entry per frame in [first, last] --
    ``{name, videoName, frameIndex, labels: [{id, category, box2d: {x1, y1, x2, y2}}]}``."""
    names = {v: k for k, v in class_dict.items()}
    out_dir = os.path.dirname(output_path)
    if out_dir and not os.path.exists(out_dir):
        os.makedirs(out_dir)
    bbox_pred = np.asarray(bbox_pred)
    t0, t1, by_t = _frames(y_out)
    base = os.path.basename(output_path)
    data = []
    for t in range(t0, t1 + 1):
        hids = np.asarray(by_t.get(t, []), dtype=np.int64)
        tracks = y_out[hids, 1].astype('int32')
        _check_unique(tracks)
        labels = [{'id': tracks[i], 'category': names[int(bbox_pred[d, 0])],
                   'box2d': {'x1': bbox_pred[d, 2], 'y1': bbox_pred[d, 3], 'x2': bbox_pred[d, 4], 'y2': bbox_pred[d, 5]}}
                  for i, d in enumerate(hids)]
        data.append({'name': base, 'videoName': base, 'frameIndex': int(t), 'labels': labels})
    with open(output_path, 'w') as f:
        json.dump(data, f, default=_json_scalar)


def _json_scalar(o):
    # the reference passes numpy scalars to json.dump, which only works for float64 (a float subclass); accept the
    # other numpy scalar types an engine result carries as well
    if isinstance(o, np.integer):
        return int(o)
    if isinstance(o, np.floating):
        return float(o)
    raise TypeError(f'Object of type {type(o).__name__} is not JSON serializable')
