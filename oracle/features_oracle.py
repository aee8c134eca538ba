"""CPU restatement of the reference's detection ingest and feature construction -- TEST INFRASTRUCTURE: only tests/
import it.  Pinned by tests/golden/features.npz, written by the reference's own
``KittiMOTDataset`` on a synthetic KITTI tree (tests/golden/make_golden_features.py)."""
import numpy as np

_KITTI_TYPES = ['Pedestrian', 'Car', 'Cyclist', 'Van', 'Truck', 'Person', 'Tram', 'Misc', 'DontCare']


def kitti_cats(cat):
    """The category filter of ``dataset/kitti_mot.py:86-91``."""
    if cat == 'All':
        return ['Pedestrian', 'Car', 'Cyclist', 'Van', 'DontCare']
    if cat == 'Car':
        return ['Car', 'Van', 'DontCare']
    return [cat, 'DontCare']


def parse_kitti_detections(lines_by_frame, cat):
    """``load_detections`` (``dataset/kitti_mot.py:311-365``) over the frames in order: rows
    ``[fr, -1, cat_id, -10, x1, y1, x2, y2, -1, -1, -1, -1000, -1000, -1000, -10, score]`` (float32), types outside
    the category filter and every 'Van' dropped."""
    ids = {c: i + 1 for i, c in enumerate(_KITTI_TYPES)}
    cats = kitti_cats(cat)
    rows = []
    for fr in sorted(lines_by_frame):
        for line in lines_by_frame[fr]:
            tmp = line.split(',')
            if tmp[0] not in cats or tmp[0] == 'Van':
                continue
            rows.append([fr, -1, ids[tmp[0]], -10, float(tmp[1]), float(tmp[2]), float(tmp[3]), float(tmp[4]), -1, -1, -1,
                         -1000, -1000, -1000, -10, float(tmp[5])])
    return np.asarray(rows, dtype=np.float32).reshape(-1, 16)


def parse_bdd100k_detections(lines_by_frame, cat):
    """``load_detections`` of the BDD100K loader (``dataset/bdd100k_mot.py:295-350``): same row layout, category ids of
    its ``class_dict``; types outside ``self.cats`` (``:83-86``), distractor types and scores <= 0.8 dropped."""
    class_dict = {'pedestrian': 1, 'rider': 2, 'car': 3, 'bus': 4, 'truck': 5, 'train': 6, 'motorcycle': 7, 'bicycle': 8}
    distractors = {'other person': 9, 'trailer': 9, 'other vehicle': 9, 'crowd': -1}
    cat_ids = {**class_dict, **distractors}
    cats = (list(class_dict) if cat == 'All' else [cat]) + list(distractors)
    rows = []
    for fr in sorted(lines_by_frame):
        for line in lines_by_frame[fr]:
            tmp = line.split(',')
            cid, score = cat_ids[tmp[0]], float(tmp[5])
            if tmp[0] not in cats or tmp[0] in distractors or score <= 0.8:
                continue
            rows.append([fr, -1, cid, -10, float(tmp[1]), float(tmp[2]), float(tmp[3]), float(tmp[4]), -1, -1, -1,
                         -1000, -1000, -1000, -10, score])
    return np.asarray(rows, dtype=np.float32).reshape(-1, 16)


def norm_constants(dataset, detections, feats, ncat):
    """Hard-coded mean / std rows (``dataset/kitti_mot.py:155-177``, ``dataset/bdd100k_mot.py:154-176``)."""
    two_d = {('kitti', 'centertrack'): ([0.78, 544.57, 171.58, 71.54, 61.50], [0.14, 285.65, 13.94, 69.92, 47.39]),
             ('kitti', 'rrc'): ([0.91, 577.11, 178.39, 102.48, 58.36], [0.21, 301.75, 11.55, 78.83, 44.66]),
             ('bdd100k', 'hin'): ([0.94, 545.84, 329.28, 85.19, 71.47], [0.07, 294.88, 81.51, 93.51, 75.72]),
             ('bdd100k', 'libra'): ([0.94, 545.84, 329.28, 85.19, 71.47], [0.07, 294.88, 81.51, 93.51, 75.72])}
    mean, std = [0.5] * ncat, [0.5] * ncat
    if '2d' in feats:
        m, s = two_d[(dataset, detections)]
        mean, std = mean + m, std + s
    if 'temp' in feats:
        mean, std = mean + [0.0, 0.0], std + [1.0, 1.0]
    return np.asarray([mean], np.float32), np.asarray([std], np.float32)


def build_features(bbox_pred, ncat, feats, mean, std, fr_range=30):
    """``dataset/kitti_mot.py:545-566`` without the visual block: one-hot category | [score, xc, yc, w, h] |
    [sin, cos](pi (frame mod fr_range) / fr_range) (``:414-420``), then ``(features - mean) / std``."""
    b = np.asarray(bbox_pred, np.float32)
    out = [np.eye(ncat, dtype=np.float32)[b[:, 2].astype('int64') - 1]]
    if '2d' in feats:
        out.append(np.stack((b[:, 15], (b[:, 4] + b[:, 6]) / 2.0, (b[:, 5] + b[:, 7]) / 2.0, b[:, 6] - b[:, 4],
                             b[:, 7] - b[:, 5]), axis=1))
    if 'temp' in feats:
        f = np.mod(b[:, 0:1], fr_range) * np.pi / fr_range
        out.append(np.concatenate((np.sin(f), np.cos(f)), axis=1))
    x = np.concatenate(out, axis=1).astype(np.float32)
    return (x - mean) / std if x.shape[0] else x
