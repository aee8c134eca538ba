"""Detection ingest and feature construction: the step in front of the hot path.

``load_kitti_detections`` / ``load_bdd100k_detections`` read the per-frame detection text files the reference's loaders
read (``dataset/kitti_mot.py:311-365``, ``dataset/bdd100k_mot.py:295-350``; host I/O), ``build_features`` turns the detection records into the normalised
feature matrix ``X`` the tracker consumes (``dataset/kitti_mot.py:545-566``, ``dataset/bdd100k_mot.py:530-551``) with one
CUDA kernel (``tmpnn_build_features``).  The 'vis' block (the embedding CNN) is out of scope.
"""
import os

import numpy as np
import torch

from . import _lib as L

_KITTI_TYPES = ['Pedestrian', 'Car', 'Cyclist', 'Van', 'Truck', 'Person', 'Tram', 'Misc', 'DontCare']
# (dataset, detections) -> mean, std of [score, xc, yc, w, h] (dataset/kitti_mot.py:155-177, bdd100k_mot.py:154-176)
_NORM_2D = {('kitti', 'centertrack'): ([0.78, 544.57, 171.58, 71.54, 61.50], [0.14, 285.65, 13.94, 69.92, 47.39]),
            ('kitti', 'rrc'): ([0.91, 577.11, 178.39, 102.48, 58.36], [0.21, 301.75, 11.55, 78.83, 44.66]),
            ('bdd100k', 'hin'): ([0.94, 545.84, 329.28, 85.19, 71.47], [0.07, 294.88, 81.51, 93.51, 75.72]),
            ('bdd100k', 'libra'): ([0.94, 545.84, 329.28, 85.19, 71.47], [0.07, 294.88, 81.51, 93.51, 75.72])}


def kitti_category_filter(cat):
    """Detection types kept for ``--category`` (``dataset/kitti_mot.py:86-91``); 'Van' is always dropped afterwards."""
    if cat == 'All':
        return ['Pedestrian', 'Car', 'Cyclist', 'Van', 'DontCare']
    if cat == 'Car':
        return ['Car', 'Van', 'DontCare']
    return [cat, 'DontCare']


def parse_kitti_detection_lines(lines_by_frame, cat='All'):
    """{frame: iterable of ``type,x1,y1,x2,y2,score`` lines} -> ``bbox_pred [ND, 16] float32`` in frame order."""
    ids = {c: i + 1 for i, c in enumerate(_KITTI_TYPES)}
    keep = set(kitti_category_filter(cat)) - {'Van'}
    rows = []
    for fr in sorted(lines_by_frame):
        for line in lines_by_frame[fr]:
            tmp = line.rstrip('\n').split(',')
            if tmp[0] in keep:
                rows.append((fr, -1, ids[tmp[0]], -10, float(tmp[1]), float(tmp[2]), float(tmp[3]), float(tmp[4]), -1, -1, -1,
                             -1000, -1000, -1000, -10, float(tmp[5])))
    return np.asarray(rows, dtype=np.float32).reshape(-1, 16)


def load_kitti_detections(detections_path, seq, frames, cat='All'):
    """Reads ``<detections_path>/<seq>/%04d.txt`` for every frame (pinned host buffer ready for an async upload)."""
    lines = {}
    for fr in frames:
        with open(os.path.join(detections_path, seq, '%.4d.txt' % (fr,))) as f:
            lines[fr] = f.readlines()
    b = torch.from_numpy(parse_kitti_detection_lines(lines, cat))
    return b.pin_memory() if torch.cuda.is_available() and b.numel() else b


_BDD_CLASSES = {'pedestrian': 1, 'rider': 2, 'car': 3, 'bus': 4, 'truck': 5, 'train': 6, 'motorcycle': 7, 'bicycle': 8}
_BDD_DISTRACTORS = ('other person', 'trailer', 'other vehicle', 'crowd')
BDD_SCORE_FLOOR = 0.8  # detections with score <= 0.8 never reach the tracker (dataset/bdd100k_mot.py:340-341)


def parse_bdd100k_detection_lines(lines_by_frame, cat='All'):
    """{frame: iterable of ``type,x1,y1,x2,y2,score`` lines} -> ``bbox_pred [ND, 16] float32`` in frame order, with the
    BDD100K rules (``dataset/bdd100k_mot.py:319-350``): the eight tracked classes (or the one ``--category``), distractor
    classes and every detection scoring <= 0.8 dropped; an unknown type is a ``KeyError`` like in the reference."""
    keep = set(_BDD_CLASSES) if cat == 'All' else {cat}
    rows = []
    for fr in sorted(lines_by_frame):
        for line in lines_by_frame[fr]:
            tmp = line.rstrip('\n').split(',')
            if tmp[0] not in _BDD_CLASSES and tmp[0] not in _BDD_DISTRACTORS:
                raise KeyError(tmp[0])
            score = float(tmp[5])
            if tmp[0] in keep and tmp[0] in _BDD_CLASSES and score > BDD_SCORE_FLOOR:
                rows.append((fr, -1, _BDD_CLASSES[tmp[0]], -10, float(tmp[1]), float(tmp[2]), float(tmp[3]), float(tmp[4]), -1, -1,
                             -1, -1000, -1000, -1000, -10, score))
    return np.asarray(rows, dtype=np.float32).reshape(-1, 16)


def load_bdd100k_detections(detections_path, seq, frames, cat='All'):
    """Reads ``<detections_path>/<seq>/%04d.txt`` for every frame; a frame without a file has no detections
    (``dataset/bdd100k_mot.py:321-324``)."""
    lines = {}
    for fr in frames:
        try:
            with open(os.path.join(detections_path, seq, '%.4d.txt' % (fr,))) as f:
                lines[fr] = f.readlines()
        except OSError:
            lines[fr] = []
    b = torch.from_numpy(parse_bdd100k_detection_lines(lines, cat))
    return b.pin_memory() if torch.cuda.is_available() and b.numel() else b


def norm_constants(dataset, detections, feats, ncat):
    """The reference's hard-coded mean / std rows for the requested feature groups (``'2d' in feats`` is a substring
    test there too)."""
    mean, std = [0.5] * ncat, [0.5] * ncat
    if '2d' in feats:
        m, s = _NORM_2D[(dataset, detections)]
        mean, std = mean + m, std + s
    if 'temp' in feats:
        mean, std = mean + [0.0, 0.0], std + [1.0, 1.0]
    return torch.tensor(mean, dtype=torch.float32), torch.tensor(std, dtype=torch.float32)


def build_features(bbox_pred, ncat, feats='2d', dataset='kitti', detections='centertrack', fr_range=30, device=None):
    """``bbox_pred [ND, 16]`` (host or device) -> ``X [ND, F] float32`` on the GPU, normalised like the reference."""
    if 'vis' in feats:
        raise NotImplementedError("'vis' features come from the embedding CNN, which is out of scope")
    dev = torch.device(device) if device is not None else (bbox_pred.device if torch.is_tensor(bbox_pred) and bbox_pred.is_cuda
                                                           else torch.device('cuda', torch.cuda.current_device()))
    b = torch.as_tensor(bbox_pred, dtype=torch.float32).to(dev, non_blocking=True).contiguous()
    mean, std = norm_constants(dataset, detections, feats, ncat)
    mean, std = mean.to(dev), std.to(dev)
    nd, F = int(b.shape[0]), int(mean.numel())
    x = torch.empty((nd, F), dtype=torch.float32, device=dev)
    L.call('tmpnn_build_features', L.ptr(b), nd, int(ncat), int('2d' in feats), int('temp' in feats), int(fr_range), L.ptr(mean),
           L.ptr(std), L.ptr(x), F, L.stream())
    return x
