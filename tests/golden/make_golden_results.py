"""Golden result files: written by the UNMODIFIED reference's ``store_kitti_results`` /
``store_bdd100k_results`` (build container only; ``/root/reference`` does not exist on the GPU box).

    python tests/golden/make_golden_results.py

Inputs are seeded synthetic detections + track assignments (``results_inputs``, also used by the test)."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

KITTI_CLASSES = {'Pedestrian': 1, 'Car': 2, 'Cyclist': 3}
BDD_CLASSES = {'pedestrian': 1, 'rider': 2, 'car': 3, 'bus': 4, 'truck': 5, 'train': 6, 'motorcycle': 7, 'bicycle': 8}


def results_inputs(seed, ncls, nd=60, frames=9, gap=True):
    """bbox_pred [nd, 14] float64 and y_out [nd, 2] int64 with unique ids per frame, untracked rows (-1), an empty
    frame and low-score tracks (the KITTI writer's Car suppression)."""
    rs = np.random.RandomState(seed)
    ts = np.sort(rs.randint(0, frames, size=nd))
    if gap:
        ts[ts == 4] = 5                                  # frame 4 has no detections at all
    b = np.zeros((nd, 14))
    b[:, 0] = rs.randint(1, ncls + 1, size=nd)
    b[:, 1] = rs.uniform(-3.14, 3.14, size=nd)
    b[:, 2:4] = rs.uniform(0, 600, size=(nd, 2))
    b[:, 4:6] = b[:, 2:4] + rs.uniform(5, 200, size=(nd, 2))
    b[:, 6:13] = rs.uniform(-20, 60, size=(nd, 7))
    b[:, 13] = rs.uniform(0.3, 1.0, size=nd)
    y = np.stack((ts, np.full(nd, -1)), 1).astype(np.int64)
    nxt = 0
    for t in np.unique(ts):
        idx = np.flatnonzero(ts == t)
        ids = rs.permutation(max(12, idx.size))[:idx.size]   # unique within the frame, shared across frames
        keep = rs.uniform(size=idx.size) > 0.15
        y[idx[keep], 1] = ids[keep]
        nxt += 1
    low = np.isin(y[:, 1], (0, 1, 2))                    # three tracks never score above 0.65 ...
    b[low, 13] = rs.uniform(0.3, 0.65, size=int(low.sum()))
    if ncls == 3:
        b[y[:, 1] == 0, 0] = 2                           # ... one of them all Car, one with a Cyclist in it, one mixed
        b[y[:, 1] == 1, 0] = np.where(rs.uniform(size=int((y[:, 1] == 1).sum())) < 0.5, 2, 3)
    return b, y


if __name__ == '__main__':
    sys.path.insert(0, '/root/reference')
    import trackmpnn_b200.overlay as overlay
    overlay.install(stub_missing=True)                   # the dataset modules import DCNv2 / cv2-free stubs only
    from dataset.kitti_mot import store_kitti_results
    from dataset.bdd100k_mot import store_bdd100k_results
    out = os.path.join(HERE, 'results')
    os.makedirs(out, exist_ok=True)
    # The reference's BDD writer hands numpy int32 track ids to json.dump, which no json version serialises
    # (TypeError at dataset/bdd100k_mot.py:66).  The function itself stays unmodified; only the encoder's fallback is
    # taught to write numpy integers as plain ints, which is the evident intent.
    import json
    json.JSONEncoder.default = lambda self, o: int(o) if isinstance(o, np.integer) else float(o)
    for seed in (1, 2, 3):
        b, y = results_inputs(seed, 3)
        store_kitti_results(b, y.copy(), KITTI_CLASSES, os.path.join(out, f'kitti_{seed}.txt'))
        b, y = results_inputs(10 + seed, 8)
        store_bdd100k_results(b, y.copy(), BDD_CLASSES, os.path.join(out, f'bdd_{seed}.json'))
    print('wrote', sorted(os.listdir(out)))
