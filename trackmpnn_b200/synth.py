"""Seeded synthetic detection streams shaped like the reference's datasets.

The reference builds one feature row per detection as
``[one-hot category | score, xc, yc, w, h]`` and standardises it with
hard-coded constants (reference ``dataset/kitti_mot.py:155-177, 545-566`` and
``dataset/bdd100k_mot.py:154-176``).  Real KITTI / BDD100K files are not
available, so tests, goldens and the bench use this generator: persistent
objects moving at constant velocity with jitter, a miss rate (mirrors the
reference's ``dropout_ratio`` of 0.2, ``dataset/kitti_mot.py:102``) and a
false-positive rate (``track_id = -1``).

Only numpy's ``RandomState`` is used, so a (seed, config) pair gives the same
stream on every machine.
"""
import numpy as np

# (mean, std) of [score, xc, yc, w, h]; category one-hots use 0.5 / 0.5.
_NORM = {
    "kitti": ([0.78, 544.57, 171.58, 71.54, 61.50], [0.14, 285.65, 13.94, 69.92, 47.39], 3, (375.0, 1242.0)),
    "bdd": ([0.94, 545.84, 329.28, 85.19, 71.47], [0.07, 294.88, 81.51, 93.51, 75.72], 8, (720.0, 1280.0)),
}


def num_features(dataset="kitti"):
    return _NORM[dataset][2] + 5


def num_categories(dataset="kitti"):
    return _NORM[dataset][2]


def make_sequence(seed, frames, dets_per_frame, dataset="kitti", poisson=True,
                  miss_rate=0.2, fp_rate=0.1, timestamps=None):
    """Returns ``X [1, ND, F] float32`` and ``y [1, ND, 2] float32`` (= ``[ts, track_id]``).

    ``frames``: number of timesteps T (timestamps 0..T-1) unless ``timestamps``
    (an increasing list of integer timestamps, e.g. a training chunk
    ``[0,1,2,3,4,7,8]`` as built by reference ``dataset/kitti_mot.py:220-227``)
    is given.  Detections are emitted in time order, as the reference's loaders do.
    """
    rs = np.random.RandomState(seed)
    mean2d, std2d, ncat, (img_h, img_w) = _NORM[dataset]
    if timestamps is None:
        timestamps = list(range(frames))
    tp_target = dets_per_frame * (1.0 - fp_rate)
    live_target = tp_target / (1.0 - miss_rate)
    death = 0.03

    objs = []  # [id, cat, xc, yc, w, h, vx, vy]
    next_id = 0

    def spawn():
        nonlocal next_id
        w = rs.uniform(20, 160)
        h = w * rs.uniform(0.5, 1.2)
        o = [next_id, rs.randint(ncat), rs.uniform(0, img_w), rs.uniform(0.3 * img_h, 0.8 * img_h),
             w, h, rs.normal(0, 6.0), rs.normal(0, 1.5)]
        next_id += 1
        return o

    n0 = rs.poisson(live_target) if poisson else int(round(live_target))
    for _ in range(n0):
        objs.append(spawn())

    rows_x, rows_y = [], []
    t_prev = timestamps[0]
    for t in timestamps:
        dt = t - t_prev
        t_prev = t
        # advance the world dt steps
        for _ in range(dt):
            objs = [o for o in objs if rs.uniform() > death]
            nb = rs.poisson(death * live_target) if poisson else (1 if rs.uniform() < death * live_target else 0)
            for _ in range(nb):
                objs.append(spawn())
            for o in objs:
                o[2] += o[6] + rs.normal(0, 1.0)
                o[3] += o[7] + rs.normal(0, 0.5)
        frame = []
        for o in objs:
            if rs.uniform() < miss_rate:
                continue
            score = float(np.clip(rs.normal(0.85, 0.1), 0.3, 1.0))
            frame.append((o[0], o[1], score, o[2] + rs.normal(0, 1.5), o[3] + rs.normal(0, 1.0),
                          o[4] * (1 + rs.normal(0, 0.02)), o[5] * (1 + rs.normal(0, 0.02))))
        nfp = rs.poisson(dets_per_frame * fp_rate) if poisson else int(round(dets_per_frame * fp_rate))
        for _ in range(nfp):
            w = rs.uniform(20, 160)
            frame.append((-1, rs.randint(ncat), float(np.clip(rs.normal(0.5, 0.15), 0.3, 1.0)),
                          rs.uniform(0, img_w), rs.uniform(0.3 * img_h, 0.8 * img_h), w, w * rs.uniform(0.5, 1.2)))
        order = rs.permutation(len(frame))
        for k in order:
            tid, cat, score, xc, yc, w, h = frame[k]
            onehot = np.zeros(ncat, dtype=np.float32)
            onehot[cat] = 1.0
            f2d = (np.array([score, xc, yc, w, h], dtype=np.float64) - np.array(mean2d)) / np.array(std2d)
            rows_x.append(np.concatenate(((onehot - 0.5) / 0.5, f2d.astype(np.float32))))
            rows_y.append((float(t), float(tid)))
    F = ncat + 5
    X = np.asarray(rows_x, dtype=np.float32).reshape(1, -1, F)
    y = np.asarray(rows_y, dtype=np.float32).reshape(1, -1, 2)
    return X, y


def train_chunk_timestamps(seed, cur_win_size=5, ret_win_size=0, start=0):
    """Timestamps of one training chunk: ``cur_win_size`` consecutive frames plus two
    "skip" frames (reference ``dataset/kitti_mot.py:220-227``)."""
    rs = np.random.RandomState(seed)
    fr = list(range(start, start + cur_win_size))
    skip = int(rs.randint(start + cur_win_size, start + cur_win_size + ret_win_size + 1))
    return fr + [skip, skip + 1]
