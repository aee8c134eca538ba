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
    Vectorised per frame so that bench-sized streams (256 x 200 x 80) build in seconds.
    """
    rs = np.random.RandomState(seed)
    mean2d, std2d, ncat, (img_h, img_w) = _NORM[dataset]
    mean2d = np.asarray(mean2d, np.float64); std2d = np.asarray(std2d, np.float64)
    if timestamps is None:
        timestamps = list(range(frames))
    tp_target = dets_per_frame * (1.0 - fp_rate)
    live_target = tp_target / (1.0 - miss_rate)
    death = 0.03
    next_id = 0

    def spawn(n):
        nonlocal next_id
        w = rs.uniform(20, 160, n)
        o = np.stack((np.arange(next_id, next_id + n, dtype=np.float64), rs.randint(ncat, size=n).astype(np.float64),
                      rs.uniform(0, img_w, n), rs.uniform(0.3 * img_h, 0.8 * img_h, n), w,
                      w * rs.uniform(0.5, 1.2, n), rs.normal(0, 6.0, n), rs.normal(0, 1.5, n)), 1)
        next_id += n
        return o  # columns: id, cat, xc, yc, w, h, vx, vy

    objs = spawn(rs.poisson(live_target) if poisson else int(round(live_target)))
    xs, ys = [], []
    t_prev = timestamps[0]
    for t in timestamps:
        for _ in range(t - t_prev):  # advance the world
            objs = objs[rs.uniform(size=objs.shape[0]) > death]
            nb = rs.poisson(death * live_target) if poisson else int(rs.uniform() < death * live_target)
            if nb:
                objs = np.concatenate((objs, spawn(nb)), 0)
            objs[:, 2] += objs[:, 6] + rs.normal(0, 1.0, objs.shape[0])
            objs[:, 3] += objs[:, 7] + rs.normal(0, 0.5, objs.shape[0])
        t_prev = t
        seen = objs[rs.uniform(size=objs.shape[0]) >= miss_rate]
        n_tp = seen.shape[0]
        tp = np.stack((seen[:, 0], seen[:, 1], np.clip(rs.normal(0.85, 0.1, n_tp), 0.3, 1.0),
                       seen[:, 2] + rs.normal(0, 1.5, n_tp), seen[:, 3] + rs.normal(0, 1.0, n_tp),
                       seen[:, 4] * (1 + rs.normal(0, 0.02, n_tp)), seen[:, 5] * (1 + rs.normal(0, 0.02, n_tp))), 1)
        nfp = rs.poisson(dets_per_frame * fp_rate) if poisson else int(round(dets_per_frame * fp_rate))
        w = rs.uniform(20, 160, nfp)
        fp = np.stack((-np.ones(nfp), rs.randint(ncat, size=nfp).astype(np.float64),
                       np.clip(rs.normal(0.5, 0.15, nfp), 0.3, 1.0), rs.uniform(0, img_w, nfp),
                       rs.uniform(0.3 * img_h, 0.8 * img_h, nfp), w, w * rs.uniform(0.5, 1.2, nfp)), 1)
        fr = np.concatenate((tp, fp), 0)[rs.permutation(n_tp + nfp)]  # columns: id, cat, score, xc, yc, w, h
        onehot = np.zeros((fr.shape[0], ncat), np.float32)
        onehot[np.arange(fr.shape[0]), fr[:, 1].astype(np.int64)] = 1.0
        f2d = ((fr[:, 2:7] - mean2d) / std2d).astype(np.float32)
        xs.append(np.concatenate(((onehot - 0.5) / 0.5, f2d), 1))
        ys.append(np.stack((np.full(fr.shape[0], float(t)), fr[:, 0]), 1))
    F = ncat + 5
    X = np.concatenate(xs, 0).astype(np.float32).reshape(1, -1, F)
    y = np.concatenate(ys, 0).astype(np.float32).reshape(1, -1, 2)
    return X, y


def train_chunk_timestamps(seed, cur_win_size=5, ret_win_size=0, start=0):
    """Timestamps of one training chunk: ``cur_win_size`` consecutive frames plus two
    "skip" frames (reference ``dataset/kitti_mot.py:220-227``)."""
    rs = np.random.RandomState(seed)
    fr = list(range(start, start + cur_win_size))
    skip = int(rs.randint(start + cur_win_size, start + cur_win_size + ret_win_size + 1))
    return fr + [skip, skip + 1]
