"""Golden feature fixtures: the UNMODIFIED reference ``KittiMOTDataset`` (``dataset/kitti_mot.py``) and
``BDD100kMOTDataset`` (``dataset/bdd100k_mot.py``) run on small synthetic dataset trees written to a temp dir (build container only; ``/root/reference`` does not exist on the GPU box).

    python tests/golden/make_golden_features.py

Stores, per configuration, the detection files' text, the ``bbox_pred`` array the dataset parsed and the normalised
feature matrix it built (``dataset/kitti_mot.py:311-365, 545-566``; ``dataset/bdd100k_mot.py:295-350, 530-551``) in
``tests/golden/features.npz`` (keys ``<config>/<sequence>/{features,bbox_pred}``; BDD configs start with ``bdd_``)."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
CATS = ['Pedestrian', 'Car', 'Cyclist', 'Van', 'Truck', 'Person', 'Tram', 'Misc', 'DontCare']


def detection_lines(seed, frames=7):
    """{frame: [csv line, ...]} in the format ``load_detections`` reads: ``type,x1,y1,x2,y2,score``."""
    rs = np.random.RandomState(seed)
    out = {}
    for fr in range(frames):
        lines = []
        for _ in range(rs.randint(0, 26)):
            cat = CATS[rs.randint(0, len(CATS) - 1)]   # no 'DontCare' detections: the reference's one-hot would index out of range
            x1, y1 = rs.uniform(0, 1100), rs.uniform(0, 300)
            w, h = rs.uniform(8, 200), rs.uniform(8, 150)
            lines.append('%s,%.2f,%.2f,%.2f,%.2f,%.4f' % (cat, x1, y1, x1 + w, y1 + h, rs.uniform(0.05, 1.0)))
        out[fr] = lines
    return out


BDD_CATS = ['pedestrian', 'rider', 'car', 'bus', 'truck', 'train', 'motorcycle', 'bicycle', 'other person', 'trailer',
            'other vehicle', 'crowd']
BDD_MISSING = {('b0001', 3)}   # a frame without a detection file: the loader treats it as empty


def bdd_detection_lines(seed, frames=6):
    """Like ``detection_lines`` for the BDD100K loader: its twelve type names, scores on both sides of the 0.8 floor
    (one exactly 0.8), 1280x720 boxes."""
    rs = np.random.RandomState(seed)
    out = {}
    for fr in range(frames):
        lines = []
        for k in range(rs.randint(0, 31)):
            cat = BDD_CATS[rs.randint(0, len(BDD_CATS))]
            x1, y1 = rs.uniform(0, 1150), rs.uniform(0, 620)
            w, h = rs.uniform(8, 220), rs.uniform(8, 160)
            score = 0.8 if k == 4 else rs.uniform(0.6, 1.0)
            lines.append('%s,%.2f,%.2f,%.2f,%.2f,%.4f' % (cat, x1, y1, x1 + w, y1 + h, score))
        out[fr] = lines
    return out


if __name__ == '__main__':
    sys.path.insert(0, '/root/reference')
    import PIL.Image
    import trackmpnn_b200.overlay as overlay
    overlay.install(stub_missing=True)
    from dataset.kitti_mot import KittiMOTDataset
    store = {}
    with tempfile.TemporaryDirectory() as root:
        for seq, seed in (('0000', 1), ('0001', 2)):
            dets = detection_lines(seed)
            os.makedirs(os.path.join(root, 'testing', 'image_02', seq))
            os.makedirs(os.path.join(root, 'testing', 'centertrack_detections', seq))
            os.makedirs(os.path.join(root, 'testing', 'rrc_detections', seq))
            for fr, lines in dets.items():
                PIL.Image.new('RGB', (1242, 375)).save(os.path.join(root, 'testing', 'image_02', seq, '%.6d.png' % fr))
                for d in ('centertrack_detections', 'rrc_detections'):
                    with open(os.path.join(root, 'testing', d, seq, '%.4d.txt' % fr), 'w') as f:
                        f.write(''.join(l + '\n' for l in lines))
        for name, kw in (('all_2d', dict(cat='All', detections='centertrack', feats='2d')),
                         ('car_2d_temp', dict(cat='Car', detections='centertrack', feats='2d+temp')),
                         ('ped_rrc', dict(cat='Pedestrian', detections='rrc', feats='2d+temp'))):
            ds = KittiMOTDataset(root, 'test', kw['cat'], kw['detections'], kw['feats'], 'espv2', 5, 0, None, False, False)
            for i in range(len(ds)):
                feats, bbox_pred, bbox_gt, loss = ds[i]
                store[f'{name}/{i}/features'] = feats.numpy()
                store[f'{name}/{i}/bbox_pred'] = bbox_pred
    from dataset.bdd100k_mot import BDD100kMOTDataset
    with tempfile.TemporaryDirectory() as root:
        for seq, seed in (('b0000', 11), ('b0001', 12)):
            dets = bdd_detection_lines(seed)
            os.makedirs(os.path.join(root, 'testing', 'image_02', seq))
            for d in ('hin_detections', 'libra_detections'):
                os.makedirs(os.path.join(root, 'testing', d, seq))
            for fr, lines in dets.items():
                PIL.Image.new('RGB', (1280, 720)).save(os.path.join(root, 'testing', 'image_02', seq, '%.4d.jpg' % fr))
                if (seq, fr) in BDD_MISSING:
                    continue
                for d in ('hin_detections', 'libra_detections'):
                    with open(os.path.join(root, 'testing', d, seq, '%.4d.txt' % fr), 'w') as f:
                        f.write(''.join(l + '\n' for l in lines))
        for name, kw in (('bdd_all_2d', dict(cat='All', detections='hin', feats='2d')),
                         ('bdd_all_2d_temp', dict(cat='All', detections='libra', feats='2d+temp')),
                         ('bdd_car_2d_temp', dict(cat='car', detections='hin', feats='2d+temp'))):
            ds = BDD100kMOTDataset(root, 'test', kw['cat'], kw['detections'], kw['feats'], 'espv2', 5, 0, None, False, False)
            for i in range(len(ds)):
                feats, bbox_pred, bbox_gt, loss = ds[i]
                store[f'{name}/{i}/features'] = feats.numpy()
                store[f'{name}/{i}/bbox_pred'] = bbox_pred
    np.savez_compressed(os.path.join(HERE, 'features.npz'), **store)
    print({k: v.shape for k, v in store.items()})
