"""Feature construction: the oracle restatement against the matrices the reference's own
``KittiMOTDataset`` built on a synthetic KITTI tree (tests/golden/features.npz)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
from make_golden_features import BDD_MISSING, bdd_detection_lines, detection_lines  # noqa: E402

from oracle import features_oracle as FO  # noqa: E402

CONFIGS = {'all_2d': dict(cat='All', detections='centertrack', feats='2d'),
           'car_2d_temp': dict(cat='Car', detections='centertrack', feats='2d+temp'),
           'ped_rrc': dict(cat='Pedestrian', detections='rrc', feats='2d+temp')}
BDD_CONFIGS = {'bdd_all_2d': dict(cat='All', detections='hin', feats='2d'),
               'bdd_all_2d_temp': dict(cat='All', detections='libra', feats='2d+temp'),
               'bdd_car_2d_temp': dict(cat='car', detections='hin', feats='2d+temp')}
GOLD = np.load(os.path.join(HERE, 'golden', 'features.npz'))


@pytest.mark.parametrize('name', sorted(CONFIGS))
@pytest.mark.parametrize('seq', [0, 1])
def test_oracle_matches_reference_dataset(name, seq):
    kw = CONFIGS[name]
    lines = detection_lines(seq + 1)
    bbox = FO.parse_kitti_detections(lines, kw['cat'])
    np.testing.assert_array_equal(bbox, GOLD[f'{name}/{seq}/bbox_pred'])
    mean, std = FO.norm_constants('kitti', kw['detections'], kw['feats'], 3)
    x = FO.build_features(bbox, 3, kw['feats'], mean, std)
    want = GOLD[f'{name}/{seq}/features']
    assert x.shape == want.shape and x.dtype == np.float32
    np.testing.assert_array_equal(x, want)


def bdd_lines(seq):
    """The detection files the golden script wrote for BDD sequence ``seq`` (a missing file = an empty frame)."""
    name = 'b%04d' % seq
    return {fr: ([] if (name, fr) in BDD_MISSING else lines) for fr, lines in bdd_detection_lines(11 + seq).items()}


@pytest.mark.parametrize('name', sorted(BDD_CONFIGS))
@pytest.mark.parametrize('seq', [0, 1])
def test_bdd100k_oracle_matches_reference_dataset(name, seq):
    kw = BDD_CONFIGS[name]
    bbox = FO.parse_bdd100k_detections(bdd_lines(seq), kw['cat'])
    np.testing.assert_array_equal(bbox, GOLD[f'{name}/{seq}/bbox_pred'])
    mean, std = FO.norm_constants('bdd100k', kw['detections'], kw['feats'], 8)
    x = FO.build_features(bbox, 8, kw['feats'], mean, std)
    np.testing.assert_array_equal(x, GOLD[f'{name}/{seq}/features'])


@pytest.mark.parametrize('name', sorted(BDD_CONFIGS))
def test_bdd100k_host_ingest_matches_reference_dataset(name, tmp_path):
    """The product's host-side reader (file I/O + parse, no GPU work) on the same files, incl. the missing one."""
    from trackmpnn_b200 import features as F
    kw = BDD_CONFIGS[name]
    for seq in (0, 1):
        d = tmp_path / ('b%04d' % seq)
        d.mkdir()
        lines = bdd_detection_lines(11 + seq)
        for fr, ls in lines.items():
            if ('b%04d' % seq, fr) not in BDD_MISSING:
                (d / ('%.4d.txt' % fr)).write_text(''.join(l + '\n' for l in ls))
        b = F.load_bdd100k_detections(str(tmp_path), 'b%04d' % seq, sorted(lines), kw['cat'])
        np.testing.assert_array_equal(b.numpy(), GOLD[f'{name}/{seq}/bbox_pred'])
    with pytest.raises(KeyError):
        F.parse_bdd100k_detection_lines({0: ['tricycle,1,2,3,4,0.9']})


def test_kitti_host_ingest_matches_reference_dataset():
    from trackmpnn_b200 import features as F
    for name, kw in CONFIGS.items():
        for seq in (0, 1):
            b = F.parse_kitti_detection_lines(detection_lines(seq + 1), kw['cat'])
            np.testing.assert_array_equal(b, GOLD[f'{name}/{seq}/bbox_pred'])
