"""Feature construction: the oracle restatement against the matrices the reference's own
``KittiMOTDataset`` built on a synthetic KITTI tree (tests/golden/features.npz)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
from make_golden_features import detection_lines  # noqa: E402

from oracle import features_oracle as FO  # noqa: E402

CONFIGS = {'all_2d': dict(cat='All', detections='centertrack', feats='2d'),
           'car_2d_temp': dict(cat='Car', detections='centertrack', feats='2d+temp'),
           'ped_rrc': dict(cat='Pedestrian', detections='rrc', feats='2d+temp')}
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
