"""GPU: the feature-construction kernel against the oracle on the reference's golden detection records, and on a
bench-sized random batch."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
from make_golden_features import BDD_MISSING, bdd_detection_lines, detection_lines  # noqa: E402

from oracle import features_oracle as FO  # noqa: E402

pytestmark = pytest.mark.gpu
CONFIGS = {'all_2d': dict(cat='All', detections='centertrack', feats='2d'),
           'car_2d_temp': dict(cat='Car', detections='centertrack', feats='2d+temp'),
           'ped_rrc': dict(cat='Pedestrian', detections='rrc', feats='2d+temp')}
GOLD = np.load(os.path.join(HERE, 'golden', 'features.npz'))


@pytest.mark.parametrize('name', sorted(CONFIGS))
def test_golden_records(name):
    from trackmpnn_b200 import features as F
    kw = CONFIGS[name]
    for seq in (0, 1):
        b = F.parse_kitti_detection_lines(detection_lines(seq + 1), kw['cat'])
        np.testing.assert_array_equal(b, GOLD[f'{name}/{seq}/bbox_pred'])
        x = F.build_features(b, 3, kw['feats'], 'kitti', kw['detections']).cpu().numpy()
        want = GOLD[f'{name}/{seq}/features']
        ncol = 8   # one-hot + 2d: same fp32 operations, bit-exact
        np.testing.assert_array_equal(x[:, :ncol], want[:, :ncol])
        np.testing.assert_allclose(x[:, ncol:], want[:, ncol:], atol=2e-7, rtol=0)   # sinf / cosf vs numpy's float32 sin / cos


@pytest.mark.parametrize('name,kw', [('bdd_all_2d', dict(cat='All', detections='hin', feats='2d')),
                                     ('bdd_all_2d_temp', dict(cat='All', detections='libra', feats='2d+temp')),
                                     ('bdd_car_2d_temp', dict(cat='car', detections='hin', feats='2d+temp'))])
def test_bdd100k_golden_records(name, kw):
    from trackmpnn_b200 import features as F
    for seq in (0, 1):
        lines = {fr: ([] if ('b%04d' % seq, fr) in BDD_MISSING else ls) for fr, ls in bdd_detection_lines(11 + seq).items()}
        b = F.parse_bdd100k_detection_lines(lines, kw['cat'])
        np.testing.assert_array_equal(b, GOLD[f'{name}/{seq}/bbox_pred'])
        x = F.build_features(b, 8, kw['feats'], 'bdd100k', kw['detections']).cpu().numpy()
        want = GOLD[f'{name}/{seq}/features']
        np.testing.assert_array_equal(x[:, :13], want[:, :13])
        np.testing.assert_allclose(x[:, 13:], want[:, 13:], atol=2e-7, rtol=0)


def test_large_batch_matches_oracle():
    from trackmpnn_b200 import features as F
    rs = np.random.RandomState(0)
    nd = 300000
    b = np.zeros((nd, 16), np.float32)
    b[:, 0] = rs.randint(0, 500, nd)
    b[:, 2] = rs.randint(1, 9, nd)
    b[:, 4:6] = rs.uniform(0, 1000, (nd, 2)); b[:, 6:8] = b[:, 4:6] + rs.uniform(1, 300, (nd, 2))
    b[:, 15] = rs.uniform(0, 1, nd)
    x = F.build_features(b, 8, '2d+temp', 'bdd100k', 'libra').cpu().numpy()
    mean, std = FO.norm_constants('bdd100k', 'libra', '2d+temp', 8)
    want = FO.build_features(b, 8, '2d+temp', mean, std)
    np.testing.assert_array_equal(x[:, :13], want[:, :13])
    np.testing.assert_allclose(x[:, 13:], want[:, 13:], atol=2e-7, rtol=0)
