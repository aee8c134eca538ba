"""GPU: the whole chain around the hot path on files -- KITTI detection text files -> ``features.load_kitti_detections``
-> ``features.build_features`` (kernel) -> ``TrackEngine`` (all sequences in lock step) -> ``results.store_kitti_results``
-- against the oracle chain (``oracle/features_oracle`` + ``oracle/infer_loop``) on the same files: decoded track ids
bit-exact, result files identical."""
import os

import numpy as np
import pytest
import torch

from oracle import features_oracle as FO
from oracle.infer_loop import run_infer

TYPES = ['Pedestrian', 'Car', 'Cyclist']
CLASS_DICT = {'Pedestrian': 1, 'Car': 2, 'Cyclist': 3}
# seeds chosen on the oracle (test_chain_seeds_keep_a_margin below) so that no score sits within 1e-4 of the 0.5 threshold
SEEDS = [6, 9, 16, 18]


def detection_files(seed, frames=12):
    """{frame: [csv line]} of a few objects moving at constant velocity with jitter, misses and false positives."""
    rs = np.random.RandomState(1000 + seed)
    objs = [dict(t=TYPES[rs.randint(3)], x=rs.uniform(50, 1000), y=rs.uniform(120, 260), w=rs.uniform(30, 120),
                 h=rs.uniform(30, 100), vx=rs.uniform(-12, 12), vy=rs.uniform(-2, 2)) for _ in range(rs.randint(3, 7))]
    out = {}
    for fr in range(frames):
        lines = []
        for o in objs:
            if rs.rand() < 0.15:
                continue
            x, y = o['x'] + o['vx'] * fr + rs.normal(0, 1.5), o['y'] + o['vy'] * fr + rs.normal(0, 1.0)
            lines.append('%s,%.2f,%.2f,%.2f,%.2f,%.4f' % (o['t'], x, y, x + o['w'], y + o['h'], rs.uniform(0.5, 1.0)))
        if rs.rand() < 0.3:
            x, y = rs.uniform(0, 1100), rs.uniform(100, 300)
            lines.append('%s,%.2f,%.2f,%.2f,%.2f,%.4f' % (TYPES[rs.randint(3)], x, y, x + 40, y + 40, rs.uniform(0.3, 0.6)))
        out[fr] = lines
    return out


def _params(seed=5):
    """Decision-exercising weights (x20 on the matrices, edge-head bias 0), as the golden fixtures use."""
    from trackmpnn_b200.models.track_mpnn import TrackMPNN
    torch.manual_seed(seed)
    model = TrackMPNN('2d', 3, 64, 0, 'diff')
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() >= 2:
                p.mul_(20.0)
        model.output_transform_edge.bias.fill_(0.0)
    return model.eval()


def _oracle_chain(params, lines):
    bbox = FO.parse_kitti_detections(lines, 'All')
    mean, std = FO.norm_constants('kitti', 'centertrack', '2d', 3)
    X = FO.build_features(bbox, 3, '2d', mean, std)
    y = np.stack((bbox[:, 0], -np.ones(bbox.shape[0], np.float32)), 1)
    y_out, stats = run_infer(params, X, y, record_margin=True)
    return bbox, X, y_out, stats


def test_chain_seeds_keep_a_margin():
    params = {k: v.numpy() for k, v in _params().state_dict().items()}
    for sd in SEEDS:
        _, _, y_out, stats = _oracle_chain(params, detection_files(sd))
        assert stats['margin'] > 1e-4, (sd, stats['margin'])
        assert (y_out[:, 1] >= 0).any()


@pytest.mark.gpu
def test_detection_files_to_result_files(tmp_path):
    from trackmpnn_b200 import features as F, results as R
    from trackmpnn_b200.engine import TrackEngine
    dev = torch.device('cuda:0')
    model = _params().to(dev)
    params = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    seqs, boxes, wants = [], [], []
    for sd in SEEDS:
        lines = detection_files(sd)
        d = tmp_path / 'dets' / ('%04d' % sd)
        d.mkdir(parents=True)
        for fr, ls in lines.items():
            (d / ('%.4d.txt' % fr)).write_text(''.join(l + '\n' for l in ls))
        b = F.load_kitti_detections(str(tmp_path / 'dets'), '%04d' % sd, sorted(lines), 'All')
        X = F.build_features(b, 3, '2d', 'kitti', 'centertrack', device=dev)
        bbox_o, X_o, y_out_o, _ = _oracle_chain(params, lines)
        np.testing.assert_array_equal(b.numpy(), bbox_o)
        np.testing.assert_array_equal(X.cpu().numpy(), X_o)        # one-hot + 2d columns: same fp32 operations
        seqs.append((X.cpu().numpy(), np.stack((b.numpy()[:, 0], -np.ones(b.shape[0], np.float32)), 1)))
        boxes.append(b.numpy())
        wants.append(y_out_o)
    outs, stats = TrackEngine(model, seqs, cur_win_size=5, ret_win_size=0).run().results()
    for sd, b, got, want in zip(SEEDS, boxes, outs, wants):
        np.testing.assert_array_equal(got, want[:, 1])
        y_out = np.stack((b[:, 0].astype(np.int64), got), 1)
        R.store_kitti_results(b[:, 2:], y_out.copy(), CLASS_DICT, str(tmp_path / 'ours' / ('%04d.txt' % sd)))
        R.store_kitti_results(b[:, 2:], want.copy(), CLASS_DICT, str(tmp_path / 'oracle' / ('%04d.txt' % sd)))
        ours = (tmp_path / 'ours' / ('%04d.txt' % sd)).read_text()
        assert ours == (tmp_path / 'oracle' / ('%04d.txt' % sd)).read_text()
        assert len(ours.splitlines()) > 0
