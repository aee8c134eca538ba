"""GPU: the reference's error conventions and degenerate returns at the drop-in boundary:
``AssertionError`` for batch != 1 (``utils/graph.py:117,212``), for a track id used twice in a frame (``:178``) and for
``t_st > t_ed`` (``:359``); ``initialize_graph`` answers 7 x ``None`` when there is nothing to pair (``:132-133``);
``update_graph`` on a timestep without detections leaves the graph as it is and returns zero feature rows (``:284,295``)."""
import numpy as np
import pytest
import torch

from trackmpnn_b200 import synth

pytestmark = pytest.mark.gpu


def _stream(seed=30, frames=8, dets=4, timestamps=None):
    X, y = synth.make_sequence(seed, frames, dets, 'kitti', timestamps=timestamps)
    dev = torch.device('cuda:0')
    return torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev)


def test_batch_size_must_be_one():
    from trackmpnn_b200.utils.graph import initialize_graph
    X, y = _stream()
    with pytest.raises(AssertionError, match='batch size 1'):
        initialize_graph(torch.cat((X, X)), torch.cat((y, y)), 0, 'test', True)


def test_nothing_to_pair_returns_nones():
    from trackmpnn_b200.utils.graph import initialize_graph
    X, y = _stream(frames=1)
    assert initialize_graph(X, y, 0, 'test', True) == (None,) * 7          # a single timestep
    X, y = _stream()
    assert initialize_graph(X, y, int(y[0, :, 0].max()), 'test', True) == (None,) * 7   # only the last timestep is left
    y2 = y.clone()
    y2[0, :, 1] = -1
    assert initialize_graph(X, y2, 0, 'train', True) == (None,) * 7        # training chunk without any labelled track
    assert initialize_graph(X, y2, 0, 'test', True)[0] is not None


def test_duplicate_track_id_in_a_frame_asserts():
    from trackmpnn_b200.utils.graph import initialize_graph
    X, y = _stream()
    first = torch.nonzero(y[0, :, 0] == y[0, 0, 0])[:, 0]
    assert first.numel() >= 2
    y[0, first[:2], 1] = 7
    with pytest.raises(AssertionError):
        initialize_graph(X, y, 0, 'train', True)


def test_update_on_an_empty_timestep_changes_nothing():
    from trackmpnn_b200.utils.graph import initialize_graph, update_graph
    X, y = _stream(timestamps=[0, 1, 3, 4, 5])                              # no detections at t = 2
    y_pred, feats, node_adj, edge_adj, labels, t_st, t_end = initialize_graph(X, y, 0, 'test', True)
    assert t_st == 2
    scores = torch.full((y_pred.shape[0], 2), 0.5, device=y_pred.device)
    y_pred2, feats2, node_adj2, edge_adj2, labels2 = update_graph(node_adj, labels, scores, y_pred, X, y, 2,
                                                                  use_hungraian=False, mode='test', cuda=True)
    assert feats2.shape == (0, X.shape[2])
    np.testing.assert_array_equal(y_pred2[:, :2].cpu().numpy(), y_pred[:, :2].cpu().numpy())
    np.testing.assert_array_equal(node_adj2.to_dense().cpu().numpy(), node_adj.to_dense().cpu().numpy())


def test_prune_window_must_be_ordered():
    from trackmpnn_b200.utils.graph import initialize_graph, prune_graph
    X, y = _stream()
    y_pred, feats, node_adj, edge_adj, labels, t_st, t_end = initialize_graph(X, y, 0, 'test', True)
    n = y_pred.shape[0]
    states = torch.zeros((n, 64), device=y_pred.device)
    scores = torch.full((n, 2), 0.5, device=y_pred.device)
    with pytest.raises(AssertionError, match='t_st'):
        prune_graph(states, node_adj, labels, scores, y_pred, 3, 1, 0.5, True)
