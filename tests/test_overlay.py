"""The module overlay that lets the reference's drivers import this library under the reference's module names
(INTEGRATION.md section 3).  Import-level only: no compute, no GPU."""
import inspect
import sys

import pytest


@pytest.fixture
def overlay():
    import trackmpnn_b200.overlay as ov
    saved = {k: sys.modules.get(k) for k in list(ov._NAMES) + list(ov._OPTIONAL)}
    yield ov
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def test_reference_import_lines_resolve_to_this_library(overlay):
    overlay.install()
    # the import lines of the reference's infer.py:8-11 / train.py:13-16
    from models.track_mpnn import TrackMPNN
    from models.loss import FocalLoss, CELoss, create_targets
    from utils.graph import initialize_graph, update_graph, prune_graph, decode_tracks, hungarian
    import trackmpnn_b200.utils.graph as ours
    assert initialize_graph is ours.initialize_graph and decode_tracks is ours.decode_tracks
    assert 'use_hungraian' in inspect.signature(update_graph).parameters      # the reference's spelling is the API
    assert 'use_hungraian' in inspect.signature(decode_tracks).parameters
    assert [p for p in inspect.signature(TrackMPNN.__init__).parameters][1:6] == \
        ['features', 'ncategories', 'nhidden', 'nattheads', 'msg_type']
    assert callable(FocalLoss) and callable(CELoss) and callable(create_targets) and callable(prune_graph) and callable(hungarian)


def test_three_output_default_and_state_dict_names(overlay):
    overlay.install(three_outputs=True)
    from models.track_mpnn import TrackMPNN
    m = TrackMPNN('2d', 3, 64, 0, 'diff')
    assert m.return_attention is False          # infer.py:51 / train.py:68 unpack three values
    keys = set(m.state_dict())
    for k in ('input_transforms.0.0.weight', 'input_transforms.0.1.running_mean', 'input_transforms.0.3.bias',
              'factor_grus.0.edge_gru.weight_ih', 'factor_grus.0.node_gru.bias_hh', 'output_transform_node.weight',
              'output_transform_edge.bias'):
        assert k in keys                        # reference snapshots load with strict=True (infer.py:108)
    assert type(m).__name__ == 'TrackMPNN'
    overlay.install(three_outputs=False)
    import importlib
    tm = importlib.import_module('trackmpnn_b200.models.track_mpnn')
    assert tm.TrackMPNN('2d', 3, 64, 0, 'diff').return_attention is True


def test_missing_optional_modules_are_stubbed_loudly(overlay):
    sys.modules.pop('motmetrics', None)
    overlay.install(stub_missing=True)
    import motmetrics
    if getattr(motmetrics, '__file__', None) is None:   # stubbed (not installed in this image)
        with pytest.raises(ImportError):
            motmetrics.MOTAccumulator()


@pytest.mark.skipif(not __import__('os').path.isdir('/root/reference/dataset'), reason='reference tree not present')
def test_reference_dataset_module_imports_on_top_of_the_overlay(overlay):
    """With the reference tree on sys.path, its own dataset module (which imports models.loss.EmbeddingLoss /
    FairMOTLoss, losses of the out-of-scope embedding CNN) imports next to this library's hot-path modules."""
    sys.path.insert(0, '/root/reference')
    try:
        for k in [k for k in sys.modules if k == 'dataset' or k.startswith('dataset.')]:
            sys.modules.pop(k)
        overlay.install(stub_missing=True)
        from dataset.kitti_mot import KittiMOTDataset
        from models.loss import FairMOTLoss, CELoss
        from utils.graph import update_graph
        assert KittiMOTDataset.__module__ == 'dataset.kitti_mot'
        assert CELoss.__module__ == 'trackmpnn_b200.models.loss' and update_graph.__module__ == 'trackmpnn_b200.utils.graph'
        assert FairMOTLoss.__module__.startswith('_tmpnn_reference_')
    finally:
        sys.path.remove('/root/reference')
        for k in [k for k in sys.modules if k == 'dataset' or k.startswith('dataset.') or k.startswith('_tmpnn_reference_')]:
            sys.modules.pop(k)
