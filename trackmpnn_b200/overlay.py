"""Module overlay: lets the reference's own drivers (``infer.py``, ``train.py``, ``ablation.py``) run on this
library without edits.

    import trackmpnn_b200.overlay as overlay
    overlay.install()                       # before the driver is imported
    import runpy; runpy.run_path('/path/to/TrackMPNN/infer.py', run_name='__main__')

``install`` registers this package's modules under the names the reference imports -- ``models.track_mpnn``,
``models.layers``, ``models.loss``, ``utils.graph``, ``utils.metrics`` -- so that ``from models.track_mpnn import TrackMPNN`` and
``from utils.graph import initialize_graph, ...`` resolve here while everything else (datasets, option parsing,
metrics) still comes from the reference tree on ``sys.path``.  The reference's drivers unpack three values from
``model(...)`` although its module returns four; ``three_outputs=True`` (default) makes the
overlaid ``TrackMPNN`` default to ``return_attention=False`` for them.  ``stub_missing=True`` additionally provides
empty stand-ins for optional third-party imports the drivers make at import time but the hot path never calls
(``matplotlib`` plotting, ``motmetrics``, the un-vendored ``DCNv2`` extension), when those are not installed.
"""
import importlib
import importlib.util
import sys
import types

_NAMES = {
    'models.track_mpnn': 'trackmpnn_b200.models.track_mpnn',
    'models.layers': 'trackmpnn_b200.models.layers',
    'models.loss': 'trackmpnn_b200.models.loss',
    'utils.graph': 'trackmpnn_b200.utils.graph',
    'utils.metrics': 'trackmpnn_b200.metrics',   # create_mot_accumulator / calc_mot_metrics without motmetrics
}
_OPTIONAL = ('matplotlib', 'matplotlib.pyplot', 'motmetrics', 'models.dla.DCNv2', 'models.dla.DCNv2.dcn_v2')


def install(three_outputs=True, stub_missing=False):
    """Registers the overlay; returns the dict {reference module name: module}.  Idempotent."""
    out = {}
    for ref_name, ours in _NAMES.items():
        mod = importlib.import_module(ours)
        extra = _reference_leftovers(ref_name, mod)
        if extra:
            # names of the reference module that are NOT part of the hot path (e.g. models.loss.EmbeddingLoss /
            # FairMOTLoss of the visual-embedding CNN, which dataset/kitti_mot.py imports) keep coming from the
            # reference's own file; everything this library defines wins
            shim = types.ModuleType(ref_name)
            shim.__dict__.update(extra)
            shim.__dict__.update({k: v for k, v in mod.__dict__.items() if not k.startswith('__')})
            shim.__file__ = getattr(mod, '__file__', None)
            mod = shim
        sys.modules[ref_name] = mod
        out[ref_name] = mod
    if three_outputs:
        tm = out['models.track_mpnn']
        if not getattr(tm, '_tmpnn_overlay_three', False):
            base = tm.TrackMPNN

            class TrackMPNN(base):  # same name: state_dict keys and pickles are unaffected
                def __init__(self, features, ncategories, nhidden, nattheads, msg_type, return_attention=False, **kw):
                    super().__init__(features, ncategories, nhidden, nattheads, msg_type,
                                     return_attention=return_attention, **kw)

            TrackMPNN.__module__ = base.__module__
            TrackMPNN.__qualname__ = base.__qualname__
            shim = types.ModuleType('models.track_mpnn')
            shim.__dict__.update(tm.__dict__)
            shim.TrackMPNN = TrackMPNN
            shim._tmpnn_overlay_three = True
            sys.modules['models.track_mpnn'] = shim
            out['models.track_mpnn'] = shim
    if stub_missing:
        for name in _OPTIONAL:
            if name in sys.modules:
                continue
            try:
                found = importlib.util.find_spec(name) is not None
            except (ImportError, ValueError, AttributeError, TypeError):
                found = False
            if not found:
                m = types.ModuleType(name)
                m.__path__ = []   # a package: its stubbed submodules register themselves in sys.modules
                m.__getattr__ = lambda attr, _n=name: _missing(_n, attr)
                sys.modules[name] = m
    return out


def _reference_leftovers(ref_name, ours):
    """Public names the reference's own ``ref_name`` module defines and ``ours`` does not, loaded from the reference
    tree found on ``sys.path`` ({} when the tree is not there)."""
    import os
    rel = os.path.join(*ref_name.split('.')) + '.py'
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for root in sys.path:
        cand = os.path.join(root or '.', rel)
        if os.path.isfile(cand) and os.path.abspath(root or '.') != here:
            spec = importlib.util.spec_from_file_location('_tmpnn_reference_' + ref_name.replace('.', '_'), cand)
            ref = importlib.util.module_from_spec(spec)
            try:
                spec.loader.exec_module(ref)
            except Exception:
                return {}
            return {k: v for k, v in ref.__dict__.items() if not k.startswith('_') and not hasattr(ours, k)}
    return {}


def _missing(module, attr):
    if attr.startswith('__') and attr.endswith('__'):
        raise AttributeError(attr)

    def fail(*a, **k):
        raise ImportError(f'{module}.{attr} was stubbed by trackmpnn_b200.overlay (the package is not installed); '
                          'it is not part of the message-passing hot path')
    return fail


def uninstall():
    for ref_name in _NAMES:
        sys.modules.pop(ref_name, None)
