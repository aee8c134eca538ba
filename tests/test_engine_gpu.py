"""GPU: the batched TrackEngine (device-resident, sync-free, CUDA-graph replayed) against the
oracle's inference loop, sequence by sequence: decoded track ids bit-exact, work counters equal."""
import numpy as np
import pytest
import torch

from oracle import trackmpnn_oracle as O
from oracle.infer_loop import run_infer
from trackmpnn_b200 import synth

pytestmark = pytest.mark.gpu


def _model(dev, dataset='kitti', msg_type='diff', scale=20.0, edge_bias=0.0, seed=5, heads=0):
    from trackmpnn_b200.models.track_mpnn import TrackMPNN
    torch.manual_seed(seed)
    model = TrackMPNN('2d', synth.num_categories(dataset), 64, heads, msg_type)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2 and '.gat.' not in name:
                p.mul_(scale)
        if edge_bias is not None:
            model.output_transform_edge.bias.fill_(edge_bias)
    return model.to(dev).eval()


def _params(model):
    return {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}


def _sequences(seeds, dataset='kitti', gap=()):
    seqs = []
    for i, sd in enumerate(seeds):
        T, D = 8 + (sd % 7), 4 + (sd % 4)
        ts = None
        if sd in gap:   # a hole longer than the window forces the re-initialisation path
            ts = list(range(0, 4)) + list(range(12, 12 + T))
        X, y = synth.make_sequence(sd, T, D, dataset, timestamps=ts)
        seqs.append((X[0], y[0]))
    return seqs


GAT_SEEDS = [108, 41, 52, 68, 49, 44, 105]


@pytest.mark.parametrize('cfg', [
    # seeds chosen (on the oracle) so that no score is within 5e-4 of the 0.5 decision threshold
    dict(seeds=[30, 34, 48, 58, 65], msg_type='diff', ret=0, graph=False, gap=()),
    dict(seeds=[72, 77, 78, 81, 84, 96], msg_type='diff', ret=0, graph=True, gap=()),
    dict(seeds=[35, 36, 40], msg_type='concat', ret=2, graph=True, gap=()),
    dict(seeds=[30, 49, 34, 52, 54], msg_type='diff', ret=0, graph=True, gap=(49, 52, 54)),
    dict(seeds=[71, 72], msg_type='diff', ret=0, graph=False, gap=(), stock=True),
    dict(seeds=[30, 34, 48, 58, 65], msg_type='diff', ret=0, graph=False, gap=(), tensor=True),
    dict(seeds=[30, 49, 34, 52, 54, 72, 77], msg_type='diff', ret=0, graph=True, gap=(49, 52, 54), tensor=True),
    dict(seeds=[30, 49, 34, 52, 54, 72, 77], msg_type='diff', ret=0, graph=True, gap=(49, 52, 54), tensor=True, kernel='gather'),
    dict(seeds=[35, 36, 40], msg_type='concat', ret=2, graph=True, gap=(), tensor=True),
    dict(seeds=[108, 84, 77, 100, 72, 32, 56], msg_type='diff', ret=0, graph=True, gap=(), hungarian=True),
    dict(seeds=GAT_SEEDS, msg_type='diff', ret=0, graph=True, gap=(), heads=2),
    dict(seeds=[84, 100, 48, 125, 56], msg_type='diff', ret=2, graph=False, gap=(), hungarian=True),
])
@pytest.mark.parametrize('deferred', [False, True])
def test_engine_matches_oracle(cfg, deferred):
    from trackmpnn_b200.engine import TrackEngine
    dev = torch.device('cuda:0')
    stock = cfg.get('stock', False)
    model = _model(dev, msg_type=cfg['msg_type'], scale=1.0 if stock else 20.0, edge_bias=None if stock else 0.0,
                   heads=cfg.get('heads', 0))
    params = _params(model)
    seqs = _sequences(cfg['seeds'], gap=cfg['gap'])
    eng = TrackEngine(model, seqs, cur_win_size=5, ret_win_size=cfg['ret'], use_cuda_graph=cfg['graph'],
                      tensor_cores=cfg.get('tensor', False), use_hungarian=cfg.get('hungarian', False),
                      deferred_compaction=deferred, tensor_kernel=cfg.get('kernel', 'auto'))
    outs, stats = eng.run().results()
    tot_e = tot_f = 0
    for (X, y), got in zip(seqs, outs):
        want, st = run_infer(params, X, y, msg_type=cfg['msg_type'], cur_win_size=5, ret_win_size=cfg['ret'],
                             record_margin=True, use_hungarian=cfg.get('hungarian', False))
        assert stock or st['margin'] > 1e-4, 'decision margin too small for a meaningful bit-exact comparison; change the seed'
        np.testing.assert_array_equal(got, want[:, 1])
        tot_e += st['edge_updates']; tot_f += st['frames']
    assert stats['edge_updates'] == tot_e
    assert stats['frames'] == tot_f
    # a second run on the same engine (CUDA graph re-used) gives the same answer
    outs2, stats2 = eng.run().results()
    for a, b in zip(outs, outs2):
        np.testing.assert_array_equal(a, b)
    assert stats2 == stats


@pytest.mark.parametrize('tensor', [False, True])
@pytest.mark.parametrize('ticks', [1, 2, 5, 8])
def test_deferred_compaction_is_bit_identical(tensor, ticks):
    """Skipping the physical move of the hidden states (the next step reads them through the position maps)
    changes nothing: graphs, scores and logits after any number of frames equal the plain engine's bit for bit,
    and so do the states once read through the maps."""
    from trackmpnn_b200.engine import TrackEngine
    dev = torch.device('cuda:0')
    model = _model(dev)
    seqs = _sequences([30, 49, 34, 52, 54, 72, 77], gap=(49, 52))
    a = TrackEngine(model, seqs, cur_win_size=5, ret_win_size=1, use_cuda_graph=False, tensor_cores=tensor,
                    deferred_compaction=False)
    b = TrackEngine(model, seqs, cur_win_size=5, ret_win_size=1, use_cuda_graph=False, tensor_cores=tensor,
                    deferred_compaction=True)
    a.run(max_ticks=ticks); b.run(max_ticks=ticks)
    torch.cuda.synchronize()
    a.ga.check_status(); b.ga.check_status()
    na, nb = a.ga.n_rows.cpu().numpy(), b.ga.n_rows.cpu().numpy()
    np.testing.assert_array_equal(na, nb)
    assert na.sum() > 0
    hb_buf = b.h_alt if (ticks & 1) else b.h_cur   # the buffer the last step wrote
    for s in range(len(seqs)):
        ra = slice(s * a.cap_rows, s * a.cap_rows + int(na[s]))
        rb = slice(s * b.cap_rows, s * b.cap_rows + int(nb[s]))
        for name in ('ts', 'det', 'ass', 'src', 'dst', 'score', 'logit'):
            np.testing.assert_array_equal(getattr(a.ga, name)[ra].cpu().numpy(), getattr(b.ga, name)[rb].cpu().numpy(), err_msg=name)
        ha = a.h_cur[ra].cpu().numpy()
        hb = hb_buf[b.ga.phys[rb].long()].cpu().numpy()
        np.testing.assert_array_equal(ha, hb)


@pytest.mark.parametrize('ret', [0, 2])
def test_structured_index_equals_general_index(ret):
    """The engine's sort-free index (incidence lists derived from the block boundaries) is entry for
    entry the general one (atomic fill + per-segment sort) on graphs in the middle of a run."""
    from trackmpnn_b200.engine import TrackEngine
    from trackmpnn_b200.device_graph import SlabIndex
    dev = torch.device('cuda:0')
    model = _model(dev)
    seqs = _sequences([30, 34, 48, 58, 65, 72, 77], gap=(48,))
    eng = TrackEngine(model, seqs, cur_win_size=5, ret_win_size=ret, use_cuda_graph=False)
    for ticks in (1, 3, 6, 9):
        eng.run(max_ticks=ticks)
        torch.cuda.synchronize()
        g = eng.ga
        a = SlabIndex(g, eng.index.cap_dets, eng.index.cap_inc)
        b = SlabIndex(g, eng.index.cap_dets, eng.index.cap_inc)
        a.build(g, None, structured=False)
        b.build(g, None, structured=True)
        torch.cuda.synchronize()
        g.check_status()
        nd = int(a.n_dets.item())
        assert nd == int(b.n_dets.item()) and nd > 0
        np.testing.assert_array_equal(a.det_rows[:nd].cpu().numpy(), b.det_rows[:nd].cpu().numpy())
        sa, sb = a.seg_ptr[:2 * nd + 1].cpu().numpy(), b.seg_ptr[:2 * nd + 1].cpu().numpy()
        np.testing.assert_array_equal(sa, sb)
        np.testing.assert_array_equal(a.inc[:sa[-1]].cpu().numpy(), b.inc[:sb[-1]].cpu().numpy())
        assert sa[-1] == 2 * int(a.n_edges.item())


@pytest.mark.parametrize('tensor', [False, True])
@pytest.mark.parametrize('msg_type,seeds', [('diff', [35, 40, 41, 43, 45, 52]), ('concat', [40, 44, 48, 61, 65])])
def test_engine_two_feature_groups(msg_type, seeds, tensor):
    """features '2d+temp' (two feature groups, models/track_mpnn.py:17-33): every group has its own input transform, GRU
    cells and aggregate, the heads read both groups' states.  Decoded tracks bit-exact against the oracle loop on the FMA
    path and on the tensor-core path (edge step + detection rows in detection mode, one aggregate buffer per group kept
    for the range-overflow re-run); seeds chosen on the oracle for a decision margin > 5e-4."""
    from trackmpnn_b200.engine import TrackEngine
    from trackmpnn_b200.models.track_mpnn import TrackMPNN
    dev = torch.device('cuda:0')
    torch.manual_seed(5)
    model = TrackMPNN('2d+temp', 3, 64, 0, msg_type)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2:
                p.mul_(20.0)
        model.output_transform_edge.bias.fill_(0.0)
    model = model.to(dev).eval()
    params = _params(model)
    seqs = []
    for sd in seeds:
        X, y = synth.make_sequence(sd, 8 + (sd % 7), 4 + (sd % 4), 'kitti')
        X, y = X[0], y[0]
        ph = 2 * np.pi * y[:, 0:1] / 10.0
        seqs.append((np.concatenate((X, np.sin(ph), np.cos(ph)), 1).astype(np.float32), y))
    eng = TrackEngine(model, seqs, cur_win_size=5, ret_win_size=0, use_cuda_graph=True, tensor_cores=tensor)
    assert eng.G == 2 and eng.det_tensor == tensor
    outs, stats = eng.run().results()
    for (X, y), got in zip(seqs, outs):
        want, st = run_infer(params, X, y, features='2d+temp', msg_type=msg_type, cur_win_size=5, ret_win_size=0,
                             record_margin=True)
        assert st['margin'] > 1e-4
        np.testing.assert_array_equal(got, want[:, 1])


@pytest.mark.parametrize('deferred', [False, True])
@pytest.mark.parametrize('cfg', [dict(win=5, ret=1, seqs=[(30, 12, 6), (49, 14, 9), (34, 10, 5), (52, 11, 7), (54, 16, 4)], gap=(49, 52)),
                                 dict(win=20, ret=2, stock=True, seqs=[(803, 30, 6), (804, 27, 5), (805, 24, 7)], gap=()),
                                 dict(win=8, ret=0, stock=True, seqs=[(811, 14, 40), (812, 12, 75)], gap=()),
                                 # blocks of ~210 columns (four column chunks) x ~420 sources (14 stripes)
                                 dict(win=3, ret=0, stock=True, seqs=[(821, 8, 210), (822, 7, 140)], gap=())])
def test_block_aggregation_equals_incidence_list_aggregation(cfg, deferred):
    """tmpnn_aggregate_dets_blocks (one pass over every dense edge block: row sums + per-stripe column partials, freshly
    appended blocks skipped) against tmpnn_aggregate_dets (incidence lists) on every step of a run: same sums up to fp32
    re-association (models/layers.py:103).  Blocks wider than one 64-column chunk and longer than one 32-source stripe
    are covered by the 40 / 75 / 140 / 210 detections-per-frame streams."""
    from trackmpnn_b200.engine import TrackEngine
    dev = torch.device('cuda:0')
    stock = cfg.get('stock', False)
    model = _model(dev, scale=1.0 if stock else 20.0, edge_bias=None if stock else 0.0)
    seqs = []
    for sd, T, D in cfg['seqs']:
        ts = (list(range(0, 4)) + list(range(4 + cfg['win'] + 1, 4 + cfg['win'] + 1 + T))) if sd in cfg['gap'] else None
        X, y = synth.make_sequence(sd, T, D, 'kitti', timestamps=ts)
        seqs.append((X[0], y[0]))
    eng = TrackEngine(model, seqs, cur_win_size=cfg['win'], ret_win_size=cfg['ret'], use_cuda_graph=False,
                      deferred_compaction=deferred, block_aggregation=True)
    assert eng._agg_blocks is not None
    eng.check_aggregation = []
    eng.run()
    eng.results()
    assert len(eng.check_aggregation) >= 6
    assert max(nd for nd, _, _ in eng.check_aggregation) > 20
    assert max(ref for _, _, ref in eng.check_aggregation) > 1e-3
    for nd, diff, ref in eng.check_aggregation:
        assert diff <= 2e-6 * max(1.0, ref), (nd, diff, ref)
    eng.check_aggregation = None
    # and the engine's tracks do not depend on the form of the aggregation (decisions margin-guarded by the seeds)
    ref_eng = TrackEngine(model, seqs, cur_win_size=cfg['win'], ret_win_size=cfg['ret'], use_cuda_graph=False,
                          deferred_compaction=deferred, block_aggregation=False)
    assert ref_eng._agg_blocks is None
    a, sa = eng.run().results()
    b, sb = ref_eng.run().results()
    assert sa == sb
    if stock:
        for x, y_ in zip(a, b):
            np.testing.assert_array_equal(x, y_)


@pytest.mark.parametrize('kernel', ['fma', 'gather', 'pre'])
def test_engine_state_at_workload_size(kernel):
    """BDD-shaped sequences at the bench workload's size (~80 detections / frame, ~60 k association rows per
    window, hundreds of 128-row tiles per launch, so every pipeline stage / barrier phase of the tensor-core
    kernels wraps many times): after 7 tracked frames the engine's window graphs equal the oracle's bit for bit and the
    scores / hidden states (read through the deferred-compaction maps) agree within 1e-4."""
    from trackmpnn_b200.engine import TrackEngine
    dev = torch.device('cuda:0')
    model = _model(dev, dataset='bdd', scale=1.0, edge_bias=None)   # stock init: BASELINE configs
    params = _params(model)
    seqs = []
    for sd in (11, 12, 13):
        X, y = synth.make_sequence(sd, 12, 80, 'bdd')
        seqs.append((X[0], y[0]))
    frames = 7
    ticks = 2 + frames   # engine ticks are absolute timesteps; the first two frames are consumed by the initialisation
    eng = TrackEngine(model, seqs, cur_win_size=5, ret_win_size=0, use_cuda_graph=False, tensor_cores=kernel != 'fma',
                      tensor_kernel=kernel if kernel != 'fma' else 'auto')
    eng.run(max_ticks=ticks)
    torch.cuda.synchronize()
    eng.ga.check_status()
    n_rows = eng.ga.n_rows.cpu().numpy()
    hb = eng.h_alt if (ticks & 1) else eng.h_cur   # the buffer the last step wrote
    for s, (X, y) in enumerate(seqs):
        _, st = run_infer(params, X, y, ncategories=synth.num_categories('bdd'), cur_win_size=5, max_frames=frames,
                          keep_state=True)
        g, h, sc = st['state']['g'], st['state']['h'], st['state']['scores']
        assert g.n > 20000 and int(n_rows[s]) == g.n
        rows = slice(s * eng.cap_rows, s * eng.cap_rows + g.n)
        for name in ('ts', 'det', 'ass', 'src', 'dst'):
            np.testing.assert_array_equal(getattr(eng.ga, name)[rows].cpu().numpy(), getattr(g, name), err_msg=name)
        np.testing.assert_allclose(eng.ga.score[rows].cpu().numpy(), sc[:, 1], atol=1e-4, rtol=0)
        np.testing.assert_allclose(hb[eng.ga.phys[rows].long()].cpu().numpy(), h, atol=1e-4, rtol=0)


def _edge_case_sequences():
    """Degenerate streams next to a normal one: a single timestep (initialize_graph has nothing to pair: no tracks), two
    timesteps (only the initial forward runs), holes shorter than the window, one detection per frame, no detection."""
    out = []
    for seed, frames, dets, kw in ((30, 10, 5, {}), (31, 1, 5, {}), (34, 2, 3, {}),
                                   (32, None, 4, dict(timestamps=[0, 1, 3, 4, 7, 8, 9, 12])),
                                   (33, 12, 1, dict(poisson=False, miss_rate=0.0, fp_rate=0.0))):
        X, y = synth.make_sequence(seed, frames, dets, 'kitti', **kw)
        out.append((X[0], y[0]))
    return out


@pytest.mark.parametrize('graph', [False, True])
def test_engine_edge_case_sequences(graph):
    from trackmpnn_b200.engine import TrackEngine
    dev = torch.device('cuda:0')
    model = _model(dev)
    params = _params(model)
    seqs = _edge_case_sequences()
    empty = (np.zeros((0, 8), np.float32), np.zeros((0, 2), np.float32))   # the reference raises IndexError on this one
    eng = TrackEngine(model, seqs[:2] + [empty] + seqs[2:], cur_win_size=5, ret_win_size=0, use_cuda_graph=graph)
    outs, stats = eng.run().results()
    assert outs[2].shape == (0,)
    outs = outs[:2] + outs[3:]
    tot_e = tot_f = 0
    for (X, y), got in zip(seqs, outs):
        want, st = run_infer(params, X, y, record_margin=True)
        assert st['margin'] > 1e-4
        np.testing.assert_array_equal(got, want[:, 1])
        tot_e += st['edge_updates']; tot_f += st['frames']
    assert (outs[1] == -1).all() and (outs[2] == -1).all()    # nothing to pair / never decoded
    assert stats['edge_updates'] == tot_e and stats['frames'] == tot_f


def test_engine_reports_capacity_overflow():
    """A slab too small for the window graph sets the sticky device flag; results() raises instead of returning tracks."""
    from trackmpnn_b200 import _lib as L
    from trackmpnn_b200.engine import TrackEngine
    dev = torch.device('cuda:0')
    model = _model(dev)
    seqs = _sequences([30, 34])
    eng = TrackEngine(model, seqs, cur_win_size=5, ret_win_size=0, use_cuda_graph=False, cap_rows=48)
    with pytest.raises(L.TmpnnError, match='capacity'):
        eng.run().results()


# (seed, frames, detections / frame) chosen on the oracle: no score within 5e-4 of the 0.5 decision threshold
WIDE = {
    'w8': dict(win=8, ret=1, seqs=[(219, 25, 3), (234, 22, 3), (240, 28, 3), (276, 28, 3), (297, 22, 3), (312, 28, 3)]),
    'w8_hungarian': dict(win=8, ret=2, hungarian=True, seqs=[(210, 25, 3), (219, 25, 3), (234, 22, 3), (240, 28, 3), (288, 22, 3), (393, 28, 3)]),
    # 486 / 507 are sparse streams with holes as long as the window: the graph empties exactly one frame before detections
    # return, where infer.py:64-69 re-initialises from the CURRENT frame (its test reads the previous iteration's feats)
    'w12': dict(win=12, ret=0, seqs=[(207, 22, 3), (210, 25, 3), (273, 25, 3), (367, 29, 4), (387, 22, 3), (402, 28, 3), (486, 26, 3), (507, 29, 3)]),
    'w20': dict(win=20, ret=3, seqs=[(240, 28, 3), (273, 25, 3), (284, 27, 5), (402, 28, 3), (477, 22, 3), (498, 25, 3), (486, 26, 3), (507, 29, 3)]),
    # stock init: nothing is ever associated, the window holds every detection of 20 frames and all pairs between them
    'w20_stock': dict(win=20, ret=0, stock=True, seqs=[(801, 30, 6), (802, 27, 5)]),
    'w20_stock_tensor': dict(win=20, ret=2, stock=True, tensor=True, seqs=[(803, 30, 6), (804, 27, 5), (805, 24, 7)]),
}


@pytest.mark.parametrize('name', sorted(WIDE))
@pytest.mark.parametrize('graph', [False, True])
def test_engine_wide_windows(name, graph):
    """--cur-win-size 8 / 12 / 20 (BASELINE configs[3] uses 20) with retention windows: decoded tracks bit-exact against the
    oracle loop, work counters equal (reference infer.py:82-87: t_upto = t_cur - cur_win_size + 2)."""
    from trackmpnn_b200.engine import TrackEngine
    cfg = WIDE[name]
    dev = torch.device('cuda:0')
    stock = cfg.get('stock', False)
    model = _model(dev, scale=1.0 if stock else 20.0, edge_bias=None if stock else 0.0)
    params = _params(model)
    seqs = []
    for sd, T, D in cfg['seqs']:
        X, y = synth.make_sequence(sd, T, D, 'kitti')
        seqs.append((X[0], y[0]))
    eng = TrackEngine(model, seqs, cur_win_size=cfg['win'], ret_win_size=cfg['ret'], use_cuda_graph=graph,
                      tensor_cores=cfg.get('tensor', False), use_hungarian=cfg.get('hungarian', False))
    outs, stats = eng.run().results()
    tot_e = tot_f = 0
    for (X, y), got in zip(seqs, outs):
        want, st = run_infer(params, X, y, cur_win_size=cfg['win'], ret_win_size=cfg['ret'], record_margin=True,
                             use_hungarian=cfg.get('hungarian', False))
        assert stock or st['margin'] > 1e-4, 'decision margin too small for a meaningful bit-exact comparison; change the seed'
        np.testing.assert_array_equal(got, want[:, 1])
        tot_e += st['edge_updates']; tot_f += st['frames']
    assert stats['edge_updates'] == tot_e and stats['frames'] == tot_f


@pytest.mark.parametrize('n_frame', [1500, 3000])
def test_decode_walk_beyond_shared_memory(n_frame):
    """Windows with more detection rows than the walk's shared-memory arrays hold (4096 by default, 8192 at most): 3 frames of
    1500 (4500 rows: the large shared-memory launch) / 3000 (9000 rows: global-memory scratch) detections joined by random
    association rows; decode_tracks through the drop-in API against the oracle: track ids and the surviving graph bit-exact."""
    from trackmpnn_b200.device_graph import WindowGraph
    from trackmpnn_b200.utils.graph import decode_tracks
    dev = torch.device('cuda:0')
    rs = np.random.RandomState(n_frame)
    n_e = 700
    ts, det, src, dst = [], [], [], []
    row0 = {}
    n = 0
    for t in range(3):
        if t > 0:   # association rows into frame t, sorted by (source, target) like the reference's blocks
            s_rows = np.concatenate([row0[u] + np.arange(n_frame) for u in range(t)])
            pairs = sorted(set(zip(rs.choice(s_rows, n_e).tolist(), rs.randint(n_frame, size=n_e).tolist())))
            tgt0 = n + len(pairs)
            for a, j in pairs:
                ts.append(-1); det.append(-1); src.append(a); dst.append(tgt0 + j)
            n += len(pairs)
        row0[t] = n
        for j in range(n_frame):
            ts.append(t); det.append(t * n_frame + j); src.append(-1); dst.append(-1)
        n += n_frame
    ts, det, src, dst = (np.asarray(v, np.int64) for v in (ts, det, src, dst))
    p = rs.uniform(0.0, 1.0, n).astype(np.float32)
    p[np.abs(p - 0.5) < 1e-3] = 0.9
    scores = np.stack((1 - p, p), 1).astype(np.float32)
    states = rs.normal(size=(n, 64)).astype(np.float32)
    wg = WindowGraph(n, n, dev, with_labels=True)
    for name, v in (('ts', ts), ('det', det), ('src', src), ('dst', dst)):
        getattr(wg.g, name)[:n] = torch.from_numpy(v.astype(np.int32)).to(dev)
    wg.g.ass[:n] = -1
    y_out = np.stack((np.repeat(np.arange(3), n_frame), np.full(3 * n_frame, -1)), 1).astype(np.int64)
    y_out_o = y_out.copy()
    g = O.Graph(ts, det, np.full(n, -1, np.int64), src, dst, np.zeros(n, np.int64))
    g_o, y_out_o, h_o, sc_o, _ = O.decode_tracks(g, states, scores, y_out_o, 2, 1, False)
    y_pred, y_out, h, node_adj, labels, sc = decode_tracks(
        torch.from_numpy(states).to(dev), wg.adjacency(False), wg.labels(), torch.from_numpy(scores).to(dev), wg.y_pred(),
        y_out, 2, 1, use_hungraian=False, cuda=True)
    np.testing.assert_array_equal(y_out, y_out_o)
    assert (np.bincount(y_out[:, 1][y_out[:, 1] >= 0]) > 1).sum() > 50   # chains were walked
    np.testing.assert_array_equal(y_pred.cpu().numpy(), g_o.y_pred())
    np.testing.assert_array_equal(h.cpu().numpy(), h_o)


@pytest.mark.parametrize('graph', [False, True])
def test_engine_no_tp_classifier(graph):
    """infer.py's --no-tp-classifier (infer.py:54-57, 77-80): detection scores forced to (0, 1) before association and
    decoding.  The node head's bias is set to 0 so that the classifier rejects detections: every one of these streams
    decodes differently with and without it (checked on the oracle)."""
    from trackmpnn_b200.engine import TrackEngine
    dev = torch.device('cuda:0')
    model = _model(dev)
    with torch.no_grad():
        model.output_transform_node.bias.fill_(0.0)
    params = _params(model)
    seqs = _sequences([34, 48, 58, 65, 72, 77, 78, 81, 84])
    eng = TrackEngine(model, seqs, cur_win_size=5, ret_win_size=1, use_cuda_graph=graph, tp_classifier=False)
    outs, stats = eng.run().results()
    differs = 0
    for (X, y), got in zip(seqs, outs):
        want, st = run_infer(params, X, y, cur_win_size=5, ret_win_size=1, record_margin=True, tp_classifier=False)
        assert st['margin'] > 1e-4
        np.testing.assert_array_equal(got, want[:, 1])
        with_tp, _ = run_infer(params, X, y, cur_win_size=5, ret_win_size=1)
        differs += int((with_tp[:, 1] != want[:, 1]).any())
    assert differs >= 5


def test_engine_recaptures_after_weight_change():
    """A captured CUDA graph bakes in the pointers of the packed weight images; after an in-place parameter update
    (optimizer.step, load_state_dict) the engine must re-capture instead of replaying stale images: a re-used engine
    equals a fresh one bit for bit, before and after the change."""
    from trackmpnn_b200.engine import TrackEngine
    dev = torch.device('cuda:0')
    model = _model(dev)
    seqs = _sequences([30, 34, 48, 58, 65, 72, 77])
    eng = TrackEngine(model, seqs, cur_win_size=5, ret_win_size=0, use_cuda_graph=True)
    outs0, stats0 = eng.run().results()
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() >= 2:
                p.mul_(0.8)
        model.output_transform_edge.bias.add_(0.3)
    outs1, stats1 = eng.run().results()
    fresh = TrackEngine(model, seqs, cur_win_size=5, ret_win_size=0, use_cuda_graph=False)
    outs2, stats2 = fresh.run().results()
    assert stats1 == stats2
    for a, b in zip(outs1, outs2):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(eng.ga.score.cpu().numpy()[:64], fresh.ga.score.cpu().numpy()[:64])
    assert any((a != b).any() for a, b in zip(outs0, outs1)) or stats0 != stats1, 'the weight change should change the tracking'
    # load_state_dict copies in place as well
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        model.output_transform_edge.bias.add_(-0.3)
    eng.run().results()
    model.load_state_dict(sd)
    outs3, stats3 = eng.run().results()
    assert stats3 == stats1
    for a, b in zip(outs1, outs3):
        np.testing.assert_array_equal(a, b)


def test_tensor_core_range_overflow_reruns_on_fma():
    """States beyond the fp16 split's range (|h| > 6e4): the tensor-core step raises TMPNN_FLAG_TC_RANGE on the device and
    the conditional FMA launch behind it re-runs the association rows; the result equals the FMA engine's and results()
    reports a note instead of raising."""
    from trackmpnn_b200 import _lib as L
    from trackmpnn_b200.engine import TrackEngine
    dev = torch.device('cuda:0')
    model = _model(dev, scale=1.0, edge_bias=None)
    with torch.no_grad():   # a huge second Linear of the input transform: detection states of ~1e5
        model.input_transforms[0][3].weight.mul_(3e7)
    seqs = _sequences([30, 34, 48, 58])
    a = TrackEngine(model, seqs, cur_win_size=5, ret_win_size=0, use_cuda_graph=False, tensor_cores=True)
    b = TrackEngine(model, seqs, cur_win_size=5, ret_win_size=0, use_cuda_graph=False, tensor_cores=False)
    a.run(max_ticks=2); b.run(max_ticks=2)
    outs_a, _ = a.results()
    b.results()
    assert a.ga.notes & L.NOTE_TC_RANGE_RERUN and not (b.ga.notes & L.NOTE_TC_RANGE_RERUN)
    assert float(torch.maximum(a.h_cur.abs().max(), a.h_alt.abs().max())) > 6e4
    n = int(a.ga.n_rows[0])
    np.testing.assert_array_equal(a.ga.logit[:n].cpu().numpy(), b.ga.logit[:n].cpu().numpy())
    np.testing.assert_array_equal(a.h_cur[a.ga.phys[:n].long()].cpu().numpy(), b.h_cur[b.ga.phys[:n].long()].cpu().numpy())


@pytest.mark.parametrize('gap', [4, 5, 6, 7])
@pytest.mark.parametrize('graph', [False, True])
def test_engine_reinitialises_like_infer_py(gap, graph):
    """Holes of window - 1 ... window + 2 frames (window 5): infer.py:64-69 re-initialises when the graph its last decode left
    AND the rows its last update added are both empty -- with a hole exactly as long as the window that happens on a frame
    that HAS detections (initialize_graph from the current frame), one frame later on an empty one."""
    from trackmpnn_b200.engine import TrackEngine
    dev = torch.device('cuda:0')
    model = _model(dev)
    params = _params(model)
    seeds = {4: [900, 901, 902], 5: [901, 902, 903], 6: [900, 902, 904], 7: [900, 902, 903]}[gap]
    ts = [0, 1, 2, 3] + list(range(3 + gap, 3 + gap + 6))
    seqs = []
    for sd in seeds:
        X, y = synth.make_sequence(sd, None, 4, 'kitti', timestamps=ts)
        seqs.append((X[0], y[0]))
    outs, stats = TrackEngine(model, seqs, cur_win_size=5, ret_win_size=0, use_cuda_graph=graph).run().results()
    tot_e = tot_f = 0
    for (X, y), got in zip(seqs, outs):
        want, st = run_infer(params, X, y, cur_win_size=5, ret_win_size=0, record_margin=True)
        assert st['margin'] > 1e-4
        np.testing.assert_array_equal(got, want[:, 1])
        tot_e += st['edge_updates']; tot_f += st['frames']
    assert stats['edge_updates'] == tot_e and stats['frames'] == tot_f


@pytest.mark.parametrize('kernel', ['fma', 'gather', 'pre'])
def test_step_at_workload_size_with_decision_exercising_weights(kernel):
    """Workload-size window graphs (BDD shape, ~80 detections / frame, > 20 k association rows per sequence) with the
    x20 "decision-exercising" weights (the weight set where a single-pass reduced-precision GEMM fails the 1e-4
    bar by 100x).  At this size some score always sits within 1e-5 of the 0.5 threshold, so a free-running comparison cannot be
    margin-guarded; instead the graph is grown with stock weights (decision free, checked elsewhere), then the weights are
    scaled in place and three message-passing steps are run on the frozen graph (infer.py's forward with no new rows), each
    compared with the oracle's forward on the same graph and input state: states and logits within 1e-4."""
    from trackmpnn_b200 import _lib as L
    from trackmpnn_b200.engine import TrackEngine
    dev = torch.device('cuda:0')
    model = _model(dev, dataset='bdd', scale=1.0, edge_bias=None)
    seqs = []
    for sd in (11, 12, 13):
        X, y = synth.make_sequence(sd, 12, 80, 'bdd')
        seqs.append((X[0], y[0]))
    ticks = 9
    eng = TrackEngine(model, seqs, cur_win_size=5, ret_win_size=0, use_cuda_graph=False, tensor_cores=kernel != 'fma',
                      tensor_kernel=kernel if kernel != 'fma' else 'auto')
    eng.run(max_ticks=ticks)
    torch.cuda.synchronize()
    eng.ga.check_status()
    g = eng.ga
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2:
                p.mul_(20.0)
        model.output_transform_edge.bias.fill_(0.0)
    params = _params(model)
    h_in, h_out = (eng.h_alt, eng.h_cur) if (ticks & 1) else (eng.h_cur, eng.h_alt)   # h_in: what the last step wrote
    n_rows = g.n_rows.cpu().numpy()
    eng.n_new.zero_()    # no new rows: the forward is a pure message-passing step
    worst_h = worst_l = 0.0
    for step in range(3):
        refs = []
        for s in range(len(seqs)):
            rows = slice(s * eng.cap_rows, s * eng.cap_rows + int(n_rows[s]))
            og = O.Graph(*[getattr(g, k)[rows].cpu().numpy().astype(np.int64) for k in ('ts', 'det', 'ass', 'src', 'dst')])
            assert og.n > 20000
            hs = h_in[g.phys[rows].long()].cpu().numpy()
            refs.append(O.forward(params, np.zeros((0, 13), np.float32), hs, og, ncategories=8))
        eng._forward(g, h_in, h_out)
        torch.cuda.synchronize()
        g.check_status()
        assert not (g.notes & L.NOTE_TC_RANGE_RERUN)
        for s, (so, lo, ho) in enumerate(refs):
            rows = slice(s * eng.cap_rows, s * eng.cap_rows + int(n_rows[s]))
            dh = float(np.abs(h_out[rows].cpu().numpy() - ho).max())
            dl = float(np.abs(g.logit[rows].cpu().numpy() - lo[:, 0]).max())
            worst_h, worst_l = max(worst_h, dh), max(worst_l, dl)
            assert float(np.abs(ho).max()) > (0.05 if step else 0.0)   # the states are not tiny any more
        L.call('tmpnn_graph_phys_identity', g.c, L.stream())   # the state now sits at the logical rows of h_out
        h_in, h_out = h_out, h_in
    assert worst_h <= 1e-4 and worst_l <= 1e-4, (worst_h, worst_l)


@pytest.mark.parametrize('kernel', ['gather', 'pre'])
def test_tensor_core_split_keeps_precision_for_small_weights(kernel):
    """fp16 hi / lo split of the weights: the image is pre-scaled by a power of two (k_pack_gru_tc), so the residuals of small
    weights are normal fp16 numbers.  GRU weights of sigma = 1e-4 (100x below the stock init: without the scale every weight
    itself is an fp16 subnormal, ~8 significant bits, and the residual is zero) against detection states of O(10) (input
    transform x 1000), so that the gate pre-activations are O(0.01 - 0.1): the tensor-core step agrees with the fp32 FMA step
    to 3e-6 absolute (the ex2 / rcp gate formulation's own floor is ~3e-7) -- an 8-bit weight would be off by ~1e-4."""
    from trackmpnn_b200.engine import TrackEngine
    dev = torch.device('cuda:0')
    model = _model(dev, scale=1.0, edge_bias=None)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if 'factor_grus' in name and p.dim() >= 2:
                p.mul_(0.01)
        model.input_transforms[0][3].weight.mul_(1000.0)
    seqs = _sequences([30, 34, 48, 58, 65])
    a = TrackEngine(model, seqs, cur_win_size=5, ret_win_size=0, use_cuda_graph=False, tensor_cores=True, tensor_kernel=kernel)
    b = TrackEngine(model, seqs, cur_win_size=5, ret_win_size=0, use_cuda_graph=False, tensor_cores=False)
    a.run(max_ticks=3); b.run(max_ticks=3)
    a.results(); b.results()
    n = a.ga.n_rows.cpu().numpy()
    np.testing.assert_array_equal(n, b.ga.n_rows.cpu().numpy())
    n_edge, scale, err = 0, 0.0, 0.0
    for s in range(len(seqs)):
        rows = slice(s * a.cap_rows, s * a.cap_rows + int(n[s]))
        edge = (a.ga.ts[rows] < 0).cpu().numpy()
        ha = a.h_alt[a.ga.phys[rows].long()].cpu().numpy()[edge]     # 3 ticks: the last step wrote h_alt
        hb = b.h_alt[b.ga.phys[rows].long()].cpu().numpy()[edge]
        n_edge += int(edge.sum())
        scale = max(scale, float(np.abs(hb).max()))
        err = max(err, float(np.abs(ha - hb).max()))
    assert n_edge > 100 and scale > 3e-3 and err <= 3e-6, (n_edge, scale, err)
