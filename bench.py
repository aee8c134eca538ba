#!/usr/bin/env python
"""bench.py -- TrackMPNN hot path on B200: edge-updates/s and tracked frames/s.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path
    python bench.py --impl reference [--gpus N] ...                # CPU baseline arm (oracle port)

Workload (BASELINE.json configs[2], "C3"): BDD100K-shaped synthetic detection streams
(8 categories, F = 13, ~Poisson(80) detections / frame, --cur-win-size 5, greedy decode,
stock random-init weights, torch.manual_seed(5)), 256 independent sequences PER GPU
(weak scaling: sequences shard across ranks with no communication).  One "step" = one
full tracking pass over the rank's batch: initialise, then for every frame
update_graph -> TrackMPNN.forward -> decode_tracks (reference infer.py:48-87).

metric: edge_updates_per_s = sum over forward calls of the edge rows in the graph / time
(SURVEY.md section 8d); frames_per_s is reported beside it.  `value` is measured with the
inputs resident in HBM; `e2e` re-uploads the detections from pinned host memory and reads
the decoded tracks back every step.  `roofline` / `roofline_aggregation` / `roofline_compaction` time the fused edge
step, the detection aggregation and the window slide launch by launch with CUDA events.  Side legs (N = 1): `c1`
(configs[0]: one KITTI-shaped sequence through the drop-in loop and the S = 1 engine), `train` (configs[1]: training
chunks through the drop-in modules, and 32 at a time through the batched trainer), `cpu_baseline` (oracle port);
N > 1 adds `train_ddp` (configs[4]: data-parallel batched training, one NCCL gradient all-reduce per step).
configs[3] (stress graph) is the same program with --win 20 --dets 180 --seqs-per-gpu 4.  See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_EDGE_UPDATE = 536   # h read 256 + h' write 256 + logit/score 8 + src/dst 8 + 2 incidences 8
FLOP_PER_ROW_UPDATE = 49.5e3  # two 192x64 GEMVs + gates + head


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--seqs-per-gpu', type=int, default=256)
    ap.add_argument('--frames', type=int, default=200)
    ap.add_argument('--dets', type=int, default=80)
    ap.add_argument('--dataset', default='bdd', choices=['bdd', 'kitti'])
    ap.add_argument('--win', type=int, default=5)
    ap.add_argument('--no-cuda-graph', action='store_true')
    ap.add_argument('--no-deferred', action='store_true', help='move the hidden states in every window slide (A/B switch)')
    ap.add_argument('--list-aggregation', action='store_true', help='aggregate through the incidence lists (A/B switch)')
    ap.add_argument('--fma-dets', action='store_true', help='detection rows on the fp32 FMA kernel (A/B switch)')
    ap.add_argument('--cpu-frames', type=int, default=0, help='frames of the CPU-baseline sample (0 = auto)')
    ap.add_argument('--skip-cpu', action='store_true')
    ap.add_argument('--skip-e2e', action='store_true')
    ap.add_argument('--skip-train', action='store_true')
    ap.add_argument('--train-chunks', type=int, default=6, help='BPTT chunks timed by the training leg')
    ap.add_argument('--train-dets', type=int, default=40)
    ap.add_argument('--train-batch', type=int, default=32, help='chunks per batch of the batched trainer')
    ap.add_argument('--skip-strong', action='store_true', help='N > 1: skip the strong-scaling leg')
    ap.add_argument('--skip-check', action='store_true', help='skip the same-tracks-on-every-rank check')
    ap.add_argument('--skip-c4', action='store_true', help='N = 1: skip the configs[3] stress-graph leg')
    ap.add_argument('--c4-win', type=int, default=20)
    ap.add_argument('--c4-dets', type=int, default=200)
    ap.add_argument('--c4-seqs', type=int, default=4)
    ap.add_argument('--c4-frames', type=int, default=26)
    return ap.parse_args()


def workload_name(a):
    return (f'C3 {a.dataset}-shaped synthetic streams, F={8 + 5 if a.dataset == "bdd" else 3 + 5}, ~Poisson({a.dets}) dets/frame, '
            f'{a.frames} frames, win {a.win}, greedy decode, stock init, {a.seqs_per_gpu} sequences per GPU')


def make_sequences(a, seed0, count):
    from trackmpnn_b200 import synth
    seqs = []
    for i in range(count):
        X, y = synth.make_sequence(seed0 + i, a.frames, a.dets, a.dataset)
        seqs.append((X[0], y[0]))
    return seqs


def measured_traffic(kernel):
    """DRAM bytes per association row of `kernel` from this round's ncu capture of the shipped library
    (profiles/r02_dram_traffic.json, written by profiles/make_dram_traffic.py from `ncu --metrics dram__bytes_*`;
    the file names the kernel, the command and the git revision it was taken at).  Returns (bytes per row, source)."""
    p = os.path.join(ROOT, 'profiles', 'r02_dram_traffic.json')
    try:
        with open(p) as f:
            d = json.load(f)
        k = d[kernel]
        return float(k['dram_bytes_per_edge_row']), (f"profiles/r02_dram_traffic.json: ncu dram__bytes_read+write of {k.get('kernel', kernel)} "
                                                      f"per association row x rows per launch (captured at {d.get('git', '?')})")
    except Exception:
        return None, None


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def mark(self):
        """Start of the timed region: nvidia-smi was started a little earlier so that its start-up latency does not
        eat a short region; only samples taken from here on count (all of them if the region was shorter than one
        sampling period)."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        t_mark = getattr(self, 't_mark', 0.0)
        rows = [r for t, r in self.rows if t >= t_mark] or [r for _, r in self.rows[-2:]]
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for nme, v in zip(names, r[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(nme)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def cpu_oracle_sample(a, frames_cap, seed=5):
    """One sequence of the workload through the oracle's inference loop, first `frames_cap` frames."""
    from oracle import trackmpnn_oracle as O
    from oracle.infer_loop import run_infer
    from trackmpnn_b200 import synth
    ncat = synth.num_categories(a.dataset)
    params = O.init_params('2d', ncat, 64, 'diff', seed=5)
    X, y = synth.make_sequence(seed, a.frames, a.dets, a.dataset)
    t0 = time.perf_counter()
    _, st = run_infer(params, X, y, ncategories=ncat, cur_win_size=a.win, max_frames=frames_cap)
    dt = time.perf_counter() - t0
    return st, dt


def run_reference(a, rank):
    """CPU baseline arm.  The reference is Python and does not travel to the GPU box, so this times the
    oracle port (edge-list numpy restatement, validated against the reference's golden vectors) on the
    host cores: numpy BLAS threads for the GEMMs, one thread for the graph bookkeeping."""
    if rank != 0:
        return
    frames_cap = a.cpu_frames or 60
    for _ in range(a.warmup):
        cpu_oracle_sample(a, 2)
    edges = frames = 0
    t = 0.0
    for k in range(a.steps):
        st, dt = cpu_oracle_sample(a, frames_cap, seed=5 + k)
        edges += st['edge_updates']; frames += st['frames']; t += dt
    val = edges / t
    sample = f'1 sequence of the workload, first {frames_cap} frames per step (oracle port, numpy)'
    out = {'impl': 'reference', 'metric': 'edge_updates_per_s', 'value': val, 'unit': 'edge-updates/s',
           'frames_per_s': frames / t, 'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup,
           'ms_per_step': 1e3 * t / max(1, a.steps), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
           'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': workload_name(a), 'sample': sample},
           'cpu_baseline': {'value': val, 'unit': 'edge-updates/s', 'cores': os.cpu_count(), 'kind': 'port', 'sample': sample},
           'e2e': {'value': val, 'unit': 'edge-updates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
           'gpu_launches': 0}
    print(json.dumps(out), flush=True)


# ---- training leg: BASELINE.json configs[1] (KITTI-shaped --category=All, ~40 dets/frame, forward + backward) ----
def train_chunk_cuda(model, opt, X, y):
    """One chunk of the reference's train.py:65-134 through the drop-in modules (tp_classifier on):
    teacher-forced graph growth, TrackMPNN.forward per step (autograd Function), CE + BCE, backward, Adam."""
    import torch
    from trackmpnn_b200.utils.graph import initialize_graph, update_graph
    from trackmpnn_b200.models.loss import create_targets, CELoss, FocalLoss
    ce, fn, fe = CELoss(), FocalLoss(gamma=0), FocalLoss(gamma=0)
    edges = 0

    def losses(scores, logits, y_pred, labels, node_adj):
        nonlocal edges
        idx_edge = torch.nonzero((y_pred[:, 0] == -1))[:, 0]
        idx_node = torch.nonzero((y_pred[:, 0] != -1))[:, 0]
        edges += int(idx_edge.numel())
        targets = create_targets(labels, node_adj, idx_node)
        lc = ce(logits, targets, node_adj, idx_node)
        lf = fn(scores[idx_node, 0], targets[idx_node]) + fe(scores[idx_edge, 0], targets[idx_edge])
        return torch.cat((1 - scores, scores), dim=1), lc + lf

    opt.zero_grad()
    y_pred, feats, node_adj, edge_adj, labels, t_st, t_end = initialize_graph(X, y, t_st=0, mode='train', cuda=True)
    scores, logits, states, _ = model(feats, None, node_adj, edge_adj)
    scores, loss = losses(scores, logits, y_pred, labels, node_adj)
    for t_cur in range(t_st, t_end):
        y_pred, feats, node_adj, edge_adj, labels = update_graph(node_adj, labels, scores, y_pred, X, y, t_cur,
                                                                 use_hungraian=False, mode='train', cuda=True)
        scores, logits, states, _ = model(feats, states, node_adj, edge_adj)
        scores, l = losses(scores, logits, y_pred, labels, node_adj)
        loss = loss + l
    loss.backward()
    opt.step()
    return edges, float(loss.item())


def run_train_batched(a, dev, model, chunk, steps=9):
    """The same training step for B chunks at once (trackmpnn_b200/train_engine.py): per message-passing step ONE
    block-diagonal graph, so every kernel runs once for the whole batch; BatchNorm statistics and the BCE means stay
    per chunk, the batch loss is the sum of the chunk losses (verified against the chunk-by-chunk path in
    tests/test_train_engine_gpu.py and against oracle/train_ref.py at this size).  Forward of the association rows and the two
    backward contractions run on tcgen05; the backward kernels accumulate into one flat gradient buffer; the whole optimizer
    step (zero, forward, CE + BCE, backward, Adam) is replayed as CUDA graphs (GraphedTrainStep).  The graphs depend on the
    labels only (teacher forcing), so a batch is built once and replayed; `value` times the optimizer steps,
    `value_incl_graph_build` charges a full rebuild of the batch's graphs to every step, which is what the reference's loop does
    (train.py:92-104)."""
    import torch
    from trackmpnn_b200.train_engine import GraphedTrainStep, TrainBatch
    B = a.train_batch
    chunks = []
    for i in range(B):
        Xn, yn = chunk(3000 + i)
        chunks.append((torch.from_numpy(Xn).to(dev), torch.from_numpy(yn).to(dev)))
    TrainBatch(chunks, dev)   # untimed: first use of the graph kernels (module load, allocator warm-up)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    batch = TrainBatch(chunks, dev)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    step = GraphedTrainStep(model, batch, lr=1e-4, weight_decay=5e-4)

    def timed(fn):
        fn(); fn(); fn()
        torch.cuda.synchronize()
        times = []
        for _ in range(steps):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        return float(np.median(times)), float(np.mean(times))

    dt_eager, _ = timed(step.eager)
    step.capture()
    dt, dt_mean = timed(step.replay)
    lv = float(step.loss.item())
    return {'chunks_per_batch': B, 'message_passing_steps': len(batch.steps), 'edge_rows_per_batch': int(batch.edge_rows),
            'value': batch.edge_rows / dt, 'unit': 'edge-updates/s (forward+backward+optimizer)', 'chunks_per_s': B / dt,
            'ms_per_batch': 1e3 * dt, 'ms_per_batch_mean': 1e3 * dt_mean, 'ms_per_batch_eager_launches': 1e3 * dt_eager,
            'graph_build_ms_per_chunk': 1e3 * t_build / B, 'loss': lv,
            'value_incl_graph_build': batch.edge_rows / (dt + t_build), 'chunks_per_s_incl_graph_build': B / (dt + t_build),
            'graph_builder': batch.builder, 'api': 'GraphedTrainStep.replay() (CUDA graphs: zero + forward + losses + backward | Adam)'}


def run_train_leg(a, dev):
    import torch
    from trackmpnn_b200 import synth
    from trackmpnn_b200.models.track_mpnn import TrackMPNN
    torch.manual_seed(5)
    model = TrackMPNN('2d', synth.num_categories('kitti'), 64, 0, 'diff').to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=5e-4)  # train.py:329

    def chunk(seed):
        ts = synth.train_chunk_timestamps(seed, 5, 2)
        Xn, yn = synth.make_sequence(seed, None, a.train_dets, 'kitti', timestamps=ts)
        return Xn, yn

    data = [chunk(1000 + i) for i in range(a.train_chunks + 2)]
    dd = [(torch.from_numpy(Xn).to(dev), torch.from_numpy(yn).to(dev)) for Xn, yn in data]
    for X, y in dd[:2]:
        train_chunk_cuda(model, opt, X, y)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    edges = 0
    for X, y in dd[2:]:
        e, _ = train_chunk_cuda(model, opt, X, y)
        edges += e
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out = {'workload': f'C2 kitti-shaped training chunks (5 frames + 2 skip frames, ~Poisson({a.train_dets}) dets/frame, F=8, '
                       f'tp_classifier, teacher forcing, BPTT over all steps, CE+BCE, Adam), drop-in modules, 1 chunk at a time',
           'value': edges / dt, 'unit': 'edge-updates/s (forward+backward+optimizer)', 'chunks_per_s': a.train_chunks / dt,
           'ms_per_chunk': 1e3 * dt / a.train_chunks, 'edge_rows_per_chunk': edges // max(1, a.train_chunks)}
    out['batched'] = run_train_batched(a, dev, model, chunk)
    if not a.skip_cpu:
        from oracle import train_ref as T, trackmpnn_oracle as O
        params = O.init_params('2d', 3, 64, 'diff', seed=5)
        Xn, yn = data[2]
        t0 = time.perf_counter()
        r = T.train_chunk(params, Xn, yn)
        dtc = time.perf_counter() - t0
        ec = sum(int((g.ts < 0).sum()) for g in r['graphs'])
        out['cpu_baseline'] = {'value': ec / dtc, 'unit': 'edge-updates/s', 'cores': os.cpu_count(), 'kind': 'port',
                               'seconds': dtc, 'sample': '1 chunk of the workload (oracle/train_ref.py: torch fp32 ops on '
                                                         'edge lists + autograd, no optimizer step)'}
    return out


def run_c1_leg(a, dev):
    """BASELINE.json configs[0]: one KITTI-shaped sequence (--category=Car shape: ~10 detections / frame, 100 frames,
    window 5, F = 8), the reference's own CPU-runnable case.  Timed three ways: the reference's driver loop
    (infer.py:48-87) over the drop-in modules (one sequence, tensors in / tensors out, host-visible counts every
    call), the batched TrackEngine with S = 1 (CUDA-graph replay), and the oracle port on the host cores."""
    import torch
    from trackmpnn_b200 import synth
    from trackmpnn_b200.engine import TrackEngine
    from trackmpnn_b200.models.track_mpnn import TrackMPNN
    from trackmpnn_b200.utils.graph import initialize_graph, update_graph, decode_tracks
    torch.manual_seed(5)
    model = TrackMPNN('2d', synth.num_categories('kitti'), 64, 0, 'diff').to(dev).eval()
    Xn, yn = synth.make_sequence(5, 100, 10, 'kitti')
    X, y = torch.from_numpy(Xn).to(dev), torch.from_numpy(yn).to(dev)

    def two(sc):
        return torch.cat((1 - sc, sc), dim=1)

    def dropin():
        y_out = yn[0].astype(np.int64); y_out[:, 1] = -1
        frames = edges = 0
        with torch.no_grad():
            y_pred, feats, node_adj, edge_adj, labels, t_st, t_end = initialize_graph(X, y, 0, 'test', True)
            scores, logits, states, _ = model(feats, None, node_adj, edge_adj)
            scores = two(scores)
            for t in range(t_st, t_end):
                y_pred, feats, node_adj, edge_adj, labels = update_graph(node_adj, labels, scores, y_pred, X, y, t,
                                                                         use_hungraian=False, mode='test', cuda=True)
                scores, logits, states, _ = model(feats, states, node_adj, edge_adj)
                scores = two(scores)
                edges += int((y_pred[:, 0] == -1).sum())
                t_upto = t_end if t == t_end - 1 else t - 5 + 2
                y_pred, y_out, states, node_adj, labels, scores = decode_tracks(
                    states, node_adj, labels, scores, y_pred, y_out, t_upto, 0, use_hungraian=False, cuda=True)
                frames += 1
        return frames, edges

    dropin()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    frames, edges = dropin()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out = {'workload': 'C1 kitti-shaped, 1 sequence, 100 frames, ~Poisson(10) dets/frame, win 5, greedy decode, stock init',
           'dropin': {'frames_per_s': frames / dt, 'edge_updates_per_s': edges / dt, 'api': 'initialize_graph / update_graph / '
                      'TrackMPNN.forward / decode_tracks, one call each per frame (infer.py:48-87)'}}
    eng = TrackEngine(model, [(Xn[0], yn[0])], cur_win_size=5, ret_win_size=0)
    for _ in range(3):
        eng.run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        eng.run()
    _, st = eng.results()
    dt = (time.perf_counter() - t0) / reps
    out['engine_s1'] = {'frames_per_s': st['frames'] / dt, 'edge_updates_per_s': st['edge_updates'] / dt,
                        'api': 'TrackEngine(S = 1).run(), CUDA-graph replay (launch-latency bound: ~1 k rows per window)'}
    if not a.skip_cpu:
        from oracle import trackmpnn_oracle as O
        from oracle.infer_loop import run_infer
        params = O.init_params('2d', synth.num_categories('kitti'), 64, 'diff', seed=5)
        t0 = time.perf_counter()
        _, stc = run_infer(params, Xn, yn, ncategories=synth.num_categories('kitti'), cur_win_size=5)
        dtc = time.perf_counter() - t0
        out['cpu_baseline'] = {'frames_per_s': stc['frames'] / dtc, 'edge_updates_per_s': stc['edge_updates'] / dtc,
                               'cores': os.cpu_count(), 'kind': 'port',
                               'note': 'the reference itself measured 14.1 frames/s on this shape (SURVEY.md section 6)'}
    return out


def run_c4_leg(a, dev):
    """BASELINE.json configs[3]: stress graph, --cur-win-size 20, ~200 detections / frame (10^5 - 10^7 association rows per
    window), single-GPU roofline study of the aggregation and the GRU step.  Same engine, same kernels; the reference cannot
    run this size (dense N x N > host RAM)."""
    import copy
    import torch
    from trackmpnn_b200 import synth
    from trackmpnn_b200.engine import TrackEngine
    from trackmpnn_b200.models.track_mpnn import TrackMPNN
    b = copy.copy(a)
    b.win, b.dets, b.frames, b.seqs_per_gpu = a.c4_win, a.c4_dets, a.c4_frames, a.c4_seqs
    torch.manual_seed(5)
    model = TrackMPNN('2d', synth.num_categories(b.dataset), 64, 0, 'diff').to(dev).eval()
    seqs = make_sequences(b, 7000, b.seqs_per_gpu)
    eng = TrackEngine(model, seqs, cur_win_size=b.win, ret_win_size=0, use_cuda_graph=not a.no_cuda_graph,
                      block_aggregation=False if a.list_aggregation else 'auto')
    for _ in range(2):
        eng.run()
    eng.results()
    prepare_passes(eng)
    r = time_passes(eng, 2)
    roof, agg, comp, phases = rooflines(b, eng, r, r['ms'])
    sec = r['ms'] * 1e-3
    out = {'workload': f'C4 stress graph: {b.seqs_per_gpu} sequences x {b.frames} frames, ~Poisson({b.dets}) dets/frame, win {b.win}, '
                       'greedy decode, stock init', 'value': r['edges'] / sec, 'unit': 'edge-updates/s', 'frames_per_s': r['frames'] / sec,
           'edge_rows_per_window_mean': int(roof['edge_rows_per_launch'] / b.seqs_per_gpu), 'cap_rows_per_sequence': eng.cap_rows,
           'max_detection_rows_per_window': eng.max_dets, 'roofline': roof, 'roofline_aggregation': agg, 'roofline_compaction': comp,
           'phases': phases['share']}
    del eng
    torch.cuda.empty_cache()
    return out


def run_train_ddp_leg(a, dev, rank, world, barrier, reduce_):
    """BASELINE.json configs[4]: data-parallel training.  Every rank runs forward + losses + backward of its own
    batch of chunks (batched trainer); the backward kernels accumulate into views of ONE flat gradient buffer
    (trackmpnn_b200.parallel.FlatGradients), which the ranks all-reduce over NCCL as it is -- no pack / unpack -- then every
    rank takes the same Adam step.  BatchNorm statistics stay per chunk (no SyncBN), as in the reference.
    Gradient check: the all-reduced gradient of a small per-rank batch equals the gradient rank 0
    computes alone for the union of all ranks' chunks."""
    import torch
    import torch.distributed as dist
    from trackmpnn_b200 import parallel, synth
    from trackmpnn_b200.models.track_mpnn import TrackMPNN
    from trackmpnn_b200.train_engine import GraphedTrainStep, TrainBatch, batch_loss
    torch.manual_seed(5)
    model = TrackMPNN('2d', synth.num_categories('kitti'), 64, 0, 'diff').to(dev).train()
    flat = parallel.FlatGradients(model)
    params = list(model.parameters())
    B, n_steps = a.train_batch, a.train_chunks

    def chunk(seed):
        ts = synth.train_chunk_timestamps(seed, 5, 2)
        Xn, yn = synth.make_sequence(seed, None, a.train_dets, 'kitti', timestamps=ts)
        return torch.from_numpy(Xn).to(dev), torch.from_numpy(yn).to(dev)

    # ---- gradient check: sum over ranks of the per-rank gradients == single-GPU gradient of the summed loss -------------
    Bc = 4
    flat.zero()
    batch_loss(model, TrainBatch([chunk(5000 + rank * Bc + i) for i in range(Bc)], dev)).backward()
    flat.allreduce(average=False)
    g_ddp = flat.flat.clone()
    flat.zero()
    batch_loss(model, TrainBatch([chunk(5000 + i) for i in range(Bc * world)], dev)).backward()   # every rank: the union
    g_one = flat.flat.clone()
    err = float((g_ddp - g_one).abs().max() / g_one.abs().max())
    err = reduce_(err, dist.ReduceOp.MAX)
    # BatchNorm running statistics moved during the check: every rank did the same two forward passes -> still identical
    # ---- timed steps ----------------------------------------------------------------------------------------------------
    batch = TrainBatch([chunk(2000 + rank * 1000 + i) for i in range(B)], dev)
    step = GraphedTrainStep(model, batch, lr=1e-4, weight_decay=5e-4, allreduce=True)
    step.capture()          # three real (all-reduced) steps on every rank, then [zero + forward + backward] | [Adam] as graphs
    flat, nfl = step.flat, int(step.flat.flat.numel())
    ar_ms = 0.0
    for i in range(n_steps + 2):
        if i == 2:
            barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
        step._g_fb.replay()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        flat.allreduce(average=True)
        a1.record()
        step._g_opt.replay()
        if i >= 2:
            torch.cuda.synchronize()
            ar_ms += a0.elapsed_time(a1)
    torch.cuda.synchronize(); barrier()
    dt = reduce_(time.perf_counter() - t0, dist.ReduceOp.MAX)
    tot_edges = reduce_(batch.edge_rows * n_steps, dist.ReduceOp.SUM)
    # every rank must hold the same parameters after the same averaged step
    chk = torch.stack([p.detach().double().sum() for p in params]).sum()
    lo, hi = reduce_(float(chk), dist.ReduceOp.MIN), reduce_(float(chk), dist.ReduceOp.MAX)
    return {'workload': f'C5 data-parallel training: {world} ranks x {n_steps} optimizer steps x {B} kitti-shaped chunks per rank '
                        f'(~Poisson({a.train_dets}) dets/frame), batched forward + CE/BCE + backward per rank, one NCCL all-reduce of '
                        'the flat gradient buffer the backward kernels wrote (no pack / unpack), Adam',
            'value': tot_edges / dt, 'unit': 'edge-updates/s (forward+backward+allreduce+optimizer, all ranks)',
            'chunks_per_s': world * n_steps * B / dt, 'allreduce_floats': int(nfl),
            'allreduce_ms_avg_incl_peer_wait': ar_ms / max(1, n_steps),
            'replicas_in_sync': bool(abs(hi - lo) <= 1e-9 * max(1.0, abs(hi))),
            'ddp_gradient_vs_single_gpu': {'max_abs_err_over_max_abs_grad': err, 'ok': bool(err <= 1e-4),
                                           'what': f'all-reduced sum of {world} x {Bc}-chunk gradients vs one GPU over the {Bc * world} chunks '
                                                   '(another grouping of the per-rank sums: equal to round-off, not bit-identical; each run by itself is bit-reproducible)'}}


def prepare_passes(eng):
    """One more untimed eager pass with the per-launch event lists attached (nvidia-smi starts up meanwhile)."""
    import torch
    eng.use_cuda_graph = False
    eng.profile, eng.profile_compact, eng.profile_phases = [], [], []
    eng.run(); torch.cuda.synchronize()
    eng.profile, eng.profile_compact, eng.profile_phases = [], [], []


def time_passes(eng, steps, clocks=None):
    """K passes of `eng` bracketed by CUDA events on the current stream -- eager launches after prepare_passes() (the edge
    kernel, the aggregation and the window slide are then additionally bracketed launch by launch: eng.profile*), CUDA-graph
    replay otherwise.  Returns this rank's numbers.  No collective in here: the caller puts a barrier + synchronize on both
    sides and reduces the numbers over the ranks."""
    import torch
    from trackmpnn_b200 import _lib as L
    if clocks is not None:
        clocks.mark()
    l0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    edges = frames = dets = 0
    for _ in range(steps):
        eng.run()
        edges += eng.edge_updates.clone(); frames += eng.frames_done.clone(); dets += eng.det_updates.clone()
    e1.record()
    torch.cuda.synchronize()
    out = dict(ms=e0.elapsed_time(e1), edges=int(edges.item()), frames=int(frames.item()), dets=int(dets.item()),
               launches=L.launch_count() - l0, prof=eng.profile, cprof=eng.profile_compact, pprof=eng.profile_phases)
    eng.profile = eng.profile_compact = eng.profile_phases = None
    eng.ga.check_status()
    return out


def rooflines(a, eng, r, ms):
    """The three HBM rooflines of the pass `r` (time_passes): the fused edge step, the detection aggregation, the window slide.
    `frac` is ALGORITHMIC bytes / time / peak; `frac_dram` is the DRAM traffic an ncu capture of the same
    kernel measured (profiles/r02_dram_traffic.json, bytes per association row) / time / peak."""
    prof, cprof, pprof = r['prof'], r['cprof'], r['pprof']
    k_ms = sum(p[0].elapsed_time(p[1]) for p in prof)
    k_edges = sum(int(p[2].item()) for p in prof)
    fresh = sum(int(p[5].item()) for p in prof)   # association rows appended this frame: state exactly 0, alias the slab's zero row
    n_l = max(1, len(prof))
    hbm_peak, peak_src = measured_peaks()
    achieved = BYTES_PER_EDGE_UPDATE * k_edges / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    tk = {'auto': 'pre'}.get(eng.tensor_kernel, eng.tensor_kernel)
    kname = ({'pre': 'k_mp_edge_tc3 (fused edge step: endpoints prepared once per detection, far-endpoint images fetched by TMA tile::gather4, '
                     '24 tcgen05.mma kind::f16 per tile (3-term fp16 split, N = 192), TMEM accumulators, elect.sync MMA issuer warp, two '
                     'epilogue teams, own-row images triple-buffered and reused in place as the transpose buffer)',
              'gather': 'k_mp_edge_tc (fused gather-diff + GRU + head; tcgen05.mma kind::f16, 3-term fp16 split, TMEM accumulators)'}[tk]
             if eng.tensor else 'k_mp_edge<64> (fused gather-diff + GRU + head, fp32 FMA path)')
    tr_e, src_e = measured_traffic('k_mp_edge_tc3') if eng.tensor else (None, None)
    blocks = getattr(eng, '_agg_blocks', None) is not None
    if blocks:   # the block pass and the per-detection combine are timed (and counted) together
        tr_a, src_a = measured_traffic('k_aggregate_blocks')
        tr_c, _ = measured_traffic('k_aggregate_blocks_combine')
        tr_a = tr_a + tr_c if tr_a is not None and tr_c is not None else None
    else:
        tr_a, src_a = measured_traffic('k_aggregate_dets')
    traffic = tr_e * k_edges / n_l if tr_e else None
    roof = {'kernel': kname, 'bound': 'hbm', 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': achieved / hbm_peak,
            'frac_algorithmic': achieved / hbm_peak,
            'frac_dram': (traffic / (k_ms / n_l * 1e-3) / 1e9 / hbm_peak) if traffic and k_ms > 0 else None,
            'traffic': traffic, 'traffic_source': src_e,
            'algorithmic_bytes_per_launch': BYTES_PER_EDGE_UPDATE * k_edges / n_l,
            'peak_source': peak_src, 'launches_timed': len(prof), 'avg_launch_ms': k_ms / n_l,
            'share_of_step': k_ms / ms if ms > 0 else None,
            'fp32_tflops': FLOP_PER_ROW_UPDATE * k_edges / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0,
            'algorithmic_bytes_per_edge_update': BYTES_PER_EDGE_UPDATE, 'edge_rows_per_launch': k_edges / n_l,
            'fresh_rows_share': fresh / max(1, k_edges)}
    # detection aggregation (K1): algorithmic bytes = one read of every association row THAT HOLDS STATE (rows appended this
    # frame are exactly zero and alias one zero row per slab: nothing to read) + one 256 B sum written per detection
    a_ms = sum(p[3].elapsed_time(p[4]) for p in prof)
    agg_bytes = 256.0 * (k_edges - fresh) + 256.0 * r['dets']
    tr_agg = tr_a * k_edges / n_l if tr_a else None
    agg = {'kernel': ('k_aggregate_blocks + k_aggregate_blocks_combine (one pass over every dense edge block: row sums of 32-source '
                      'stripes + per-stripe column partials, every association row read once; blocks appended this frame skipped)'
                      if blocks else 'k_aggregate_dets (CSR segmented signed sum of incident association rows, CTA per detection)'),
           'bound': 'hbm', 'achieved': agg_bytes / (a_ms * 1e-3) / 1e9 if a_ms > 0 else 0.0, 'peak': hbm_peak, 'unit': 'GB/s',
           'frac': agg_bytes / (a_ms * 1e-3) / 1e9 / hbm_peak if a_ms > 0 else 0.0,
           'frac_dram': (tr_agg / (a_ms / n_l * 1e-3) / 1e9 / hbm_peak) if tr_agg and a_ms > 0 else None,
           'avg_launch_ms': a_ms / n_l, 'share_of_step': a_ms / ms if ms > 0 else None,
           'algorithmic_bytes': '256 B per association row that holds state (rows appended this frame are zero and not read) + 256 B per detection',
           'traffic': tr_agg, 'traffic_source': src_a}
    # window slide (K4): keep-mask scan + order-preserving compaction with index remap.  With deferred compaction
    # the states stay where they are: per surviving row 36 B of metadata are read and written, 12 B of position maps
    # written, 4 B of new_of_old written and read back for the two endpoints; per row 1 B keep flag + 4 B new_of_old.
    c_ms = sum(p[0].elapsed_time(p[1]) for p in cprof)
    rows_in = sum(int(p[2].item()) for p in cprof)
    rows_out = sum(int(p[3].item()) for p in cprof)
    per_kept = (36 + 36 + 12 + 8) if eng.deferred else (36 + 36 + 8 + 512 * eng.G)
    c_bytes = float(per_kept) * rows_out + 9.0 * rows_in
    comp = {'kernel': 'tmpnn_graph_compact (k_compact_count / scan / map / move: prefix-sum stream compaction + src/dst remap'
                      + (', position maps instead of moving the states)' if eng.deferred else ', states moved)'),
            'bound': 'hbm', 'achieved': c_bytes / (c_ms * 1e-3) / 1e9 if c_ms > 0 else 0.0, 'peak': hbm_peak, 'unit': 'GB/s',
            'frac': c_bytes / (c_ms * 1e-3) / 1e9 / hbm_peak if c_ms > 0 else 0.0, 'avg_launch_ms': c_ms / max(1, len(cprof)),
            'share_of_step': c_ms / ms if ms > 0 else None, 'rows_in': rows_in, 'rows_kept': rows_out,
            'algorithmic_bytes': f'{per_kept} B per surviving row + 9 B per row', 'traffic': None}
    # the reference's three phases per frame (update_graph 20 % / forward 42 % / decode_tracks 33 % of its loop)
    ph_ms = [sum(p[i].elapsed_time(p[i + 1]) for p in pprof) for i in range(3)]
    ph_tot = sum(ph_ms) or 1.0
    phases = {'ticks_timed': len(pprof), 'update_ms_per_tick': ph_ms[0] / max(1, len(pprof)),
              'forward_ms_per_tick': ph_ms[1] / max(1, len(pprof)), 'decode_ms_per_tick': ph_ms[2] / max(1, len(pprof)),
              'share': {'update': ph_ms[0] / ph_tot, 'forward': ph_ms[1] / ph_tot, 'decode': ph_ms[2] / ph_tot},
              'reference_share': {'update': 0.20, 'forward': 0.42, 'decode': 0.33},
              'what': 'update = tmpnn_graph_append (+ Hungarian re-association); forward = input transform + incidence index + '
                      'aggregation + association-row and detection-row steps; decode = association + track walk + window slide'}
    return roof, agg, comp, phases


def sharding_check(a, dev, rank, world):
    """Identical per-sequence tracks whatever GPU and batch a sequence lands in: every rank tracks
    the SAME two check sequences inside a batch of its own (rank-specific) other sequences, with decision-exercising weights
    (x20, edge-head bias 0) so that associations, chain walks and deletions all carry load; rank 0 additionally tracks the two
    alone.  Returns this rank's track arrays of the two sequences (int32 tensor on `dev`) and rank 0's stand-alone ones."""
    import torch
    from trackmpnn_b200 import synth
    from trackmpnn_b200.engine import TrackEngine
    from trackmpnn_b200.models.track_mpnn import TrackMPNN
    torch.manual_seed(5)
    model = TrackMPNN('2d', synth.num_categories(a.dataset), 64, 0, 'diff')
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() >= 2:
                p.mul_(20.0)
        model.output_transform_edge.bias.fill_(0.0)
    model = model.to(dev).eval()
    frames = min(a.frames, 30)

    def seq(seed):
        X, y = synth.make_sequence(seed, frames, a.dets, a.dataset)
        return X[0], y[0]

    check = [seq(90001), seq(90002)]
    mine = [seq(91000 + 10 * rank + i) for i in range(1 + rank % 3)]
    eng = TrackEngine(model, mine[:1] + check + mine[1:], cur_win_size=a.win, ret_win_size=0)
    outs, _ = eng.run().results()
    got = np.concatenate(outs[1:3]).astype(np.int32)
    alone = None
    if rank == 0:
        outs0, _ = TrackEngine(model, check, cur_win_size=a.win, ret_win_size=0).run().results()
        alone = np.concatenate(outs0).astype(np.int32)
    del eng
    return torch.from_numpy(got).to(dev), alone


def main():
    a = parse()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if a.impl == 'reference':
        run_reference(a, rank)
        return

    import torch
    import torch.distributed as dist
    from trackmpnn_b200 import parallel, synth
    from trackmpnn_b200.engine import TrackEngine
    from trackmpnn_b200.models.track_mpnn import TrackMPNN

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    # Every collective of this program is issued from main()'s top level by ALL ranks (tests/test_bench_collectives.py
    # checks that none sits under a rank test): a reduction only rank 0 joins hangs until the NCCL watchdog aborts.
    def barrier():
        if world > 1:
            dist.barrier()

    def reduce_(x, op):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=op)
            return float(t.item())
        return float(x)

    MAX, SUM, MIN = dist.ReduceOp.MAX, dist.ReduceOp.SUM, dist.ReduceOp.MIN
    torch.manual_seed(5)
    model = TrackMPNN('2d', synth.num_categories(a.dataset), 64, 0, 'diff').to(dev).eval()

    # ---- weak-scaling leg (the headline line): --seqs-per-gpu sequences on EVERY rank ------------------------------------
    seqs = make_sequences(a, 5 + rank * a.seqs_per_gpu, a.seqs_per_gpu)
    eng = TrackEngine(model, seqs, cur_win_size=a.win, ret_win_size=0, use_cuda_graph=not a.no_cuda_graph,
                      deferred_compaction=not a.no_deferred, block_aggregation=False if a.list_aggregation else 'auto', det_tensor=not a.fma_dets)
    for _ in range(max(a.warmup, 1)):
        eng.run()
    eng.results()
    clocks = ClockSampler(local)
    clocks.start()
    prepare_passes(eng)
    barrier(); torch.cuda.synchronize()
    r = time_passes(eng, a.steps, clocks)
    barrier()
    clk = clocks.stop()
    ms = reduce_(r['ms'], MAX)
    tot_edges = reduce_(r['edges'], SUM)
    tot_frames = reduce_(r['frames'], SUM)
    tot_dets = reduce_(r['dets'], SUM)
    roof, agg, comp, phases = rooflines(a, eng, r, r['ms'])
    cap_rows, deferred = eng.cap_rows, eng.deferred

    # ---- end to end: host buffers in, decoded tracks out, copies inside the timed region ---------------------------------
    e2e_dt = e2e_edges = h2d = d2h = 0
    if not a.skip_e2e:
        eng.use_cuda_graph = not a.no_cuda_graph
        x_host = eng.frames.x.cpu().pin_memory()
        tab_host = [t.cpu().pin_memory() for t in (eng.frames.frame_ptr, eng.frames.frame_dets, eng.frames.det_ptr)]
        out_host = torch.empty(eng.y_out_track.shape, dtype=eng.y_out_track.dtype).pin_memory()
        h2d = x_host.numel() * 4 + sum(t.numel() * 4 for t in tab_host)
        d2h = out_host.numel() * 4
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            eng.frames.x.copy_(x_host, non_blocking=True)
            for dst, src in zip((eng.frames.frame_ptr, eng.frames.frame_dets, eng.frames.det_ptr), tab_host):
                dst.copy_(src, non_blocking=True)
            eng.run()
            out_host.copy_(eng.y_out_track, non_blocking=True)
            torch.cuda.synchronize()
            e2e_edges += int(eng.edge_updates.item())
        e2e_dt = time.perf_counter() - t0
        barrier()
    e2e_dt = reduce_(e2e_dt, MAX)
    e2e_tot = reduce_(e2e_edges, SUM)
    e2e = None
    if not a.skip_e2e:
        e2e = {'value': e2e_tot / e2e_dt, 'unit': 'edge-updates/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
               'ms_per_step': 1e3 * e2e_dt / a.steps, 'api': 'TrackEngine.run() + results copy, CUDA-graph replay'}
    del eng
    torch.cuda.empty_cache()

    # ---- strong-scaling leg: BASELINE.json configs[2] literally -- --seqs-per-gpu sequences IN TOTAL, bin-packed onto the
    # ranks by cost (parallel.partition_sequences over parallel.sequence_cost); at N = 1 it is the weak leg ----------------
    strong = None
    if world > 1 and not a.skip_strong:
        all_seqs = seqs if rank == 0 else make_sequences(a, 5, a.seqs_per_gpu)
        costs = [parallel.sequence_cost(np.bincount(np.asarray(y)[:, 0].astype(np.int64)), a.win) for _, y in all_seqs]
        bins = parallel.partition_sequences(costs, world)
        load = [sum(costs[i] for i in b) for b in bins]
        eng = TrackEngine(model, [all_seqs[i] for i in bins[rank]], cur_win_size=a.win, ret_win_size=0,
                          use_cuda_graph=not a.no_cuda_graph, deferred_compaction=not a.no_deferred)
        for _ in range(max(a.warmup, 1)):
            eng.run()
        eng.results()
        # 32 sequences per GPU make a frame ~1 ms of ~45 launches: timed as the engine is used (CUDA-graph replay); the
        # per-launch roofline of rank 0 comes from one more, eager, pass afterwards
        barrier(); torch.cuda.synchronize()
        rs = time_passes(eng, a.steps)
        barrier()
        prepare_passes(eng)
        rp = time_passes(eng, 1)
        s_roof = rooflines(a, eng, rp, rp['ms'])[0]
        del eng, all_seqs
        torch.cuda.empty_cache()
    else:
        rs, bins, load, s_roof = dict(ms=0.0, edges=0, frames=0), None, None, None
    s_ms = reduce_(rs['ms'], MAX)
    s_ms_min = reduce_(rs['ms'], MIN)
    s_edges = reduce_(rs['edges'], SUM)
    s_frames = reduce_(rs['frames'], SUM)
    if bins is not None:
        strong = {'workload': f'{a.seqs_per_gpu} sequences in total, bin-packed onto {world} ranks by sum_t D_t x (detections of the '
                              f'previous {a.win - 1} frames) (longest-processing-time first)',
                  'value': s_edges / (s_ms * 1e-3), 'unit': 'edge-updates/s', 'frames_per_s': s_frames / (s_ms * 1e-3),
                  'ms_per_step': s_ms / a.steps, 'ms_per_step_fastest_rank': s_ms_min / a.steps,
                  'sequences_per_rank': [len(b) for b in bins], 'cost_imbalance': max(load) / (sum(load) / world),
                  'rank0_edge_kernel_frac': s_roof['frac'], 'rank0_edge_kernel_share_of_eager_step': s_roof['share_of_step'],
                  'timed_passes': 'CUDA-graph replay', 'scaling': 'strong'}
    del seqs

    # ---- the same sequence gives the same tracks on every GPU, in every batch --------------------------------------------
    sharding_identical = None
    chk_info = None
    if not a.skip_check:
        got, alone = sharding_check(a, dev, rank, world)
        if world > 1:
            gathered = [torch.empty_like(got) for _ in range(world)]
            dist.all_gather(gathered, got)
        else:
            gathered = [got]
        if rank == 0:
            per_rank = [t.cpu().numpy() for t in gathered]
            sharding_identical = bool(all(np.array_equal(p, alone) for p in per_rank))
            chk_info = {'identical': sharding_identical, 'ranks_compared': world, 'detections_compared': int(alone.size),
                        'tracks': int(alone.max()) + 1, 'multi_detection_tracks': int((np.bincount(alone[alone >= 0]) > 1).sum()),
                        'what': 'two check sequences tracked on every rank inside a rank-specific batch (x20 weights, edge-head bias 0) '
                                'and alone on rank 0: track ids compared bit for bit'}

    # ---- side legs ------------------------------------------------------------------------------------------------------
    cpu = train = c1 = None
    if rank == 0 and world == 1 and not a.skip_cpu:
        frames_cap = a.cpu_frames or 60
        stc, dtc = cpu_oracle_sample(a, frames_cap)
        cpu = {'value': stc['edge_updates'] / dtc, 'unit': 'edge-updates/s', 'cores': os.cpu_count(), 'kind': 'port',
               'frames_per_s': stc['frames'] / dtc, 'seconds': dtc,
               'sample': f'1 sequence of the workload, first {frames_cap} frames (oracle port: numpy BLAS threads for '
                         f'the GEMMs, one thread for graph bookkeeping)'}
    def side(fn):   # a side leg that fails must not take the headline line with it: its error is reported in its place
        try:
            return fn(a, dev)
        except Exception as e:   # noqa: BLE001
            import traceback
            traceback.print_exc()
            torch.cuda.empty_cache()
            return {'error': f'{type(e).__name__}: {e}'}

    if rank == 0 and world == 1 and not a.skip_train:
        train = side(run_train_leg)
        c1 = side(run_c1_leg)
    c4 = None
    if rank == 0 and world == 1 and not a.skip_c4:
        c4 = side(run_c4_leg)
    train_ddp = None
    if world > 1 and not a.skip_train:
        train_ddp = run_train_ddp_leg(a, dev, rank, world, barrier, reduce_)

    if rank == 0:
        sec = ms * 1e-3
        out = {'metric': 'edge_updates_per_s', 'value': tot_edges / sec, 'unit': 'edge-updates/s',
               'frames_per_s': tot_frames / sec, 'n_gpus': world, 'steps': a.steps, 'warmup': max(a.warmup, 1),
               'ms_per_step': ms / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
               'dtype': 'f32', 'data': 'synthetic',
               'config': {'workload': workload_name(a), 'l2': 'inputs_larger_than_l2 (state >= 4 GB per GPU vs 126 MB L2)',
                          'edge_rows_per_step_per_gpu': r['edges'] // max(1, a.steps), 'det_rows_per_step_per_gpu': r['dets'] // max(1, a.steps),
                          'frames_per_step_per_gpu': r['frames'] // max(1, a.steps), 'cap_rows_per_sequence': cap_rows,
                          'deferred_compaction': deferred, 'timed_passes': 'eager launches (edge kernel bracketed by CUDA events)',
                          'cpu_baseline_kind': 'port: the reference is Python and is not on the bench box; the CPU arm times the '
                                               'numpy oracle port, which is ~100x faster than the reference own dense N x N code'},
               'det_updates_per_s': tot_dets / sec, 'phases': phases,
               'roofline': roof, 'roofline_aggregation': agg, 'roofline_compaction': comp, 'cpu_baseline': cpu, 'e2e': e2e,
               'gpu_launches': int(r['launches']), 'clocks': clk, 'strong': strong, 'sharding_identical': sharding_identical,
               'sharding_check': chk_info, 'c1': c1, 'train': train, 'c4': c4, 'train_ddp': train_ddp}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    # stdout carries exactly one line, the JSON result: anything a library writes to file descriptor 1 while the bench
    # runs (NCCL prints its version banner there when NCCL_DEBUG is set) is routed to stderr
    sys.stdout.flush()
    _json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_json_fd, 'w')
    main()
