"""Batched, device-resident tracking engine: many independent sequences advance in lock step.

This is the loop of the reference's ``infer.py:48-87`` (initialise -> forward, then per frame
update -> forward -> decode) for S sequences at once.  Every sequence owns one slab of the
structure-of-arrays graph (``include/tmpnn.h``); per frame the engine enqueues a fixed list
of kernels whose sizes come from device memory, so the host never reads the graph back and
the whole frame can be replayed as a CUDA graph.  Sequences shard across GPUs with no
communication (one engine per rank).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .device_graph import FrameTable, SlabGraph, SlabIndex
from . import functional as F_

H = L.HIDDEN
_I32 = torch.int32


def worst_case_rows(frame_counts, window):
    """Upper bound of the rows one sequence's window graph can hold: all detections of
    ``window`` consecutive timesteps active and pairwise connected."""
    c = np.asarray(frame_counts, dtype=np.int64)
    best = 0
    for t in range(len(c)):
        w = c[max(0, t - window + 1):t + 1]
        s = int(w.sum())
        pairs = (s * s - int((w * w).sum())) // 2
        best = max(best, s + pairs)
    return best


class TrackEngine:
    def __init__(self, model, sequences, cur_win_size=5, ret_win_size=0, device=None, cap_rows=None,
                 use_cuda_graph=True, tensor_cores='auto', use_hungarian=False, structured_index=True,
                 deferred_compaction=True, tensor_kernel='auto', tp_classifier=True, block_aggregation='auto', det_tensor=True):
        """sequences: list of (X [ND, F] float32, y [ND, 2] = [ts, track_id]) host arrays.

        deferred_compaction: the slide of the window does not move the hidden states; the next step reads them
        through the position maps the compaction emits (tmpnn_graph.phys / psrc / pdst) and writes its output
        densely, so the state only ever crosses HBM once per step in each direction.

        block_aggregation: with the structured index the detections' aggregation reads each dense edge block once
        (tmpnn_aggregate_dets_blocks) instead of walking the incidence lists; 'auto' = whenever the tensor-core path is on
        (large batches).  det_tensor: the detection rows run on the tensor cores as well (tmpnn_mp_det_fwd_tc)."""
        self.model = model
        self.tp_classifier = bool(tp_classifier)   # False: infer.py's --no-tp-classifier
        self.dev = device if device is not None else next(model.parameters()).device
        if self.dev.type != 'cuda':
            raise L.TmpnnError('TrackEngine needs a CUDA device; there is no CPU path')
        if any(g.msg_type not in ('diff', 'concat') for g in model.factor_grus):
            raise ValueError('bad msg_type')
        self.W, self.R = int(cur_win_size), int(ret_win_size)
        self.S = len(sequences)
        self.G = len(model.feature_idx)
        self.ldh = self.G * H
        xs = [np.asarray(x, dtype=np.float32).reshape(-1, np.asarray(x).shape[-1]) for x, _ in sequences]
        ys = [np.asarray(y).reshape(-1, 2) for _, y in sequences]
        self.frames = FrameTable(ys, self.dev, xs)
        fp = self.frames.host_frame_ptr
        counts = fp[:, 1:] - fp[:, :-1]
        self.t_hi = int(self.frames.t_max) + 1
        if cap_rows is None:
            cap_rows = max(worst_case_rows(counts[s], self.W + self.R) for s in range(self.S))
            max_frame = int(counts.max()) if counts.size else 1
            if deferred_compaction:
                # a frame's detections are staged behind the previous (uncompacted) graph, and the slab's last
                # row is the shared all-zero row the new association rows read
                cap_rows += max_frame + 1
            cap_rows = max(64, (cap_rows + 63) // 64 * 64)
        self.cap_rows = int(cap_rows)
        self.deferred = bool(deferred_compaction)
        max_frame = int(counts.max()) if counts.size else 1
        self.cap_new = max(1, 2 * self.S * max_frame)
        max_dets = int(max(worst_case_dets(counts[s], self.W + self.R) for s in range(self.S)))
        dev = self.dev
        self.ga = SlabGraph(self.S, self.cap_rows, dev, with_labels=False, deferred=self.deferred)
        self.gb = SlabGraph(self.S, self.cap_rows, dev, with_labels=False, status=self.ga.status, deferred=self.deferred)
        self._g0, self._g1 = self.ga, self.gb
        n_all = self.S * self.cap_rows
        self.h_cur = torch.zeros((n_all, self.ldh), dtype=torch.float32, device=dev)
        self.h_alt = torch.zeros((n_all, self.ldh), dtype=torch.float32, device=dev)
        self.index = SlabIndex(self.ga, cap_dets=self.S * max_dets, cap_inc=2 * n_all)
        z = lambda n: torch.zeros(n, dtype=_I32, device=dev)
        self.st = dict(phase=z(self.S), skip_until=z(self.S), t_end=z(self.S), active=z(self.S), t_upto=z(self.S),
                       fresh=z(self.S), last_new=z(self.S))
        self.st_c = L.SeqState(*[L.ptr(self.st[k]) for k in L.SEQ_STATE_FIELDS])
        self.y_out_track = torch.full((max(1, self.frames.total_dets),), -1, dtype=_I32, device=dev)
        self.next_track_id = z(self.S)
        self.t_dev = z(1)
        self.new_rows = z(self.cap_new)
        self.new_x = z(self.cap_new)
        self.n_new = z(2)
        self.n_appended = z(self.S)
        self.append_scratch = z(int(L.lib().tmpnn_graph_append_scratch_ints(self.S, self.cap_rows)))
        self.compact_scratch = z(int(L.lib().tmpnn_graph_compact_scratch_ints(self.S, self.cap_rows)))
        self.keep = torch.zeros(n_all, dtype=torch.uint8, device=dev)
        self.new_of_old = z(n_all)
        self.agg = torch.empty((self.index.cap_dets, H), dtype=torch.float32, device=dev)
        self._aggs = None   # per-group aggregates when the detection rows run on the tensor cores (kept for the FMA re-run)
        self.a_scratch = torch.empty((self.cap_new, H), dtype=torch.float32, device=dev)
        # work counters, accumulated on the device
        self.edge_updates = torch.zeros(1, dtype=torch.int64, device=dev)
        self.det_updates = torch.zeros(1, dtype=torch.int64, device=dev)
        self.frames_done = torch.zeros(1, dtype=torch.int64, device=dev)
        self.use_cuda_graph = use_cuda_graph
        # the engine's graphs are chains of dense edge blocks: the index is derived from the block boundaries
        self.structured_index = bool(structured_index)
        # --hungarian (infer.py:143): optimal assignment per timestep instead of the greedy arg-max
        self.use_hungarian = bool(use_hungarian)
        self.max_dets = max_dets
        self.decode_scratch = z(int(L.lib().tmpnn_graph_decode_scratch_ints(self.S, max_dets)))
        self.hung_scratch = None
        if self.use_hungarian:
            nbytes = int(L.lib().tmpnn_hungarian_scratch_bytes(self.S, max_dets))
            self.hung_scratch = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=dev)
        diff = all(g.msg_type == 'diff' for g in model.factor_grus)
        # tcgen05 path when the batch can fill 128-row tiles on every SM; fp32 FMA path otherwise.
        # tensor_kernel 'pre' (= 'auto'): endpoints prepared once per detection row (both msg_types); 'gather':
        # gathered and split per association row ('diff' only)
        self.tensor_kernel = tensor_kernel
        self.tensor = (diff or tensor_kernel != 'gather') and (
            self.S * self.cap_rows >= F_.TENSOR_MIN_ROWS if tensor_cores == 'auto' else bool(tensor_cores))
        self._tc_scratch = {}
        self._gat_scratch = {}
        # detection rows on the tensor cores too: the edge-step kernel in detection mode (needs the prepared-endpoint kernel's
        # scratch and the structured index's segment tables)
        self.det_tensor = bool(det_tensor) and self.tensor and self.structured_index and tensor_kernel in ('auto', 'pre')
        if self.det_tensor:
            self._aggs = [self.agg] + [torch.empty_like(self.agg) for _ in range(self.G - 1)]
        # block-structured aggregation (needs the structured index): per slab one run sum per (source, edge block) and one
        # column partial per (stripe of 32 sources, detection)
        self._agg_blocks = None
        if block_aggregation == 'auto':   # three launches instead of one: only pays where the graphs are large
            block_aggregation = self.tensor
        if self.structured_index and block_aggregation:
            cap_runs = max_dets * min(self.W + self.R, 64)
            cap_cpart = self.cap_rows // 32 + max_dets + 64
            nbytes = int(L.lib().tmpnn_aggregate_blocks_scratch_bytes(self.S, cap_runs, cap_cpart))
            self._agg_blocks = dict(cap_runs=cap_runs, cap_cpart=cap_cpart,
                                    scratch=torch.empty((nbytes + 15) // 16 * 4, dtype=torch.float32, device=dev))
        self.profile = None  # when a list: (edge start, edge end, n_edges tensor, aggregation start, aggregation end, new edge rows) per step
        self.profile_compact = None  # when a list: (start, end, rows before, rows after) per window slide
        # when a list: four events per tick bracketing the reference's three phases -- update_graph (append) |
        # forward (input transform, index, aggregation, both row types) | decode_tracks (associate, walk, window slide)
        self.profile_phases = None
        # when a list (eager ticks only): every step also aggregates through the incidence lists (tmpnn_aggregate_dets) and
        # appends (detections, max |difference| to the block-structured result, max |reference|)
        self.check_aggregation = None
        self._graph = None
        self._graph_key = None
        self.ticks = 0

    # ---- one engine tick --------------------------------------------------------------------
    def _forward(self, g, h_in, h_out):
        model = self.model
        st = L.stream()
        # K0 on the new detection rows (device-side count), all feature groups
        for grp in range(self.G):
            seq = model.input_transforms[grp]
            lin1, bn, lin2 = seq[0], seq[1], seq[3]
            cols = model.feature_idx[grp]
            if bn.training:
                raise L.TmpnnError('TrackEngine is an inference engine: call model.eval() first')
            L.call('tmpnn_input_linear1', L.ptr(self.frames.x), self.frames.ldx, int(cols[0]), len(cols),
                   L.ptr(self.new_x), L.ptr(lin1.weight.detach()), L.ptr(lin1.bias.detach()), L.ptr(self.a_scratch),
                   L.ptr(self.n_new), 0, st)
            L.call('tmpnn_input_bn_relu_linear2', L.ptr(self.a_scratch), L.ptr(bn.running_mean), L.ptr(bn.running_var),
                   L.ptr(bn.weight.detach()), L.ptr(bn.bias.detach()), L.ptr(lin2.weight.detach()),
                   L.ptr(lin2.bias.detach()), L.ptr(h_in), self.ldh, grp * H, L.ptr(self.new_rows),
                   L.ptr(self.n_new), 0, st)
        self.index.build(g, self.st['active'], structured=self.structured_index)
        packs = F_.packed_cells(model)
        tc = F_.packed_cells_tc(model) if self.tensor else None
        tc_node = F_.packed_cells_tc(model, node=True) if self.det_tensor else None
        for grp in range(self.G):
            agg = self._aggs[grp] if self.det_tensor else self.agg
            concat = int(model.factor_grus[grp].msg_type == 'concat')
            if self.profile is not None:
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
            F_.aggregate_for_dets(model.factor_grus[grp], g, self.index, h_in, self.ldh, grp * H, agg, self._gat_scratch,
                                  blocks=self._agg_blocks)
            if self.check_aggregation is not None and self._agg_blocks is not None and model.factor_grus[grp].gat is None:
                ref = torch.empty_like(agg)
                L.call('tmpnn_aggregate_dets', g.c, self.index.c, L.ptr(h_in), self.ldh, grp * H, L.ptr(ref), st)
                nd = int(self.index.n_dets.item())
                self.check_aggregation.append((nd, float((ref[:nd] - agg[:nd]).abs().max()) if nd else 0.0,
                                               float(ref[:nd].abs().max()) if nd else 0.0))
            if self.profile is not None:
                a1.record()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            if self.tensor:
                F_.edge_step_tc(model, g, self.index, h_in, h_out, self.ldh, grp, self.G, tc[grp], self._tc_scratch,
                                kernel=self.tensor_kernel)
            else:
                L.call('tmpnn_mp_edge_fwd', g.c, self.index.c, L.ptr(h_in), L.ptr(h_out), self.ldh, grp, self.G, concat,
                       L.ptr(packs[grp][0]), st)
            if self.profile is not None:
                e1.record()
                self.profile.append((e0, e1, self.index.n_edges.clone(), a0, a1, self.n_new[1:2].clone()))
            if self.det_tensor:
                F_.det_step_tc(g, self.index, h_in, h_out, self.ldh, grp, self.G, tc_node[grp], agg, self._tc_scratch)
            else:
                L.call('tmpnn_mp_det_fwd', g.c, self.index.c, L.ptr(h_in), L.ptr(h_out), self.ldh, grp, self.G,
                       L.ptr(packs[grp][1]), L.ptr(agg), st)
        if self.tensor:
            F_.rerun_edges_if_out_of_range(model, g, self.index, h_in, h_out, self.ldh, packs, aggs=self._aggs)
        if not self.tp_classifier:
            # --no-tp-classifier (infer.py:54-57, 77-80): detections count as true positives in association and decoding
            L.call('tmpnn_graph_force_det_scores', g.c, self.index.c, st)
        L.call('tmpnn_graph_counters', L.ptr(self.edge_updates), L.ptr(self.index.n_edges), L.ptr(self.det_updates),
               L.ptr(self.index.n_dets), None, None, 0, None, st)

    def _start(self):
        g = self.ga
        # the initial step runs h_alt -> h_cur so that h_cur holds the current state afterwards
        # (buffer roles never change: CUDA graphs bake the pointers in)
        L.call('tmpnn_graph_append', g.c, self.frames.c, C.byref(self.st_c), None, 1, self.W, 0, L.ptr(self.h_alt),
               self.ldh, L.ptr(self.new_rows), L.ptr(self.new_x), L.ptr(self.n_new), self.cap_new,
               L.ptr(self.n_appended), L.ptr(self.append_scratch), L.stream())
        self._forward(g, self.h_alt, self.h_cur)
        if self.deferred:  # the state now sits at the logical rows of h_cur
            L.call('tmpnn_graph_phys_identity', g.c, L.stream())
        # update_graph re-associates from the previous scores before it appends (utils/graph.py:251-268).
        # Inside the loop that result is carried over from decode_tracks (same scores, and deletion
        # cannot change a survivor's association); after the initial forward it is computed here.
        self._associate(g)

    def _associate(self, g):
        if self.use_hungarian:
            L.call('tmpnn_graph_associate_hungarian', g.c, self.index.c, None, L.ptr(self.st['active']), self.max_dets, 0, 0,
                   0.5, L.ptr(self.hung_scratch), L.stream())
        else:
            L.call('tmpnn_graph_associate', g.c, self.index.c, 0, L.ptr(self.st['active']), L.stream())

    def _tick(self, flip=False):
        """One iteration of infer.py:60-87 for every sequence at t = *t_dev, then t += 1.  With deferred
        compaction the state alternates between the two buffers (flip = odd tick: h_alt -> h_cur)."""
        g, go = self.ga, self.gb
        h_in, h_out = (self.h_alt, self.h_cur) if (flip and self.deferred) else (self.h_cur, self.h_alt)
        st = L.stream()
        ph = None
        if self.profile_phases is not None:
            ph = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ph[0].record()
        if self.use_hungarian:
            # update_graph re-solves the assignment on the graph decode_tracks left behind (utils/graph.py:247-249);
            # unlike the greedy choice it is not invariant under the deletion, so it is recomputed here
            self.index.build(g, self.st['active'], structured=self.structured_index)
            self._associate(g)
        with L.nvtx_range('phase/update_graph'):
            L.call('tmpnn_graph_append', g.c, self.frames.c, C.byref(self.st_c), L.ptr(self.t_dev), 0, self.W, 0,
                   L.ptr(h_in), self.ldh, L.ptr(self.new_rows), L.ptr(self.new_x), L.ptr(self.n_new), self.cap_new,
                   L.ptr(self.n_appended), L.ptr(self.append_scratch), st)
        if ph:
            ph[1].record()
        with L.nvtx_range('phase/forward'):
            self._forward(g, h_in, h_out)
        if ph:
            ph[2].record()
        if L.nvtx_enabled():
            torch.cuda.nvtx.range_push('tmpnn/phase/decode_tracks')
        self._associate(g)
        L.call('tmpnn_graph_decode', g.c, self.index.c, self.frames.c, L.ptr(self.y_out_track),
               L.ptr(self.next_track_id), L.ptr(self.st['t_upto']), 0, L.ptr(self.st['active']), self.R,
               L.ptr(self.keep), self.max_dets, L.ptr(self.decode_scratch), st)
        # plain: move the survivors h_alt -> h_cur; deferred: emit the maps only (sequences that sat the step out
        # are copied h_in -> h_out)
        if self.profile_compact is not None:
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
        L.call('tmpnn_graph_compact', g.c, go.c, L.ptr(self.keep), L.ptr(h_out), L.ptr(h_in),
               L.ptr(self.st['active']), L.ptr(h_out if self.deferred else h_in), self.ldh, L.ptr(self.new_of_old),
               L.ptr(self.compact_scratch), st)
        if self.profile_compact is not None:
            c1.record()
            self.profile_compact.append((c0, c1, g.n_rows.sum(), go.n_rows.sum()))
        if L.nvtx_enabled():
            torch.cuda.nvtx.range_pop()
        if ph:
            ph[3].record()
            self.profile_phases.append(ph)
        L.call('tmpnn_graph_counters', None, None, None, None, L.ptr(self.frames_done), L.ptr(self.st['active']), self.S,
               L.ptr(self.t_dev), st)

    # ---- driver --------------------------------------------------------------------------------
    def reset(self):
        self.ga, self.gb = self._g0, self._g1
        for t in self.st.values():
            t.zero_()
        for g in (self.ga, self.gb):
            g.n_rows.zero_()
        self.ga.status.zero_()
        self.y_out_track.fill_(-1)
        self.next_track_id.zero_()
        self.t_dev.zero_()
        self.edge_updates.zero_()
        self.det_updates.zero_()
        self.frames_done.zero_()
        self.ticks = 0

    def run(self, max_ticks=None):
        """Tracks every sequence to its end (or ``max_ticks`` frames).  Enqueues only; call
        ``results()`` (which synchronises) for the decoded tracks."""
        self.reset()
        self._start()
        # the captured ticks bake in the device pointers of the packed weight images (functional.packed_cells*), which
        # _start() has just re-packed if a parameter changed in place (optimizer.step, load_state_dict), and of every
        # parameter / buffer tensor: a changed key means the captured graph reads freed or stale memory -> re-capture
        key = self._weights_key()
        if self._graph is not None and key != self._graph_key:
            self._graph = None
        n_ticks = self.t_hi if max_ticks is None else min(self.t_hi, int(max_ticks))
        # the two graph sets swap roles every tick -> capture two ticks per CUDA graph replay
        t = 0
        if self.use_cuda_graph and n_ticks >= 4:
            if self._graph is None:
                self._capture()
            # state after capture warm-up was rewound by _capture; replay pairs
            while t + 2 <= n_ticks:
                self._graph.replay()
                t += 2
        while t < n_ticks:
            self._tick(flip=bool(t & 1))
            self.ga, self.gb = self.gb, self.ga
            t += 1
        self.ticks = n_ticks
        return self

    def _capture(self):
        """Captures two consecutive ticks (ga -> gb -> ga) as one CUDA graph.  Capturing does not
        execute, so the engine state is untouched."""
        L.check(L.lib().tmpnn_init())
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._tick(flip=False)
            self.ga, self.gb = self.gb, self.ga
            self._tick(flip=True)
            self.ga, self.gb = self.gb, self.ga
        self._graph = g
        self._graph_key = self._weights_key()

    def _weights_key(self):
        """Identity of everything a captured tick reads from the model: parameter / buffer storage, their versions (the
        packed images are rebuilt when a version changes) and the train / eval switch."""
        ts = list(self.model.parameters()) + list(self.model.buffers())
        return tuple((t.data_ptr(), t._version) for t in ts) + (self.model.training,)

    def results(self):
        """Synchronises; returns (list of y_out [ND_s, 2] int64 arrays, stats dict)."""
        torch.cuda.synchronize(self.dev)
        self.ga.check_status()
        track = self.y_out_track.cpu().numpy().astype(np.int64)
        dp = self.frames.host_det_ptr
        ts = self.frames  # noqa: F841
        outs = []
        for s in range(self.S):
            outs.append(track[dp[s]:dp[s + 1]].copy())
        stats = dict(edge_updates=int(self.edge_updates.item()), det_updates=int(self.det_updates.item()),
                     frames=int(self.frames_done.item()), ticks=self.ticks)
        return outs, stats


def worst_case_dets(frame_counts, window):
    c = np.asarray(frame_counts, dtype=np.int64)
    best = 0
    for t in range(len(c)):
        best = max(best, int(c[max(0, t - window + 1):t + 1].sum()))
    return max(best, 1)
