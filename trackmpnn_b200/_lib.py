"""ctypes binding of ``libtmpnn_sm100a.so`` (the C ABI declared in ``include/tmpnn.h``).

There is deliberately no fallback: if the library is missing or a call fails, the
product path raises.  PyTorch is used only to own device memory and streams; every
argument that crosses this boundary is a raw device pointer or a plain integer.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libtmpnn_sm100a.so')

TILE_ROWS = 64
HIDDEN = 64

FLAG_TC_RANGE = 128
NOTE_TC_RANGE_RERUN = 512   # not an error: a tensor-core step left the fp16 split's range and was re-run on the FMA kernel
FLAG_NAMES = {
    1: 'row capacity of a slab exceeded (tmpnn_graph_append)',
    2: 'detection-row capacity of the index exceeded',
    4: 'incidence capacity of the index exceeded',
    8: 'a detection has more incident edges than one CTA can sort',
    16: 'More than one GT edge from same node!',
    32: 'more detection rows in one window than the decode walk holds',
    64: 'tensor-core kernel: mbarrier wait timed out (results invalid)',
    128: 'tensor-core kernel: activation exceeds the fp16 split range (use the FMA path)',
    256: 'structured index build: the window graph is not a chain of dense edge blocks (use tmpnn_index_build)',
}

i32p = C.POINTER(C.c_int32)
f32p = C.POINTER(C.c_float)


class Graph(C.Structure):
    _fields_ = [('num_seqs', C.c_int32), ('cap_rows', C.c_int32), ('n_rows', C.c_void_p), ('ts', C.c_void_p),
                ('det', C.c_void_p), ('ass', C.c_void_p), ('src', C.c_void_p), ('dst', C.c_void_p),
                ('label', C.c_void_p), ('score', C.c_void_p), ('logit', C.c_void_p), ('status', C.c_void_p),
                ('phys', C.c_void_p), ('psrc', C.c_void_p), ('pdst', C.c_void_p), ('phys_end', C.c_void_p)]


class InputGroup(C.Structure):   # tmpnn_input_group
    _fields_ = [('a', C.c_void_p), ('mean', C.c_void_p), ('var', C.c_void_p), ('x_idx', C.c_void_p), ('out_rows', C.c_void_p),
                ('scratch', C.c_void_p), ('n', C.c_int32), ('n_edge_rows', C.c_int32)]


class Index(C.Structure):
    _fields_ = [('cap_dets', C.c_int32), ('cap_inc', C.c_int32), ('n_dets', C.c_void_p), ('n_edges', C.c_void_p),
                ('det_rows', C.c_void_p), ('det_of_row', C.c_void_p), ('seq_det_ptr', C.c_void_p),
                ('seg_ptr', C.c_void_p), ('inc', C.c_void_p), ('tile_ptr', C.c_void_p), ('tile128_ptr', C.c_void_p),
                ('scratch', C.c_void_p)]


class Frames(C.Structure):
    _fields_ = [('t_max', C.c_int32), ('ldx', C.c_int32), ('frame_ptr', C.c_void_p), ('frame_dets', C.c_void_p),
                ('det_ptr', C.c_void_p), ('det_track', C.c_void_p)]


SEQ_STATE_FIELDS = ('phase', 'skip_until', 't_end', 'active', 't_upto', 'fresh', 'last_new')


class SeqState(C.Structure):
    _fields_ = [('phase', C.c_void_p), ('skip_until', C.c_void_p), ('t_end', C.c_void_p), ('active', C.c_void_p),
                ('t_upto', C.c_void_p), ('fresh', C.c_void_p), ('last_new', C.c_void_p)]


class TmpnnError(RuntimeError):
    pass


_lib = None

_VP = C.c_void_p
_I = C.c_int
_PROTOS = {
    'tmpnn_version': ([], C.c_int),
    'tmpnn_init': ([], C.c_int),
    'tmpnn_gru_pack_floats': ([_I], C.c_size_t),
    'tmpnn_pack_gru': ([_VP] * 6 + [_I, _VP, _VP], _I),
    'tmpnn_input_linear1': ([_VP, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _I, _VP], _I),
    'tmpnn_input_bn_stats': ([_VP, _I, _I, _VP, _VP, _VP, _VP, _VP], _I),
    'tmpnn_input_bn_relu_linear2': ([_VP] * 8 + [_I, _I, _VP, _VP, _I, _VP], _I),
    'tmpnn_index_scratch_ints': ([_I, _I, _I], C.c_size_t),
    'tmpnn_index_build': ([C.POINTER(Graph), C.POINTER(Index), _VP, _VP], _I),
    'tmpnn_index_structured_scratch_bytes': ([_I, _I], C.c_size_t),
    'tmpnn_index_build_structured': ([C.POINTER(Graph), C.POINTER(Index), _VP, _VP, _VP], _I),
    'tmpnn_aggregate_dets': ([C.POINTER(Graph), C.POINTER(Index), _VP, _I, _I, _VP, _VP], _I),
    'tmpnn_aggregate_blocks_scratch_bytes': ([_I, _I, _I], C.c_size_t),
    'tmpnn_aggregate_dets_blocks': ([C.POINTER(Graph), C.POINTER(Index), _VP, _VP, _I, _I, _VP, _VP, _I, _I, _VP], _I),
    'tmpnn_gat_aggregate_dets': ([C.POINTER(Graph), C.POINTER(Index), _VP, _I, _I, _VP, _VP, _I, _I, _VP, _VP, _VP, _VP, _VP], _I),
    'tmpnn_gat_aggregate_dets_train': ([C.POINTER(Graph), C.POINTER(Index), _VP, _I, _I, _VP, _VP, _I, _I, _VP, C.c_float] + [_VP] * 6, _I),
    'tmpnn_gat_bwd': ([C.POINTER(Graph), C.POINTER(Index), _I, _VP, _I, _I, _VP, _VP, _I, _VP, C.c_float] + [_VP] * 13, _I),
    'tmpnn_aggregate_edges': ([C.POINTER(Graph), C.POINTER(Index), _VP, _I, _I, _I, _VP, _VP], _I),
    'tmpnn_mp_step_fwd': ([C.POINTER(Graph), C.POINTER(Index), _VP, _VP, _I, _I, _I, _I, _VP, _VP, _VP, _VP], _I),
    'tmpnn_mp_edge_fwd': ([C.POINTER(Graph), C.POINTER(Index), _VP, _VP, _I, _I, _I, _I, _VP, _VP], _I),
    'tmpnn_mp_edge_fwd_on_flag': ([C.POINTER(Graph), C.POINTER(Index), _VP, _VP, _I, _I, _I, _I, _VP, _I, _VP], _I),
    'tmpnn_graph_counters': ([_VP, _VP, _VP, _VP, _VP, _VP, _I, _VP, _VP], _I),
    'tmpnn_graph_force_det_scores': ([C.POINTER(Graph), C.POINTER(Index), _VP], _I),
    'tmpnn_status_ack': ([C.POINTER(Graph), _I, _I, _VP], _I),
    'tmpnn_mp_det_fwd': ([C.POINTER(Graph), C.POINTER(Index), _VP, _VP, _I, _I, _I, _VP, _VP, _VP], _I),
    'tmpnn_tc_det_tile_table_bytes': ([_I, _I], C.c_size_t),
    'tmpnn_mp_det_fwd_tc': ([C.POINTER(Graph), C.POINTER(Index), _VP, _VP, _VP, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP], _I),
    'tmpnn_mp_det_fwd_on_flag': ([C.POINTER(Graph), C.POINTER(Index), _VP, _VP, _I, _I, _I, _VP, _VP, _I, _VP], _I),
    'tmpnn_gru_tc_pack_bytes': ([], C.c_size_t),
    'tmpnn_pack_gru_tc': ([_VP] * 6 + [_I, _VP, _VP], _I),
    'tmpnn_tc_tile_table_bytes': ([_I, _I], C.c_size_t),
    'tmpnn_mp_edge_fwd_tc_pre': ([C.POINTER(Graph), C.POINTER(Index), _VP, _VP, _I, _I, _I, _I, _VP, _VP, _VP, _VP, _VP], _I),
    'tmpnn_mp_edge_fwd_tc': ([C.POINTER(Graph), C.POINTER(Index), _VP, _VP, _I, _I, _I, _VP, _VP], _I),
    'tmpnn_mp_edge_fwd_tc_train': ([C.POINTER(Graph), C.POINTER(Index), _VP, _VP, _I, _I, _I, _VP, _VP, _VP], _I),
    'tmpnn_mp_det_fwd_train': ([C.POINTER(Graph), C.POINTER(Index), _VP, _VP, _I, _I, _I, _VP, _VP, _VP, _VP], _I),
    'tmpnn_mp_step_fwd_train': ([C.POINTER(Graph), C.POINTER(Index), _VP, _VP, _I, _I, _I, _I, _VP, _VP, _VP, _VP, _VP], _I),
    'tmpnn_mp_step_fwd_train_agg': ([C.POINTER(Graph), C.POINTER(Index), _VP, _VP, _I, _I, _I, _I, _VP, _VP, _VP, _VP, _VP], _I),
    'tmpnn_gate_bwd_partial_floats': ([], C.c_size_t),
    'tmpnn_gate_bwd': ([_I] + [_VP] * 4 + [_I, _I] + [_VP] * 17, _I),
    'tmpnn_rows_times_w': ([_VP, _I, _VP, _VP, _VP, _VP, _VP, _I, _VP, _I, _I, _VP], _I),
    'tmpnn_rows_outer': ([_VP, _I, _VP, _VP, _VP, _VP, _VP, _I, _I, _VP, _VP], _I),
    'tmpnn_bwd_tc_image_bytes': ([], C.c_size_t),
    'tmpnn_bwd_tc_partial_floats': ([], C.c_size_t),
    'tmpnn_pack_w_tc': ([_VP, _I, _I, _VP, _VP], _I),
    'tmpnn_rows_gemm_tc': ([_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _VP, _I, _VP, _VP, _I, _VP, _VP], _I),
    'tmpnn_scatter_bwd': ([C.POINTER(Graph), C.POINTER(Index), _I, _VP, _VP, _I, _VP, _VP, _I, _I, _VP], _I),
    'tmpnn_input_bwd': ([_VP, _I, _I, _I] + [_VP] * 9 + [_I, _I, _VP, _I, _I, _I] + [_VP] * 8, _I),
    'tmpnn_input_bn_groups_fwd': ([C.POINTER(InputGroup), _I] + [_VP] * 8 + [_I, _I, _VP], _I),
    'tmpnn_input_bwd_partial_floats': ([_I], C.c_size_t),
    'tmpnn_input_bwd_groups': ([_VP, _I, _I, _I, C.POINTER(InputGroup), _I] + [_VP] * 5 + [_I, _I, _I] + [_VP] * 8, _I),
    'tmpnn_build_features': ([_VP, _I, _I, _I, _I, _I, _VP, _VP, _VP, _I, _VP], _I),
    'tmpnn_loss_targets': ([C.POINTER(Graph), C.POINTER(Index), _I, _VP, _VP], _I),
    'tmpnn_loss_ce_fwd': ([C.POINTER(Index), _I] + [_VP] * 7, _I),
    'tmpnn_loss_ce_bwd': ([C.POINTER(Graph), C.POINTER(Index), _I] + [_VP] * 6, _I),
    'tmpnn_loss_wbce_fwd': ([_I] + [_VP] * 6, _I),
    'tmpnn_loss_wbce_bwd': ([_I] + [_VP] * 6, _I),
    'tmpnn_rows_move': ([_VP] * 5 + [_I, _I, _I, _VP], _I),
    'tmpnn_loss_focal_fwd': ([_I] + [_VP] * 5, _I),
    'tmpnn_loss_focal_bwd': ([_I] + [_VP] * 5, _I),
    'tmpnn_ypred_unpack': ([_VP, _I, _VP, _VP, _VP, _VP], _I),
    'tmpnn_ypred_pack': ([_VP, _VP, _VP, _I, _VP, _VP], _I),
    'tmpnn_coo_from_edges': ([_VP, _VP, _VP, _I, _I, _VP, _VP, C.c_int64, _VP, _VP], _I),
    'tmpnn_edges_from_coo': ([_VP, _VP, C.c_int64, _I, _VP, _VP, _VP], _I),
    'tmpnn_graph_associate': ([C.POINTER(Graph), C.POINTER(Index), _I, _VP, _VP], _I),
    'tmpnn_hungarian_scratch_bytes': ([_I, _I], C.c_size_t),
    'tmpnn_graph_associate_hungarian': ([C.POINTER(Graph), C.POINTER(Index), _VP, _VP, _I, _I, _I, C.c_float, _VP, _VP], _I),
    'tmpnn_lsap_scratch_bytes': ([_I, _I, _I], C.c_size_t),
    'tmpnn_lsap_solve': ([_VP, _I, _I, _I, _VP, _VP, _VP], _I),
    'tmpnn_graph_append_scratch_ints': ([_I, _I], C.c_size_t),
    'tmpnn_graph_append': ([C.POINTER(Graph), C.POINTER(Frames), C.POINTER(SeqState), _VP, _I, _I, _I, _VP, _I,
                            _VP, _VP, _VP, _I, _VP, _VP, _VP], _I),
    'tmpnn_graph_decode_scratch_ints': ([_I, _I], C.c_size_t),
    'tmpnn_graph_decode': ([C.POINTER(Graph), C.POINTER(Index), C.POINTER(Frames), _VP, _VP, _VP, _I, _VP, _I,
                            _VP, _I, _VP, _VP], _I),
    'tmpnn_graph_prune_mask': ([C.POINTER(Graph), C.POINTER(Index), _I, _I, C.c_float, _VP, _VP, _VP], _I),
    'tmpnn_graph_phys_identity': ([C.POINTER(Graph), _VP], _I),
    'tmpnn_graph_compact_scratch_ints': ([_I, _I], C.c_size_t),
    'tmpnn_graph_compact': ([C.POINTER(Graph), C.POINTER(Graph), _VP, _VP, _VP, _VP, _VP, _I, _VP, _VP, _VP], _I),
}


def lib():
    """Loads the CUDA library once; raises if it was not built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TmpnnError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; '
                             f'g.build()"` (nvcc, sm_100a).  trackmpnn_b200 has no CPU or PyTorch fallback path.')
        l = C.CDLL(os.environ.get('TMPNN_LIB', LIB_PATH))  # TMPNN_LIB: a debug build (profiles/trace_tc.py)
        l.tmpnn_last_error.restype = C.c_char_p
        for name, (args, res) in _PROTOS.items():
            f = getattr(l, name)
            f.argtypes = args
            f.restype = res
        _lib = l
    return _lib


def exported_symbols():
    return ['tmpnn_last_error'] + list(_PROTOS)


def check(rc):
    if rc != 0:
        msg = lib().tmpnn_last_error().decode()
        if rc == -1 and ('must be lesser' in msg or 'Only batch size' in msg):
            raise AssertionError(msg)
        raise TmpnnError(f'libtmpnn error {rc}: {msg}')


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), 'libtmpnn needs contiguous CUDA tensors'
    return t.data_ptr()


def stream():
    # raw handle of torch's current stream (the python Stream object costs ~20 us per call; a training chunk makes
    # ~700 of these calls)
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def launch_count():
    return _LAUNCHES[0]


_LAUNCHES = [0]
# kernels launched per C-ABI call (for bench.py's gpu_launches accounting)
KERNELS_PER_CALL = {
    'tmpnn_pack_gru': 1, 'tmpnn_input_linear1': 1, 'tmpnn_input_bn_stats': 1, 'tmpnn_input_bn_relu_linear2': 1,
    'tmpnn_index_build': 9, 'tmpnn_gat_aggregate_dets': 3, 'tmpnn_index_build_structured': 11, 'tmpnn_aggregate_dets': 1, 'tmpnn_aggregate_dets_blocks': 3, 'tmpnn_aggregate_edges': 1, 'tmpnn_mp_step_fwd': 3, 'tmpnn_mp_edge_fwd': 1, 'tmpnn_mp_det_fwd': 1, 'tmpnn_mp_det_fwd_tc': 3, 'tmpnn_mp_det_fwd_on_flag': 1, 'tmpnn_mp_edge_fwd_tc': 1, 'tmpnn_mp_edge_fwd_tc_pre': 3, 'tmpnn_pack_gru_tc': 1,
    'tmpnn_ypred_unpack': 1, 'tmpnn_ypred_pack': 1, 'tmpnn_coo_from_edges': 5, 'tmpnn_edges_from_coo': 1,
    'tmpnn_graph_associate': 2, 'tmpnn_graph_append': 4, 'tmpnn_graph_decode': 2, 'tmpnn_graph_prune_mask': 2,
    'tmpnn_graph_compact': 4, 'tmpnn_graph_phys_identity': 1, 'tmpnn_mp_edge_fwd_on_flag': 1, 'tmpnn_graph_force_det_scores': 1,
    'tmpnn_status_ack': 1, 'tmpnn_graph_counters': 1,
    'tmpnn_mp_step_fwd_train': 3, 'tmpnn_mp_edge_fwd_tc_train': 1, 'tmpnn_mp_det_fwd_train': 1, 'tmpnn_mp_step_fwd_train_agg': 2, 'tmpnn_gat_aggregate_dets_train': 3, 'tmpnn_gat_bwd': 4, 'tmpnn_gate_bwd': 2, 'tmpnn_rows_times_w': 1, 'tmpnn_rows_outer': 1, 'tmpnn_rows_gemm_tc': 2, 'tmpnn_pack_w_tc': 1, 'tmpnn_scatter_bwd': 2,
    'tmpnn_build_features': 1, 'tmpnn_input_bwd': 1, 'tmpnn_input_bwd_groups': 2, 'tmpnn_input_bn_groups_fwd': 3, 'tmpnn_loss_targets': 2, 'tmpnn_loss_ce_fwd': 2, 'tmpnn_loss_ce_bwd': 1, 'tmpnn_loss_focal_fwd': 2, 'tmpnn_loss_wbce_fwd': 2, 'tmpnn_loss_wbce_bwd': 1, 'tmpnn_rows_move': 1,
    'tmpnn_loss_focal_bwd': 1, 'tmpnn_graph_associate_hungarian': 2, 'tmpnn_lsap_solve': 1,
}


# NVTX ranges per kernel family (SURVEY.md section 5 "tracing"): set TMPNN_NVTX=1 (or call enable_nvtx()) and every C-ABI
# call is bracketed by a range "tmpnn/<family>/<entry point>" -- families: input, index, aggregate, mp_step, graph, loss,
# backward, pack, convert -- which a profiler with NVTX support groups the launch list by; off by default (two extra
# Python calls per launch otherwise).  TrackEngine adds the reference's three phases (update / forward / decode) on top.
_FAMILY_PREFIXES = (('tmpnn_input_', 'input'), ('tmpnn_index_', 'index'), ('tmpnn_aggregate', 'aggregate'),
                    ('tmpnn_gat_', 'aggregate'), ('tmpnn_mp_', 'mp_step'), ('tmpnn_graph_', 'graph'), ('tmpnn_status_', 'graph'),
                    ('tmpnn_loss_', 'loss'), ('tmpnn_gate_bwd', 'backward'), ('tmpnn_rows_', 'backward'), ('tmpnn_pack_w_tc', 'backward'),
                    ('tmpnn_scatter_bwd', 'backward'), ('tmpnn_pack_', 'pack'), ('tmpnn_ypred_', 'convert'),
                    ('tmpnn_coo_', 'convert'), ('tmpnn_edges_', 'convert'), ('tmpnn_build_features', 'input'),
                    ('tmpnn_lsap_', 'graph'))
_NVTX = [os.environ.get('TMPNN_NVTX', '0') not in ('', '0')]


def enable_nvtx(on=True):
    _NVTX[0] = bool(on)


def nvtx_enabled():
    return _NVTX[0]


def family_of(name):
    for prefix, fam in _FAMILY_PREFIXES:
        if name.startswith(prefix):
            return fam
    return 'misc'


class nvtx_range:
    """``with nvtx_range('forward'):`` -- a no-op unless NVTX ranges are enabled."""

    def __init__(self, label):
        self.label = label

    def __enter__(self):
        if _NVTX[0]:
            torch.cuda.nvtx.range_push('tmpnn/' + self.label)

    def __exit__(self, *exc):
        if _NVTX[0]:
            torch.cuda.nvtx.range_pop()


def call(name, *args):
    _LAUNCHES[0] += KERNELS_PER_CALL.get(name, 0)
    if _NVTX[0]:
        torch.cuda.nvtx.range_push(f'tmpnn/{family_of(name)}/{name}')
        try:
            check(getattr(lib(), name)(*args))
        finally:
            torch.cuda.nvtx.range_pop()
        return
    check(getattr(lib(), name)(*args))


class ZeroArena:
    """One pre-allocated float32 buffer that the batched trainer zeroes ONCE per optimizer step and the step's small
    zero-initialised scratch tensors (loss accumulators, per-segment sums, the CE gradient) are carved from, instead of one
    fill kernel each (~40 per step).  ``zeros()`` falls back to ``torch.zeros`` when no arena is active or it is full."""
    active = None

    def __init__(self, device, nfloats=1 << 23):
        self.buf = torch.zeros(nfloats, dtype=torch.float32, device=device)
        self.off = 0

    def begin(self):
        self.buf.zero_()
        self.off = 0
        ZeroArena.active = self

    @staticmethod
    def end():
        ZeroArena.active = None


def zeros(n, device):
    """n zero floats on ``device``: a 16-byte aligned slice of the active ZeroArena, else a fresh tensor."""
    a = ZeroArena.active
    n = int(n)
    if a is not None and a.buf.device == torch.device(device) and a.off + n <= a.buf.numel():
        out = a.buf[a.off:a.off + n]
        a.off += (n + 3) // 4 * 4
        return out
    return torch.zeros(n, dtype=torch.float32, device=device)

