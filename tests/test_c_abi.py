"""CPU: the drop-in boundary itself.  ``include/tmpnn.h`` is plain C (compiles as C99 with gcc), the built
``libtmpnn_sm100a.so`` loads without a GPU and exports every function the header declares, the ctypes stub
(``trackmpnn_b200/_lib.py``) binds exactly that set, and the host-only entry points (version, error slot, scratch-size
queries) answer.  No compute call is made here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'tmpnn.h')


def _declared():
    """Function names declared in the header (return type at the start of a line, ``tmpnn_*(``)."""
    text = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r'^\s*(?:const\s+)?[A-Za-z_][\w\s]*?[\s\*](tmpnn_\w+)\s*\(', text, flags=re.M)))


def _lib_path():
    from trackmpnn_b200 import _lib as L
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return L.LIB_PATH


def test_header_is_plain_c(tmp_path):
    src = tmp_path / 'use_header.c'
    src.write_text('#include "tmpnn.h"\nint main(void) { tmpnn_graph g; tmpnn_index ix; (void)g; (void)ix; return TMPNN_OK; }\n')
    r = subprocess.run(['gcc', '-std=c99', '-pedantic', '-Wall', '-Werror', '-fsyntax-only', '-I', os.path.dirname(HEADER), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_library_exports_every_declared_function():
    names = _declared()
    assert len(names) >= 50 and 'tmpnn_last_error' in names and 'tmpnn_mp_step_fwd' in names
    lib = C.CDLL(_lib_path())
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f'declared in include/tmpnn.h but not exported: {missing}'


def test_ctypes_stub_binds_the_declared_set():
    from trackmpnn_b200 import _lib as L
    bound = set(L.exported_symbols())
    declared = set(_declared())
    assert bound <= declared, f'bound but not declared: {sorted(bound - declared)}'
    # everything declared is either bound or a host-only helper the Python side has no use for
    assert declared - bound <= {'tmpnn_version', 'tmpnn_init', 'tmpnn_lsap_scratch_bytes'}, sorted(declared - bound)
    for name in L.exported_symbols():
        assert name in L.KERNELS_PER_CALL or name.endswith(('_bytes', '_ints', '_floats')) or name in (
            'tmpnn_last_error', 'tmpnn_init', 'tmpnn_version'), f'{name}: no launch count for bench.py gpu_launches'


def test_host_only_entry_points_answer_without_a_gpu():
    lib = C.CDLL(_lib_path())
    lib.tmpnn_last_error.restype = C.c_char_p
    lib.tmpnn_version.restype = C.c_int
    assert lib.tmpnn_version() > 0
    assert isinstance(lib.tmpnn_last_error(), bytes)
    lib.tmpnn_tc_tile_table_bytes.restype = C.c_size_t
    lib.tmpnn_tc_tile_table_bytes.argtypes = [C.c_int, C.c_int]
    assert lib.tmpnn_tc_tile_table_bytes(2, 1000) == 2 * 8 * 16 + 16       # 128-row tiles, one int4 each (+ one spare)
    lib.tmpnn_gru_pack_floats.restype = C.c_longlong
    lib.tmpnn_gru_pack_floats.argtypes = [C.c_int]
    assert lib.tmpnn_gru_pack_floats(128) > lib.tmpnn_gru_pack_floats(64) > 2 * 192 * 64
    # a NULL argument is refused with an error code and a message, never a crash
    lib.tmpnn_aggregate_dets.restype = C.c_int
    lib.tmpnn_aggregate_dets.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    rc = lib.tmpnn_aggregate_dets(None, None, None, 64, 0, None, None)
    assert rc < 0 and b'null' in lib.tmpnn_last_error().lower()


def test_every_entry_point_has_an_nvtx_family():
    """NVTX ranges are named tmpnn/<family>/<entry point>; no C-ABI call may fall into 'misc' (SURVEY.md section 5)."""
    from trackmpnn_b200 import _lib as L
    skip = {'tmpnn_version', 'tmpnn_init', 'tmpnn_last_error'}
    sizes = {n for n in L.exported_symbols() if n.endswith(('_bytes', '_ints', '_floats'))}
    for name in L.exported_symbols():
        if name in skip or name in sizes:
            continue
        assert L.family_of(name) != 'misc', name
    assert not L.nvtx_enabled()
    L.enable_nvtx(True)
    try:
        with L.nvtx_range('phase/forward'):   # works without a GPU or a profiler attached
            pass
    finally:
        L.enable_nvtx(False)
