"""CPU-side checks of the C ABI: the library loads, exports every symbol include/vos_prop.h
declares, and its host-only entry points (sample_frames, plan_step, decomposition) are right.
No compute calls: there is no GPU here."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import propagation_oracle as O
from tests import _golden as G

REPO = Path(__file__).resolve().parent.parent


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as ge
    ge.build()
    from vosb200 import _capi
    return _capi.lib()


def test_exports_every_declared_symbol(lib):
    from vosb200 import _capi
    header = (REPO / 'include' / 'vos_prop.h').read_text()
    declared = set(re.findall(r'\b(vosprop_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations parsed'
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.vosprop_abi_version() == 4


def test_struct_layout_matches_header(lib):
    from vosb200 import _capi
    # int32 x2, int32[32], float[32], float, int32 x4, 6 pointers
    assert C.sizeof(_capi.Step) == 8 + 128 + 128 + 4 + 16 + 4 + 48  # incl. 4 bytes padding before pointers
    assert C.sizeof(_capi.Config) == 16


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from vosb200 import _capi
    h = C.c_void_p()
    rc = lib.vosprop_create(C.byref(_capi.Config(0, 6420, 48, 0)), C.byref(h))
    assert rc == _capi.ERR_UNSUPPORTED and not h.value
    assert b'no CPU fallback' in lib.vosprop_last_error()
    from vosb200 import PropagationEngine
    with pytest.raises(RuntimeError):
        PropagationEngine(max_pixels=100)


def test_sample_frames_matches_reference_table(lib):
    from vosb200 import sample_frames
    table = np.load(G.GOLDEN / 'sample_frames.npz')['table']
    for row in table:
        ref_num, rng, t, n = (int(v) for v in row[:4])
        assert sample_frames(t, rng, ref_num) == row[4:4 + n].tolist(), (ref_num, rng, t)
    with pytest.raises(ValueError):
        sample_frames(5, 40, 2)          # the reference raises here too
    assert sample_frames(2, 40, 2) == [0, 1]


def test_plan_refs_sigma_rule(lib):
    from vosb200 import plan_refs
    for (t, n) in ((3, 9), (10, 9), (15, 9), (16, 9), (40, 9), (20, 3), (20, 4), (33, 12)):
        refs, sig = plan_refs(t, 40, n, 8.0, 21.0, False)
        assert refs == O.sample_frames(t, 40, n)
        assert sig == O.ref_sigmas(t, len(refs), 8.0, 21.0)
        _, sig0 = plan_refs(t, 40, n, 8.0, 21.0, True)
        assert all(s == 0.0 for s in sig0)


@pytest.mark.parametrize('P,R,sms', [(6420, 9, 148), (6420, 1, 148), (240, 1, 148), (240, 9, 148), (32400, 9, 148),
                                     (100, 3, 148), (6420, 20, 148), (129600, 2, 148), (6420, 9, 7)])
def test_stream_k_decomposition_covers_every_tile_once(lib, P, R, sms):
    grid, segs = C.c_int32(), C.c_int32()
    begin = (C.c_int64 * (sms + 1))()
    assert lib.vosprop_debug_decompose(P, R, sms, C.byref(grid), begin, C.byref(segs)) == 0
    tpf = (P + 127) // 128
    nt, total = R * tpf, R * tpf * tpf
    g = grid.value
    assert g == min(sms, total) and begin[0] == 0 and begin[g] == total
    sizes = [begin[c + 1] - begin[c] for c in range(g)]
    assert min(sizes) >= 1 and max(sizes) - min(sizes) <= 1          # balanced, contiguous, complete
    for c in range(g):                                                 # segments per CTA within the bound
        first_m, last_m = begin[c] // nt, (begin[c + 1] - 1) // nt
        assert last_m - first_m + 1 <= segs.value
