"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares, and
its host-only helpers agree with the oracle.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    import __graft_entry__ as g
    from batching_pomp_b200 import capi as c
    if not os.path.exists(c.LIB_PATH):
        g.build()
    return c


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "bpomp_cuda.h")).read()
    return sorted(set(re.findall(r"BPOMP_API\s+[\w\s\*]+?\b(bpomp_\w+)\s*\(", src)))


def test_header_symbols_all_exported(capi):
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = _header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export " + n


def test_binding_table_matches_header(capi):
    assert sorted(capi.SIGNATURES) == _header_symbols()


def test_no_cpu_fallback_without_gpu(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.BpompError) as ei:
        capi.Context(7, 0.0264575)
    assert "no CUDA device" in str(ei.value) or "CUDA" in str(ei.value)


def test_create_argument_errors(capi):
    lib = capi.load()
    h = ctypes.c_void_p()
    assert lib.bpomp_create(0, 0, 0.1, ctypes.byref(h)) == -1
    assert lib.bpomp_create(0, 9, 0.1, ctypes.byref(h)) == -1
    assert lib.bpomp_create(0, 7, 0.0, ctypes.byref(h)) == -1
    assert b"segment_length" in lib.bpomp_last_error(None)


def test_gap_bisect_perm_equals_reference_golden(capi):
    """The product's sample order (bpomp_bisect_perm, host side of ctx.cu) against vectors produced by the
    reference's own BisectPerm.hpp (tests/golden/bisect_perm_ref.npz, tests/golden/make_golden_ref.py)."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bisect_perm_ref.npz"))
    ns, offs, gf = z["n"], z["offsets"], z["first"]
    for i, n in enumerate(ns.tolist()):
        np.testing.assert_array_equal(capi.bisect_perm(n), gf[offs[i]:offs[i + 1]], err_msg="n=%d" % n)


@pytest.mark.parametrize("n", list(range(0, 70)) + [100, 127, 128, 129, 255, 300, 1000])
def test_gap_bisect_perm_equals_reference_algorithm(capi, orc, n):
    """The O(n log n) gap formulation (ctx.cu) against the literal O(n^2) restatement of
    util/BisectPerm.hpp:45-89 in the oracle."""
    got = capi.bisect_perm(n)
    ref, _ = orc.bisect_perm(n)
    assert got.tolist() == ref.tolist()
