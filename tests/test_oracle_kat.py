"""Known-answer tests pinning the CPU oracle to the hand-derived values of SURVEY.md §4
(derived from /root/reference/include/batching_pomp/util/BisectPerm.hpp:45-89,
src/BatchingPOMP.cpp:440-475, 510, 548-574).  The reference ships no tests of its own."""
import math

import numpy as np
import pytest


PERM_KAT = {
    1: [0], 2: [0, 1], 3: [1, 0, 2], 4: [1, 0, 2, 3], 5: [2, 0, 1, 3, 4], 6: [2, 4, 0, 1, 3, 5],
    7: [3, 1, 5, 0, 2, 4, 6], 8: [3, 1, 5, 0, 2, 4, 6, 7], 13: [6, 2, 9, 4, 11, 0, 1, 3, 5, 7, 8, 10, 12],
}


@pytest.mark.parametrize("n", sorted(PERM_KAT))
def test_bisect_perm_kat(orc, n):
    first, _ = orc.bisect_perm(n)
    assert first.tolist() == PERM_KAT[n]


def test_bisect_perm_pairs_n7(orc):
    first, second = orc.bisect_perm(7)
    assert list(zip(first.tolist(), second.tolist())) == [(3, 4), (1, 2), (5, 2), (0, 1), (2, 1), (4, 1), (6, 1)]


@pytest.mark.parametrize("n", [1, 2, 3, 10, 17, 64, 100, 257])
def test_bisect_perm_is_permutation(orc, n):
    first, _ = orc.bisect_perm(n)
    assert sorted(first.tolist()) == list(range(n))


def _golden_perm():
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bisect_perm_ref.npz"))
    return z["n"], z["offsets"], z["first"], z["second"]


def test_bisect_perm_matches_reference_golden(orc):
    """Pinned to the reference itself: tests/golden/bisect_perm_ref.npz holds BisectPerm::get(n) of the
    reference's own header compiled in the build container (tests/golden/make_golden_ref.py)."""
    ns, offs, gf, gs = _golden_perm()
    assert len(ns) > 250
    for i, n in enumerate(ns.tolist()):
        first, second = orc.bisect_perm(n)
        np.testing.assert_array_equal(first, gf[offs[i]:offs[i + 1]], err_msg="n=%d" % n)
        np.testing.assert_array_equal(second, gs[offs[i]:offs[i + 1]], err_msg="n=%d" % n)


def test_bisect_perm_matches_reference_build(orc):
    """Live check against oracle/_ref (the reference's BisectPerm.hpp compiled from where it lies); the
    library is built where /root/reference exists and travels to the GPU box with the snapshot."""
    if not orc.ref_available():
        pytest.skip("oracle/_ref/libbpomp_ref.so not built (no /root/reference at build time)")
    for n in list(range(0, 400)) + [777, 1024, 2049, 5000]:
        rf, rs = orc.ref_bisect_perm(n)
        first, second = orc.bisect_perm(n)
        np.testing.assert_array_equal(first, rf, err_msg="n=%d" % n)
        np.testing.assert_array_equal(second, rs, err_msg="n=%d" % n)


def test_nstates_floor_and_minimum(orc):
    L = 0.0264575
    assert orc.nstates(0.0, L) == 2
    assert orc.nstates(2.9 * L, L) == 2
    assert orc.nstates(3.0 * L * (1 + 1e-12), L) == 3
    assert orc.nstates(0.3473738735932844, math.sqrt(7.0) * 0.01) == 13


def test_checked_slots_and_fractions(orc):
    # n=5: slots 1..3 hold fractions 1/6, 2/6, 4/6 (perm [2,0,1,3,4]); slots 0 and 4 never checked.
    a = np.zeros(1)
    b = np.ones(1)
    st = orc.edge_states(a, b, 1.0 / 5.0)  # distance 1, L=0.2 -> n=5
    assert st.shape == (5, 1)
    np.testing.assert_array_equal(st[:, 0], np.array([3, 1, 2, 4, 5]) / 6.0)


@pytest.mark.parametrize("n,slots", [(2, []), (3, [1]), (4, [1]), (5, [1]), (8, [1, 4]), (13, [1, 6, 11]),
                                     (20, [1, 7, 13]), (100, [1, 22, 43, 64, 85])])
def test_belief_query_slots(orc, n, slots):
    step = orc.belief_step(n)
    got = list(range(1, n - 1, step)) if n > 2 else []
    assert got == slots


def test_radius_functions(orc):
    assert orc.halton_radius(100, 2) == 0.25
    assert orc.halton_radius(10 ** 4, 2) == pytest.approx(0.025, rel=1e-15)
    assert orc.halton_radius(10 ** 6, 2) == pytest.approx(0.0025, rel=1e-15)
    assert orc.halton_radius(100, 7) == pytest.approx(1.294868669807803, rel=1e-14)
    assert orc.halton_radius(10 ** 6, 7) == pytest.approx(0.3473738735932844, rel=1e-14)
    assert orc.rgg_radius(10 ** 4, 2) == pytest.approx(0.04194097563613211, rel=1e-13)
    assert orc.rgg_radius(10 ** 6, 7) == pytest.approx(0.33017274622329834, rel=1e-13)
    assert math.pow(2.0, 1 / 2.0) == 1.4142135623730951
    assert math.pow(2.0, 1 / 7.0) == 1.1040895136738123


def test_empty_belief_model_is_prior(orc):
    q = np.random.default_rng(0).random((5, 7))
    est = orc.belief_estimate(np.zeros((0, 7)), np.zeros(0), q)
    np.testing.assert_array_equal(est, 0.5)
    frm = np.zeros((1, 7))
    to = np.full((1, 7), 0.13)  # dist = 0.3439..., L=0.0264575 -> n = 13 -> 3 query slots
    pf, nl = orc.belief_edge_pfree(frm, to, math.sqrt(7.0) * 0.01, np.zeros((0, 7)), np.zeros(0))
    assert nl[0] == 3 and pf[0] == 0.125


def test_knn_estimate_formula(orc):
    pts = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 2.0]])
    vals = np.array([1.0, 0.0, 1.0])
    q = np.array([[0.0, 0.0]])
    # distances 0 (clamped to 1e-4), 1, 2 -> weights 1e4, 1, 0.5
    w = np.array([1.0 / 0.0001, 1.0, 0.5])
    dot = 0.0
    s = 0.0
    for wi, vi in zip(w, vals):
        dot += wi * vi
        s += wi
    expect = (dot / s + 0.25 * 0.5) / (1 + 0.25)
    assert orc.belief_estimate(pts, vals, q)[0] == expect
    # k limits the neighbour set
    expect1 = (1.0 + 0.125) / 1.25
    assert orc.belief_estimate(pts, vals, q, K=1)[0] == expect1


def test_edge_check_semantics_boxes(orc):
    d = 2
    L = 0.1
    lo = np.array([[0.45, -1.0]])
    hi = np.array([[0.55, 1.0]])
    frm = np.array([[0.0, 0.0], [0.0, 0.0], [0.0, 0.0]])
    to = np.array([[1.0, 0.0], [0.25, 0.0], [0.4, 0.0]])
    valid, first, ns = orc.edges_check_boxes(frm, to, L, lo, hi)
    # edge 0: n=10, perm(10) slot order; fraction 5/11=0.4545 and 6/11=0.545 are inside the wall.
    assert ns.tolist() == [10, 2, 4]
    perm10, _ = orc.bisect_perm(10)
    fr = (1 + perm10) / 11.0
    inside = (fr >= 0.45) & (fr <= 0.55)
    slots = [i for i in range(1, 9) if inside[i]]
    assert valid[0] == 0 and first[0] == slots[0]
    # edge 1: n=2 -> FREE unchecked; edge 2: n=4, checked fractions stay below 0.45
    assert valid[1] == 1 and first[1] == -1
    assert valid[2] == 1 and first[2] == -1


def test_slot0_and_last_never_checked(orc):
    # n=3: perm [1,0,2]; only slot 1 (fraction 1/4) is checked; the midpoint (slot 0, fraction 2/4) is not.
    lo = np.array([[0.49]])
    hi = np.array([[0.51]])
    valid, first, ns = orc.edges_check_boxes(np.array([[0.0]]), np.array([[1.0]]), 1.0 / 3.0, lo, hi)
    assert ns[0] == 3 and valid[0] == 1 and first[0] == -1


def test_image_world_convention(orc):
    pix = np.full((4, 8), 255, dtype=np.uint8)
    pix[3, 0] = 0      # lower-left pixel black: x in [0,1/8), y in [3/4,1]
    pix[0, 7] = 127    # below threshold -> invalid
    st = np.array([[0.01, 0.99], [0.99, 0.01], [0.5, 0.5], [1.0, 1.0], [1.0000001, 0.5], [-1e-9, 0.5],
                   [0.0, 1.0]])
    got = orc.states_valid_image(st, pix)
    assert got.tolist() == [0, 0, 1, 1, 0, 0, 0]


def test_halton_matches_definition(orc):
    h = orc.halton(1, 9, 3)
    np.testing.assert_allclose(h[:3, 0], [0.5, 0.25, 0.75], rtol=0, atol=0)
    np.testing.assert_allclose(h[:3, 1], [1 / 3, 2 / 3, 1 / 9], rtol=1e-15)
    np.testing.assert_allclose(h[4, 2], 0.04, rtol=1e-15)  # 5 in base 5 = "10" -> 1/25


def test_neighbors_inclusive_sorted(orc):
    verts = np.array([[0.0, 0.0], [0.3, 0.0], [0.0, 0.3], [0.5, 0.0], [0.3, 0.0]])
    row_ptr, col, dist = orc.neighbors_radius(verts, [0], 0.3)
    # inclusive radius, ascending (dist, id): self first, ties 1,2,4 by id
    assert col.tolist() == [0, 1, 2, 4]
    assert dist.tolist() == [0.0, 0.3, 0.3, 0.3]
    active = np.array([1, 0, 1, 1, 1], dtype=np.uint8)
    _, col2, _ = orc.neighbors_radius(verts, [0], 0.3, active=active)
    assert col2.tolist() == [0, 2, 4]
