"""Pins the CPU oracle — and through it the CUDA path — to the REFERENCE'S OWN CODE.

oracle/Makefile (target `ref`) compiles /root/reference/src/BatchingPOMP.cpp with every header it includes
(cspacebelief/KNNModel.hpp, BeliefPoint.hpp, util/Selector.hpp, util/BisectPerm.hpp, util/RoadmapFromFile.hpp,
batching/*.hpp) against the API stand-ins of oracle/standins into oracle/_ref/libbpomp_ref.so.
tests/golden/make_golden_ref.py ran that build on the scenarios of tests/refcases.py and committed the records
(tests/golden/planner_ref_*.npz, knn_ref.npz).  Here:
  * the oracle planner reproduces every record: status, best cost, alpha, radius, batch count, graph size, the
    four counters and the path states after EVERY solve() call, bit for bit (CPU);
  * where the reference build is present (build container; it also travels to the GPU box), the reference is run
    again live and must still produce the records, and KNNModel / Selector are compared directly (CPU);
  * the GPU planner reproduces the same records, and the CUDA belief estimates match KNNModel::estimate within
    1e-12 relative (-m gpu).
"""
import os
import sys
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refcases  # noqa: E402

G = os.path.join(HERE, "golden")
CASES = sorted(refcases.cases().keys())
LIVE = ["hybrid_image_2d", "vertex_boxes_7d", "single_image_2d"]  # the quick ones for the live re-run


def _golden(name, d):
    return refcases.from_npz(np.load(os.path.join(G, "planner_ref_%s.npz" % name), allow_pickle=False), d)


@pytest.mark.parametrize("name", CASES)
def test_oracle_planner_reproduces_reference_runs(orc, name):
    from batching_pomp_b200 import workloads as wl
    case = refcases.cases()[name]
    roadmap = wl.halton(case["N"], case["d"])
    op = orc.OraclePlanner(case["d"], wl.segment_length(case["d"], case["fraction"]), wl.max_extent(case["d"]), 1.0)
    rec = refcases.drive(op, case, lambda p: p.set_roadmap(roadmap, case["edges"]))
    assert refcases.same(_golden(name, case["d"]), rec) is None
    assert any(n > 0 for n in rec["path_len"])


@pytest.mark.parametrize("name", LIVE)
def test_reference_build_reproduces_its_goldens(orc, name):
    if not orc.ref_planner_available():
        pytest.skip("oracle/_ref/libbpomp_ref.so with the reference planner is not built here")
    from batching_pomp_b200 import workloads as wl
    case = refcases.cases()[name]
    roadmap = wl.halton(case["N"], case["d"])
    fn = tempfile.mktemp(suffix=".graphml")
    orc.write_graphml(fn, roadmap, case["edges"])
    try:
        rp = orc.RefPlanner(case["d"], case["fraction"])
        assert rp.segment_length == wl.segment_length(case["d"], case["fraction"])
        rec = refcases.drive(rp, case, lambda p: p.set_param("roadmap_filename", fn))
    finally:
        os.unlink(fn)
    assert refcases.same(_golden(name, case["d"]), rec) is None


def test_reference_planner_errors(orc):
    """The exceptions of setup() / setProblemDefinition() (BP.cpp:681-688, 745-747, 777-786; Selector.hpp:74)."""
    if not orc.ref_planner_available():
        pytest.skip("reference planner not built here")
    from batching_pomp_b200 import workloads as wl
    fn = tempfile.mktemp(suffix=".graphml")
    orc.write_graphml(fn, wl.halton(50, 2))
    try:
        rp = orc.RefPlanner(2, 0.01)
        rp.set_world_boxes(np.array([[0.0, 0.0]]), np.array([[0.2, 0.2]]))
        rp.set_param("batching_type", "edge")
        rp.set_param("roadmap_filename", fn)
        with pytest.raises(RuntimeError, match="Graph type needed"):
            rp.setup()
        rp.set_param("graph_type", "halton")
        rp.set_param("batching_type", "bogus")
        with pytest.raises(RuntimeError, match="Invalid batching type"):
            rp.setup()
        rp.set_param("batching_type", "vertex")
        rp.set_param("selector_type", "bogus")
        with pytest.raises(RuntimeError, match="Invalid selector type"):
            rp.setup()
        with pytest.raises(RuntimeError, match="Start configuration is in collision"):
            rp.set_problem([0.1, 0.1], [0.9, 0.9])
        rq = orc.RefPlanner(2, 0.01)
        rq.set_world_boxes(np.array([[0.0, 0.0]]), np.array([[0.2, 0.2]]))
        with pytest.raises(RuntimeError, match="Goal configuration is in collision"):
            rq.set_problem([0.9, 0.9], [0.1, 0.1])
        rr = orc.RefPlanner(2, 0.01)
        rr.set_param("batching_type", "vertex")
        with pytest.raises(RuntimeError, match="Roadmap name must be set"):
            rr.setup()
    finally:
        os.unlink(fn)


def test_knn_model_matches_reference(orc):
    """KNNModel::estimate (KNNModel.hpp:101-134): the oracle against the reference's recorded outputs (and against
    the reference build live when present), 1e-12 relative as north_star states for belief measures."""
    z = np.load(os.path.join(G, "knn_ref.npz"), allow_pickle=False)
    for tag in "abcde":
        pts, vals, q, K, want = z["pts_" + tag], z["vals_" + tag], z["q_" + tag], int(z["K_" + tag]), z["est_" + tag]
        got = orc.belief_estimate(pts, vals, q, K=K)
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)
        if orc.ref_planner_available():
            np.testing.assert_array_equal(orc.ref_knn_estimate(pts, vals, q, K=K), want)
    # the empty model returns the prior (KNNModel.hpp:104-106)
    if orc.ref_planner_available():
        assert np.all(orc.ref_knn_estimate(np.zeros((0, 3)), np.zeros(0), np.ones((4, 3)) * 0.5) == 0.5)
    assert np.all(orc.belief_estimate(np.zeros((0, 3)), np.zeros(0), np.ones((4, 3)) * 0.5) == 0.5)


def test_selector_matches_reference(orc):
    """Selector::selectEdges (Selector.hpp:95-146) run from the reference's header: normal, alternate, failfast
    (ascending probFree), maxinf (ascending |probFree - 0.5|); equal keys in edge creation order."""
    if not orc.ref_planner_available():
        pytest.skip("reference build not present")
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 40):
        p = rng.random(n)
        p[rng.integers(0, n, n // 2)] = 0.5 ** rng.integers(0, 4, n // 2)  # many exact ties, as in a first batch
        idx = np.arange(n)
        assert orc.ref_selector("normal", p).tolist() == idx.tolist()
        assert orc.ref_selector("alternate", p).tolist() == idx[::-1].tolist()
        assert orc.ref_selector("failfast", p).tolist() == np.lexsort((idx, p)).tolist()
        assert orc.ref_selector("maxinf", p).tolist() == np.lexsort((idx, np.abs(p - 0.5))).tolist()
    with pytest.raises(ValueError):
        orc.ref_selector("bogus", np.zeros(3))


# ------------------------------------------------------------------------------------------------------------
# GPU: the product against the reference's records
# ------------------------------------------------------------------------------------------------------------
class _GpuPlannerAdapter:
    """The GPU host planner behind the protocol of refcases.drive (ptc as an iteration budget)."""

    def __init__(self, case):
        from batching_pomp_b200.planner import Planner
        w = refcases.world_of(case["world"])
        kw = {"image": w[1]} if w[0] == "image" else {"boxes": (w[1], w[2])}
        self.p = Planner(case["d"], segment_fraction=case["fraction"], **kw)
        self.remaining = -1

    def set_world_image(self, pix):
        pass

    def set_world_boxes(self, lo, hi):
        pass

    def set_param(self, k, v):
        self.p.set_param(k, v)

    def setup(self):
        self.p.setup()

    def set_problem(self, s, g):
        self.p.set_problem(s, g)

    def set_ptc_budget(self, n):
        self.remaining = n

    def solve(self):
        before = self.p.counters()["searches"]
        st = self.p.solve(max_iterations=self.remaining)
        self.remaining -= self.p.counters()["searches"] - before
        return st

    def counters(self):
        return self.p.counters()

    def path_states(self):
        return self.p.path_states()

    best_cost = property(lambda self: self.p.best_cost)
    alpha = property(lambda self: self.p.alpha)
    radius = property(lambda self: self.p.radius)
    num_batches = property(lambda self: self.p.num_batches)
    exhausted = property(lambda self: self.p.exhausted)
    num_vertices = property(lambda self: self.p.num_vertices)
    num_edges = property(lambda self: self.p.num_edges)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_planner_reproduces_reference_runs(name):
    from batching_pomp_b200 import workloads as wl
    case = refcases.cases()[name]
    roadmap = wl.halton(case["N"], case["d"])
    gp = _GpuPlannerAdapter(case)
    rec = refcases.drive(gp, case, lambda p: p.p.set_roadmap(roadmap, case["edges"]))
    assert refcases.same(_golden(name, case["d"]), rec) is None
    assert gp.p.gpu_launches > 0


@pytest.mark.gpu
def test_gpu_belief_estimates_match_reference_knn_model():
    from batching_pomp_b200 import capi
    z = np.load(os.path.join(G, "knn_ref.npz"), allow_pickle=False)
    for tag in "abcde":
        pts, vals, q, K, want = z["pts_" + tag], z["vals_" + tag], z["q_" + tag], int(z["K_" + tag]), z["est_" + tag]
        ctx = capi.Context(pts.shape[1], 0.01)
        ctx.belief_config(K, 0.5, 0.25)
        ctx.belief_append(pts, vals)
        np.testing.assert_allclose(ctx.belief_estimate(q), want, rtol=1e-12, atol=0)
        ctx.close()


@pytest.mark.parametrize("d,nb,fraction,max_len", [(7, 500, 0.01, None), (2, 40, 0.002, 0.4), (3, 200, 0.01, 1.2)])
def test_oracle_edge_operator_equals_reference_code(orc, d, nb, fraction, max_len):
    """initializeEdgePoints + checkAndSetEdgeBlocked of the reference itself (src/BatchingPOMP.cpp:435-461, 496-544,
    driven on arbitrary edges through oracle/ref_planner_shim.cpp: ref_edges_check_boxes) against the oracle's edge
    operator: validity, first invalid slot and nStates bit for bit — short edges (2 states), long edges (hundreds
    of states), zero-length edges, both orientations."""
    if not orc.ref_edges_available():
        pytest.skip("oracle/_ref/libbpomp_ref.so with ref_edges_check_boxes is not built here")
    from batching_pomp_b200 import workloads as wl
    L = wl.segment_length(d, fraction)
    lo, hi = wl.box_world(d, nb, seed=3)
    frm, to = wl.random_edges(30000, d, seed=4, max_len=max_len or wl.halton_radius(10 ** 6, d))
    frm[:10] = to[:10]                       # zero length
    frm[10:5000], to[10:5000] = to[10:5000].copy(), frm[10:5000].copy()  # other orientation
    rv, rf, rn = orc.ref_edges_check_boxes(frm, to, fraction, lo, hi, nthreads=orc.max_threads())
    ov, of, on = orc.edges_check_boxes(frm, to, L, lo, hi, nthreads=orc.max_threads())
    np.testing.assert_array_equal(on, rn)
    np.testing.assert_array_equal(ov, rv)
    np.testing.assert_array_equal(of, rf)
    assert 0.02 < (rv == 0).mean() < 0.98 and rn.max() > (20 if max_len else 8)


@pytest.mark.gpu
def test_gpu_edge_operator_equals_reference_code(orc):
    """The CUDA edge kernels against the reference's own edge code (not the oracle): config-4 style edges."""
    if not orc.ref_edges_available():
        pytest.skip("oracle/_ref/libbpomp_ref.so with ref_edges_check_boxes is not built here")
    from batching_pomp_b200 import capi, workloads as wl
    d = 7
    L = wl.segment_length(d)
    lo, hi = wl.box_world(d, 500, seed=3)
    frm, to = wl.random_edges(120000, d, seed=9, max_len=wl.halton_radius(10 ** 6, d))
    frm[:2000], to[:2000] = wl.random_edges(2000, d, seed=10, max_len=2.0)  # long edges: the general kernel's share
    ctx = capi.Context(d, L)
    ctx.set_bounds(0.0, 1.0)
    ctx.set_world_boxes(lo, hi)
    gv, gf, gn = ctx.edges_check(frm, to)
    rv, rf, rn = orc.ref_edges_check_boxes(frm, to, 0.01, lo, hi, nthreads=orc.max_threads())
    np.testing.assert_array_equal(gn, rn)
    np.testing.assert_array_equal(gv, rv)
    np.testing.assert_array_equal(gf, rf)
    ctx.close()
