"""Path parity: the host planner (C++ on flat arrays + CUDA through the C ABI) against the literal CPU
restatement of batching_pomp::BatchingPOMP in the oracle, solve() call by solve() call: same status,
same path vertex sequence, bit-equal best cost, same counters (mNumEdgeChecks, mNumCollChecks,
mNumSearches, mNumLookups), same radius / alpha schedule."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wl():
    from batching_pomp_b200 import workloads
    return workloads


def _make(orc, wl, d, N, world, params, fraction, start, goal, edges=None, graphml=None):
    from batching_pomp_b200.planner import Planner
    L = wl.segment_length(d, fraction)
    roadmap = wl.halton(N, d)
    op = orc.OraclePlanner(d, L, wl.max_extent(d), 1.0)
    if world[0] == "boxes":
        op.set_world_boxes(world[1], world[2])
        gp = Planner(d, segment_fraction=fraction, boxes=(world[1], world[2]))
    else:
        op.set_world_image(world[1])
        gp = Planner(d, segment_fraction=fraction, image=world[1])
    op.set_roadmap(roadmap, edges)
    if graphml:
        from batching_pomp_b200.planner import write_graphml
        write_graphml(graphml, roadmap, edges)
        gp.set_param("roadmap_filename", graphml)
    else:
        gp.set_roadmap(roadmap, edges)
    for k, v in params.items():
        op.set_param(k, v)
        gp.set_param(k, v)
    return op, gp


def _run_both(op, gp, start, goal, single, max_calls=60, budget=400):
    if single:
        op.setup(); gp.setup()
        op.set_problem(start, goal); gp.set_problem(start, goal)
    else:
        op.set_problem(start, goal); gp.set_problem(start, goal)
        op.setup(); gp.setup()
    op.set_ptc_budget(budget)
    remaining = budget
    nsol = 0
    for call in range(max_calls):
        so = op.solve()
        before = gp.counters()["searches"]
        sg = gp.solve(max_iterations=remaining)
        remaining -= gp.counters()["searches"] - before  # the ptc is evaluated once per search iteration (+ exits)
        assert sg == so, "status differs at call %d: %s vs %s" % (call, sg, so)
        assert gp.counters() == op.counters(), "counters differ at call %d" % call
        assert gp.best_cost == op.best_cost
        assert gp.alpha == op.alpha and gp.radius == op.radius
        assert gp.num_batches == op.num_batches and gp.exhausted == op.exhausted
        np.testing.assert_array_equal(gp.path(), op.path())
        np.testing.assert_array_equal(gp.path_states(), op.path_states())
        assert gp.num_vertices == op.num_vertices and gp.num_edges == op.num_edges
        if so == "APPROXIMATE_SOLUTION" and len(op.path()):
            nsol += 1
        if so in ("EXACT_SOLUTION", "TIMEOUT"):
            break
    return nsol


def test_planner_parity_config1_hybrid_image(orc, wl):
    """BASELINE config 1 scaled to N=3000: 2-D image world, hybrid batching, halton radius."""
    pix = wl.image_world(256, seed=1)
    op, gp = _make(orc, wl, 2, 3000, ("image", pix), {"batching_type": "hybrid", "graph_type": "halton",
                                                       "increment": 0.2, "prune_threshold": 1.05}, 2e-3,
                   None, None)
    nsol = _run_both(op, gp, [0.1, 0.1], [0.9, 0.9], False)
    assert nsol >= 1 and gp.best_cost < 3.0
    assert gp.gpu_launches > 0


def test_planner_parity_edge_batching_selector(orc, wl):
    """Config 2 shape: edge batching with selector_type=normal (stop at the first blocked edge)."""
    pix = wl.image_world(256, seed=2)
    for sel in ("normal", "alternate", "failfast", "maxinf"):
        op, gp = _make(orc, wl, 2, 1500, ("image", pix), {"batching_type": "edge", "graph_type": "halton",
                                                           "selector_type": sel}, 2e-3, None, None)
        nsol = _run_both(op, gp, [0.1, 0.1], [0.9, 0.9], False, budget=250)
        assert nsol >= 1


def test_planner_parity_vertex_batching_7dof(orc, wl):
    """Config 3 shape: 7-DoF boxes, vertex batching (complete graph over the admitted vertices)."""
    d = 7
    lo, hi = wl.box_world(d, 500, seed=3)
    op, gp = _make(orc, wl, d, 1200, ("boxes", lo, hi), {"batching_type": "vertex"}, 0.01, None, None)
    nsol = _run_both(op, gp, np.full(d, 0.1), np.full(d, 0.9), False, max_calls=25, budget=120)
    assert nsol >= 1


def test_planner_parity_rgg_radius_edge_batching_7dof(orc, wl):
    d = 7
    lo, hi = wl.box_world(d, 500, seed=3)
    op, gp = _make(orc, wl, d, 1500, ("boxes", lo, hi), {"batching_type": "edge", "graph_type": "rgg",
                                                         "increment": 0.5}, 0.01, None, None)
    nsol = _run_both(op, gp, np.full(d, 0.1), np.full(d, 0.9), False, max_calls=25, budget=120)
    assert nsol >= 1


def test_planner_parity_single_batching_graphml(orc, wl, tmp_path):
    """single batching: the roadmap file carries the edges; start/goal linked within start_goal_radius."""
    d, N = 2, 800
    pix = wl.image_world(256, seed=1)
    roadmap = wl.halton(N, d)
    r = 2.0 * wl.halton_radius(N, d)
    diff = roadmap[:, None, :] - roadmap[None, :, :]
    dist = np.sqrt((diff ** 2).sum(-1))
    iu, ju = np.nonzero(np.triu(dist <= r, 1))
    edges = np.stack([iu, ju], 1).astype(np.uint32)
    op, gp = _make(orc, wl, d, N, ("image", pix), {"batching_type": "single", "start_goal_radius": r}, 2e-3, None,
                   None, edges=edges, graphml=str(tmp_path / "roadmap.graphml"))
    nsol = _run_both(op, gp, [0.1, 0.1], [0.9, 0.9], True, budget=200)
    assert nsol >= 1


def test_planner_parity_config1_full_size(orc, wl):
    """BASELINE config 1 at its real size: 2-D 1024^2 occupancy image, 10^4-vertex Halton roadmap, hybrid
    batching, resolution 1e-3 — every solve() call compared with the oracle planner."""
    pix = wl.image_world(1024, seed=1)
    op, gp = _make(orc, wl, 2, 10 ** 4, ("image", pix), {"batching_type": "hybrid", "graph_type": "halton",
                                                          "increment": 0.2, "prune_threshold": 1.05}, 1e-3,
                   None, None)
    nsol = _run_both(op, gp, [0.1, 0.1], [0.9, 0.9], False, max_calls=40, budget=120)
    assert nsol >= 3 and gp.best_cost < 2.0


def test_planner_parity_config3_full_size(orc, wl):
    """BASELINE config 3 at its real size: 7-DoF, 500 boxes, 10^6-vertex Halton roadmap, vertex batching
    (admitted sets 100, 200, 400, ... and the complete graph over them, VertexBatching.hpp:94-138)."""
    d = 7
    lo, hi = wl.box_world(d, 500, seed=3)
    op, gp = _make(orc, wl, d, 10 ** 6, ("boxes", lo, hi), {"batching_type": "vertex"}, 0.01, None, None)
    nsol = _run_both(op, gp, np.full(d, 0.1), np.full(d, 0.9), False, max_calls=30, budget=200)
    assert nsol >= 2


def test_planner_errors_mirror_reference(wl):
    from batching_pomp_b200.planner import Planner, PlannerError
    lo, hi = wl.box_world(2, 10, seed=3)
    p = Planner(2, boxes=(lo, hi))
    p.set_roadmap(wl.halton(100, 2))
    p.set_param("batching_type", "edge")
    with pytest.raises(PlannerError) as ei:
        p.setup()  # BP.cpp:681-683
    assert "Graph type needed" in str(ei.value)
    p.set_param("graph_type", "halton")
    p.set_param("batching_type", "bogus")
    with pytest.raises(PlannerError) as ei:
        p.setup()  # BP.cpp:745-747
    assert "Invalid batching type" in str(ei.value)
    p.set_param("batching_type", "vertex")
    p.set_param("selector_type", "bogus")
    with pytest.raises(PlannerError) as ei:
        p.setup()  # Selector.hpp:74
    assert ei.value.code == -2
    q = Planner(2, boxes=(np.array([[0.0, 0.0]]), np.array([[0.2, 0.2]])))
    with pytest.raises(PlannerError) as ei:
        q.set_problem([0.1, 0.1], [0.9, 0.9])  # BP.cpp:777-779
    assert "Start configuration is in collision" in str(ei.value)
    with pytest.raises(PlannerError) as ei:
        q.set_problem([0.9, 0.9], [0.1, 0.1])
    assert "Goal configuration is in collision" in str(ei.value)
    r = Planner(2, boxes=(lo, hi))
    r.set_param("batching_type", "vertex")
    with pytest.raises(PlannerError) as ei:
        r.setup()  # BP.cpp:686-688
    assert "Roadmap name must be set" in str(ei.value)
