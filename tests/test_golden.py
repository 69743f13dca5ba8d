"""Committed fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle with the
seeds recorded there).  CPU: the oracle still reproduces them.  GPU (-m gpu): the CUDA path reproduces them."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def test_oracle_reproduces_golden_edges(orc):
    g = _load("edges_boxes_7d.npz")
    v, f, n = orc.edges_check_boxes(g["frm"], g["to"], float(g["L"]), g["lo"], g["hi"])
    assert np.array_equal(v, g["valid"]) and np.array_equal(f, g["first"]) and np.array_equal(n, g["nstates"])
    g = _load("edges_image_2d.npz")
    v, f, n = orc.edges_check_image(g["frm"], g["to"], float(g["L"]), g["pix"])
    assert np.array_equal(v, g["valid"]) and np.array_equal(f, g["first"]) and np.array_equal(n, g["nstates"])
    assert 0 < int((g["valid"] == 0).sum()) < g["valid"].size


def test_oracle_reproduces_golden_neighbors_and_belief(orc):
    g = _load("neighbors_3d.npz")
    rp, col, dist = orc.neighbors_radius(g["verts"], g["query"], float(g["r"]), active=g["active"])
    assert np.array_equal(rp, g["row_ptr"]) and np.array_equal(col, g["col"]) and np.array_equal(dist, g["dist"])
    b = _load("belief_7d.npz")
    assert np.array_equal(orc.belief_estimate(b["pts"], b["vals"], b["queries"]), b["estimate"])
    pf, nl = orc.belief_edge_pfree(b["frm"], b["to"], float(b["L"]), b["pts"], b["vals"])
    assert np.array_equal(pf, b["pfree"]) and np.array_equal(nl, b["nlookups"])


def _replay_planner(solve, p, g):
    statuses, costs, paths, counters = [], [], [], []
    for _ in range(len(g["status"])):
        st = solve()
        statuses.append(st)
        costs.append(p.best_cost)
        paths.append(p.path())
        c = p.counters()
        counters.append([c["edge_checks"], c["coll_checks"], c["searches"], c["lookups"]])
        if st in ("EXACT_SOLUTION", "TIMEOUT"):
            break
    assert statuses == [str(s) for s in g["status"]]
    assert np.array_equal(np.array(costs), g["cost"])
    assert [len(x) for x in paths] == g["path_len"].tolist()
    assert np.array_equal(np.concatenate(paths), g["path_cat"])
    assert np.array_equal(np.array(counters), g["counters"])


def test_oracle_planner_reproduces_golden(orc):
    from batching_pomp_b200 import workloads
    g = _load("planner_hybrid_2d.npz")
    op = orc.OraclePlanner(2, workloads.segment_length(2, float(g["fraction"])), workloads.max_extent(2), 1.0)
    op.set_world_image(g["pix"])
    op.set_roadmap(g["roadmap"])
    op.set_param("batching_type", "hybrid")
    op.set_param("graph_type", "halton")
    op.set_problem([0.1, 0.1], [0.9, 0.9])
    op.setup()
    op.set_ptc_budget(int(g["ptc_budget"]))
    _replay_planner(op.solve, op, g)
    assert len(g["path_cat"]) > 0


@pytest.mark.gpu
def test_gpu_reproduces_golden():
    from batching_pomp_b200 import capi
    from batching_pomp_b200.planner import Planner
    g = _load("edges_boxes_7d.npz")
    ctx = capi.Context(7, float(g["L"]))
    ctx.set_world_boxes(g["lo"], g["hi"])
    v, f, n = ctx.edges_check(g["frm"], g["to"])
    assert np.array_equal(v, g["valid"]) and np.array_equal(f, g["first"]) and np.array_equal(n, g["nstates"])
    b = _load("belief_7d.npz")
    ctx.belief_append(b["pts"], b["vals"])
    assert np.array_equal(ctx.belief_estimate(b["queries"]), b["estimate"])
    pf, nl = ctx.belief_edge_pfree(b["frm"], b["to"])
    assert np.array_equal(pf, b["pfree"]) and np.array_equal(nl, b["nlookups"])
    ctx.close()
    g = _load("edges_image_2d.npz")
    ctx = capi.Context(2, float(g["L"]))
    ctx.set_world_image(g["pix"])
    v, f, n = ctx.edges_check(g["frm"], g["to"])
    assert np.array_equal(v, g["valid"]) and np.array_equal(f, g["first"]) and np.array_equal(n, g["nstates"])
    ctx.close()
    g = _load("neighbors_3d.npz")
    ctx = capi.Context(3, 0.01)
    ctx.vertices_append(g["verts"])
    ctx.vertices_set_active(np.nonzero(g["active"] == 0)[0].astype(np.uint32), 0)
    rp, col, dist = ctx.neighbors_radius(g["query"], float(g["r"]))
    assert np.array_equal(rp, g["row_ptr"]) and np.array_equal(col, g["col"]) and np.array_equal(dist, g["dist"])
    ctx.close()
    g = _load("planner_hybrid_2d.npz")
    p = Planner(2, segment_fraction=float(g["fraction"]), image=g["pix"])
    p.set_roadmap(g["roadmap"])
    p.set_param("batching_type", "hybrid")
    p.set_param("graph_type", "halton")
    p.set_problem([0.1, 0.1], [0.9, 0.9])
    p.setup()
    remaining = [int(g["ptc_budget"])]

    def solve():
        before = p.counters()["searches"]
        st = p.solve(max_iterations=remaining[0])
        remaining[0] -= p.counters()["searches"] - before
        return st
    _replay_planner(solve, p, g)
