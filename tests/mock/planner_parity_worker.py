#!/usr/bin/env python
"""Subprocess body of tests/test_host_mock_cpu.py: the C++ host planner, loaded from BPOMP_HOST_LIB (a copy
of libbpomp_host.so next to the test double of libbpomp_cuda.so, tests/mock/mock_cuda.cpp), against the
literal planner restatement of the oracle, solve() call by solve() call — the same comparison
tests/test_gpu_planner.py makes on the GPU (its helpers are reused).  TEST INFRASTRUCTURE ONLY."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

assert os.environ.get("BPOMP_HOST_LIB"), "run through tests/test_host_mock_cpu.py"

from batching_pomp_b200 import workloads as wl  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from test_gpu_planner import _make, _run_both  # noqa: E402


def case_hybrid_image():
    pix = wl.image_world(256, seed=1)
    op, gp = _make(orc, wl, 2, 1500, ("image", pix), {"batching_type": "hybrid", "graph_type": "halton",
                                                       "increment": 0.2, "prune_threshold": 1.05}, 2e-3, None, None)
    nsol = _run_both(op, gp, [0.1, 0.1], [0.9, 0.9], False, budget=200)
    assert nsol >= 1 and gp.gpu_launches > 0
    return nsol


def case_edge_selectors():
    pix = wl.image_world(256, seed=2)
    total = 0
    for sel in ("normal", "alternate", "failfast", "maxinf"):
        op, gp = _make(orc, wl, 2, 800, ("image", pix), {"batching_type": "edge", "graph_type": "halton",
                                                          "selector_type": sel}, 2e-3, None, None)
        n = _run_both(op, gp, [0.1, 0.1], [0.9, 0.9], False, budget=120)
        assert n >= 1
        total += n
    return total


def case_vertex_7d():
    d = 7
    lo, hi = wl.box_world(d, 500, seed=3)
    op, gp = _make(orc, wl, d, 600, ("boxes", lo, hi), {"batching_type": "vertex"}, 0.01, None, None)
    nsol = _run_both(op, gp, np.full(d, 0.1), np.full(d, 0.9), False, max_calls=20, budget=80)
    assert nsol >= 1
    return nsol


def case_rgg_edge_7d():
    d = 7
    lo, hi = wl.box_world(d, 500, seed=3)
    op, gp = _make(orc, wl, d, 700, ("boxes", lo, hi), {"batching_type": "edge", "graph_type": "rgg",
                                                        "increment": 0.5}, 0.01, None, None)
    nsol = _run_both(op, gp, np.full(d, 0.1), np.full(d, 0.9), False, max_calls=20, budget=80)
    assert nsol >= 1
    return nsol


def case_single_graphml():
    import tempfile
    d, N = 2, 600
    pix = wl.image_world(256, seed=1)
    roadmap = wl.halton(N, d)
    r = 2.0 * wl.halton_radius(N, d)
    diff = roadmap[:, None, :] - roadmap[None, :, :]
    dist = np.sqrt((diff ** 2).sum(-1))
    iu, ju = np.nonzero(np.triu(dist <= r, 1))
    edges = np.stack([iu, ju], 1).astype(np.uint32)
    with tempfile.TemporaryDirectory() as td:
        op, gp = _make(orc, wl, d, N, ("image", pix), {"batching_type": "single", "start_goal_radius": r}, 2e-3, None,
                       None, edges=edges, graphml=os.path.join(td, "roadmap.graphml"))
        nsol = _run_both(op, gp, [0.1, 0.1], [0.9, 0.9], True, budget=150)
    assert nsol >= 1
    return nsol


CASES = {f.__name__[5:]: f for f in (case_hybrid_image, case_edge_selectors, case_vertex_7d, case_rgg_edge_7d,
                                     case_single_graphml)}

if __name__ == "__main__":
    name = sys.argv[1]
    n = CASES[name]()
    print("mock parity ok: %s, %d solutions" % (name, n))
