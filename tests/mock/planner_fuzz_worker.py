#!/usr/bin/env python
"""Randomised differential test of the host planner against the oracle planner (device = test double).
    BPOMP_HOST_LIB=<tmp>/libbpomp_host.so python tests/mock/planner_fuzz_worker.py <first_seed> <count>
Each seed draws the dimension, the roadmap size, the world, the batching / graph / selector types, the
increment, the prune threshold and the query.  TEST INFRASTRUCTURE ONLY."""
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


def one(seed):
    rng = np.random.default_rng(seed)
    d = int(rng.choice([2, 2, 3, 5, 7]))
    N = int(rng.integers(120, 700))
    batching = str(rng.choice(["vertex", "edge", "hybrid"]))
    params = {"batching_type": batching, "increment": float(rng.choice([0.1, 0.2, 0.34, 0.5, 1.0])),
              "prune_threshold": float(rng.choice([1.0, 1.05, 1.2, 2.0]))}
    if batching != "vertex":
        params["graph_type"] = str(rng.choice(["halton", "rgg"]))
    if rng.random() < 0.6:
        params["selector_type"] = str(rng.choice(["normal", "alternate", "failfast", "maxinf"]))
    if d == 2 and rng.random() < 0.5:
        world = ("image", wl.image_world(int(rng.choice([64, 128, 256])), seed=int(rng.integers(1, 50))))
        fraction = float(rng.choice([2e-3, 5e-3, 1e-2]))
    else:
        lo, hi = wl.box_world(d, int(rng.integers(5, 60 if d < 5 else 400)), seed=int(rng.integers(1, 50)))
        world = ("boxes", lo, hi)
        fraction = float(rng.choice([5e-3, 1e-2, 2e-2]))
    start = np.full(d, 0.1)
    goal = np.full(d, 0.9)
    single = d <= 3 and rng.random() < 0.2
    edges = graphml = None
    tmp = None
    if single:  # the roadmap file carries the edges (SingleBatching.hpp:55-91); start/goal linked within a radius
        import tempfile
        roadmap = wl.halton(N, d)
        r = float(rng.choice([1.5, 2.0, 3.0])) * wl.halton_radius(N, d)
        diff = roadmap[:, None, :] - roadmap[None, :, :]
        iu, ju = np.nonzero(np.triu(np.sqrt((diff ** 2).sum(-1)) <= r, 1))
        edges = np.stack([iu, ju], 1).astype(np.uint32)
        params = {"batching_type": "single", "start_goal_radius": r, "increment": params["increment"]}
        tmp = tempfile.TemporaryDirectory()
        graphml = os.path.join(tmp.name, "roadmap.graphml")
    op, gp = _make(orc, wl, d, N, world, params, fraction, None, None, edges=edges, graphml=graphml)
    try:
        nsol = _run_both(op, gp, start, goal, single, max_calls=25, budget=int(rng.integers(20, 90)))
    except AssertionError as e:
        raise AssertionError("seed %d (%s, d=%d, N=%d): %s" % (seed, params, d, N, e))
    finally:
        if tmp is not None:
            tmp.cleanup()
    return nsol


if __name__ == "__main__":
    first, count = int(sys.argv[1]), int(sys.argv[2])
    tot = 0
    for s in range(first, first + count):
        tot += one(s)
    print("mock fuzz ok: seeds %d..%d, %d solutions" % (first, first + count - 1, tot))
