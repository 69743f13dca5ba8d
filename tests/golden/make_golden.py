#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle (seeds recorded below).  The reference ships no
golden vectors and cannot be built here, so these fixtures pin the ORACLE (and through it
the CUDA kernels) against accidental drift; they are not outputs of the reference itself.

    python tests/golden/make_golden.py        # rewrites the fixtures
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from batching_pomp_b200 import workloads  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    # --- edge evaluation, 7-DoF boxes (config 3/4 shape): seeds 3 (world), 104 (edges)
    d = 7
    L = workloads.segment_length(d)
    lo, hi = workloads.box_world(d, 500, seed=3)
    frm, to = workloads.random_edges(1500, d, seed=104, max_len=0.8)
    v, f, n = oracle.edges_check_boxes(frm, to, L, lo, hi)
    np.savez_compressed(os.path.join(HERE, "edges_boxes_7d.npz"), L=L, lo=lo, hi=hi, frm=frm, to=to, valid=v,
                        first=f, nstates=n)
    # --- edge evaluation, 2-D image (config 1 shape): seeds 1 (image), 105 (edges)
    pix = workloads.image_world(128, seed=1)
    L2 = workloads.segment_length(2, 1e-3)
    frm2, to2 = workloads.random_edges(1500, 2, seed=105, max_len=0.08)
    v2, f2, n2 = oracle.edges_check_image(frm2, to2, L2, pix)
    np.savez_compressed(os.path.join(HERE, "edges_image_2d.npz"), L=L2, pix=pix, frm=frm2, to=to2, valid=v2,
                        first=f2, nstates=n2)
    # --- neighbour rows: Halton N=1200, 3-D, haltonRadius; every 5th vertex inactive
    N = 1200
    verts = workloads.halton(N, 3)
    active = np.ones(N, dtype=np.uint8)
    active[::5] = 0
    q = np.arange(0, N, 7, dtype=np.uint32)
    r = workloads.halton_radius(N, 3)
    rp, col, dist = oracle.neighbors_radius(verts, q, r, active=active)
    np.savez_compressed(os.path.join(HERE, "neighbors_3d.npz"), verts=verts, active=active, query=q, r=r,
                        row_ptr=rp, col=col, dist=dist)
    # --- belief: 600 points in 7-D, 300 queries, k = 15; edge free-probabilities of 200 edges
    pts = workloads.uniform(131, (600, d))
    vals = (workloads.uniform(132, (600,)) < 0.3).astype(np.float64)
    qs = workloads.uniform(133, (300, d))
    est = oracle.belief_estimate(pts, vals, qs)
    pf, nl = oracle.belief_edge_pfree(frm[:200], to[:200], L, pts, vals)
    np.savez_compressed(os.path.join(HERE, "belief_7d.npz"), pts=pts, vals=vals, queries=qs, estimate=est,
                        frm=frm[:200], to=to[:200], L=L, pfree=pf, nlookups=nl)
    # --- planner: 2-D image, Halton N=1000, hybrid batching, first solutions
    roadmap = workloads.halton(1000, 2)
    op = oracle.OraclePlanner(2, workloads.segment_length(2, 2e-3), workloads.max_extent(2), 1.0)
    op.set_world_image(pix)
    op.set_roadmap(roadmap)
    for k, val in (("batching_type", "hybrid"), ("graph_type", "halton")):
        op.set_param(k, val)
    op.set_problem([0.1, 0.1], [0.9, 0.9])
    op.setup()
    log_status, log_cost, paths, counters = [], [], [], []
    op.set_ptc_budget(60)  # bounded: at most 60 searches in total (deterministic termination condition)
    for _ in range(40):
        st = op.solve()
        log_status.append(st)
        log_cost.append(op.best_cost)
        paths.append(op.path())
        c = op.counters()
        counters.append([c["edge_checks"], c["coll_checks"], c["searches"], c["lookups"]])
        if st in ("EXACT_SOLUTION", "TIMEOUT"):
            break
    np.savez_compressed(os.path.join(HERE, "planner_hybrid_2d.npz"), pix=pix, roadmap=roadmap, fraction=2e-3,
                        ptc_budget=60,
                        status=np.array(log_status), cost=np.array(log_cost),
                        path_len=np.array([len(p) for p in paths]), path_cat=np.concatenate(paths),
                        counters=np.array(counters))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
