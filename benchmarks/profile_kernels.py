#!/usr/bin/env python
"""Small driver for ncu captures of the kernels other than the headline one:
    python benchmarks/profile_kernels.py neighbors2d | neighbors7d | belief | image
Each mode runs its kernel three times on a bench-sized input (ncu: -k regex:<kernel> -s 1 -c 1)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from batching_pomp_b200 import capi, workloads  # noqa: E402

mode = sys.argv[1]
if mode == "neighbors2d":
    N = 10 ** 6
    ctx = capi.Context(2, workloads.segment_length(2, 1e-4))
    ctx.vertices_append(workloads.halton(N, 2))
    for _ in range(3):
        print(ctx.neighbors_all(workloads.halton_radius(N, 2), fetch=False))
elif mode == "neighbors7d":
    N = 10 ** 6
    ctx = capi.Context(7, workloads.segment_length(7))
    ctx.vertices_append(workloads.halton(N, 7))
    q = (np.arange(2048, dtype=np.uint64) * (N // 2048)).astype(np.uint32)
    for _ in range(3):
        print(ctx.neighbors_radius(q, workloads.halton_radius(N, 7))[0][-1])
elif mode == "belief":
    d, M, nq = 7, 200000, 100000
    ctx = capi.Context(d, workloads.segment_length(d))
    ctx.belief_append(workloads.uniform(31, (M, d)), (workloads.uniform(32, (M,)) < 0.3).astype(np.float64))
    q = workloads.uniform(33, (nq, d))
    for _ in range(3):
        print(ctx.belief_estimate(q)[:2])
elif mode == "image":
    d = 2
    L = workloads.segment_length(d, 1e-4)
    ctx = capi.Context(d, L)
    ctx.set_world_image(workloads.image_world(4096, seed=2))
    frm, to = workloads.random_edges(5 * 10 ** 6, d, seed=4, max_len=workloads.halton_radius(10 ** 6, 2))
    for _ in range(3):
        print(int((ctx.edges_check(frm[:1500000], to[:1500000])[0] == 0).sum()))
elif mode == "indexed":  # config 3 roadmap rows as indexed edges (bench.py's indexed leg)
    d, N, nq = 7, 10 ** 6, 4096
    L = workloads.segment_length(d)
    lo, hi = workloads.box_world(d, 500, seed=3)
    verts = workloads.halton(N, d)
    ctx = capi.Context(d, L)
    ctx.set_bounds(0.0, 1.0)
    ctx.set_world_boxes(lo, hi)
    ctx.vertices_append(verts)
    q = (np.arange(nq, dtype=np.uint64) * (N // nq)).astype(np.uint32)
    rp, col, _ = ctx.neighbors_radius(q, workloads.halton_radius(N, d))
    fi = np.repeat(q, np.diff(rp.astype(np.int64)))
    for _ in range(3):
        print(int((ctx.edges_check_indexed(fi, col)[0] == 0).sum()))
elif mode == "image_indexed":  # config 2 roadmap edges (bench.py's image leg)
    d, N = 2, 10 ** 6
    verts = workloads.halton(N, d)
    ctx = capi.Context(d, workloads.segment_length(d, 1e-4))
    ctx.set_world_image(workloads.image_world(4096, seed=2))
    ctx.vertices_append(verts)
    rp, col, _ = ctx.neighbors_all(workloads.halton_radius(N, d))
    rows = np.repeat(np.arange(N, dtype=np.uint32), np.diff(rp.astype(np.int64)))
    keep = rows < col
    fi, ti = np.ascontiguousarray(rows[keep]), np.ascontiguousarray(col[keep])
    for _ in range(3):
        print(int((ctx.edges_check_indexed(fi, ti)[0] == 0).sum()))
elif mode == "pfree_fused":  # the planner's per-expansion free-probability call (one launch)
    d, M = 2, 35000
    L = workloads.segment_length(d, 1e-3)
    ctx = capi.Context(d, L)
    ctx.belief_append(workloads.uniform(31, (M, d)), (workloads.uniform(32, (M,)) < 0.3).astype(np.float64))
    frm, to = workloads.random_edges(400, d, seed=4, max_len=workloads.halton_radius(10 ** 4, 2))
    for _ in range(3):
        print(ctx.belief_edge_pfree(frm, to)[0][:2])
