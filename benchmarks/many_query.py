#!/usr/bin/env python
"""BASELINE config 5 — many-query sweep: start/goal pairs on a shared 7-DoF roadmap, sharded across the
GPUs of one box (one process per GPU, planner instances are independent: no data-path collective, only
the final gather of the per-query results).

  python benchmarks/many_query.py [--pairs 4096] [--vertices 1000000] [--solutions 1] [--check 8]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/many_query.py ...

Each query runs batching_pomp's anytime loop (vertex batching) until `--solutions` improved paths were
returned.  `--check K`: rank 0 re-plans its first K queries with the CPU oracle planner and requires
identical vertex sequences and costs."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from batching_pomp_b200 import capi, sharding, workloads  # noqa: E402
from batching_pomp_b200.planner import Planner  # noqa: E402


def plan(p, start, goal, solutions, max_calls=200):
    p.set_problem(start, goal)
    p.setup()
    found = 0
    last = None
    for _ in range(max_calls):
        st = p.solve()
        if len(p.path()) and p.best_cost != last:
            last = p.best_cost
            found += 1
        if found >= solutions or st in ("EXACT_SOLUTION", "TIMEOUT"):
            break
    return found, p.best_cost, [int(v) for v in p.path()]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4096)
    ap.add_argument("--vertices", type=int, default=10 ** 6)
    ap.add_argument("--solutions", type=int, default=1)
    ap.add_argument("--check", type=int, default=8)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")  # results are small host objects; the data path has no collective
    d = 7
    lo, hi = workloads.box_world(d, 500, seed=3, keep_free=())
    roadmap = workloads.halton(a.vertices, d)
    # start/goal pairs ~U[0,1]^7, rejected until valid (seed 5)
    ctx = capi.Context(d, workloads.segment_length(d), device=local)
    ctx.set_world_boxes(lo, hi)
    cand = workloads.uniform(5, (4 * a.pairs + 64, d))
    cand = cand[ctx.states_valid(cand) == 1][:2 * a.pairs].reshape(-1, 2, d)
    ctx.close()
    assert cand.shape[0] == a.pairs, "not enough valid start/goal candidates"
    b = sharding.shard_bounds(a.pairs, world)
    mine = range(b[rank], b[rank + 1])
    t0 = time.perf_counter()
    out = []
    t_create = 0.0
    for qi in mine:
        c0 = time.perf_counter()
        p = Planner(d, boxes=(lo, hi), device=local)
        t_create += time.perf_counter() - c0
        p.set_roadmap_view(roadmap)  # one loaded roadmap serves every planner instance of this rank
        p.set_param("batching_type", "vertex")
        q0 = time.perf_counter()
        found, cost, path = plan(p, cand[qi, 0], cand[qi, 1], a.solutions)
        out.append((qi, found, cost, path, time.perf_counter() - q0, p.counters()))
        p.close()
    elapsed = time.perf_counter() - t0
    checked = None
    if rank == 0 and a.check > 0:
        from oracle import oracle
        checked = 0
        for qi, found, cost, path, _, counters in out[:a.check]:
            op = oracle.OraclePlanner(d, workloads.segment_length(d), workloads.max_extent(d), 1.0)
            op.set_world_boxes(lo, hi)
            op.set_roadmap(roadmap)
            op.set_param("batching_type", "vertex")
            f2, c2, p2 = plan(op, cand[qi, 0], cand[qi, 1], a.solutions)
            assert (f2, c2, p2) == (found, cost, path) and op.counters() == counters, "query %d differs from the oracle" % qi
            checked += 1
    if world > 1:
        import torch.distributed as dist
        gathered = [None] * world
        dist.all_gather_object(gathered, (elapsed, out))
    else:
        gathered = [(elapsed, out)]
    if rank == 0:
        print("rank 0: %.3f s creating %d planner instances" % (t_create, len(out)), file=sys.stderr)
    if rank == 0:
        allq = [q for _, o in gathered for q in o]
        tmax = max(e for e, _ in gathered)
        solved = sum(1 for q in allq if q[1] > 0)
        lat = sorted(q[4] for q in allq)
        print(json.dumps({"section": "config5_many_query", "pairs": a.pairs, "n_gpus": world, "roadmap_vertices": a.vertices,
                          "solutions_per_query": a.solutions, "solved": solved, "seconds_max_over_ranks": tmax,
                          "queries_per_s": len(allq) / tmax, "latency_median_s": lat[len(lat) // 2],
                          "latency_p95_s": lat[int(len(lat) * 0.95)], "mean_path_vertices":
                          float(np.mean([len(q[3]) for q in allq if q[1] > 0])) if solved else None,
                          "oracle_checked_queries": checked, "sharding": "contiguous query ranges per rank, world and "
                          "roadmap replicated, final all_gather_object of the results only"}))
        sys.stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
