#!/usr/bin/env python
"""Secondary benchmarks: the BASELINE.json configs other than the headline (bench.py = config 4).

  python benchmarks/run_configs.py [--quick] [--out FILE(jsonl, appended)] [--only NAME[,NAME]]

Sections (each prints one JSON line and is also collected into --out):
  config2_neighbors      whole-roadmap densification, 2-D Halton N=10^6, r = haltonRadius(N) -> device CSR
  config2_edges_image    every roadmap edge of that CSR (u<v) against the 4096^2 occupancy image, indexed
  config3_neighbors_7d   r-disk rows of 2048 query vertices over a 7-D Halton N=10^6 roadmap
  config3_edges_indexed  10^7 roadmap edges (pairs from those rows) against the 500-box world, indexed
  belief_knn             kNN collision-measure queries: M belief points, nq queries (k=15)
  planner_config1        time to first / best path, 2-D image, N=10^4, hybrid batching: GPU planner vs CPU oracle
  planner_config3        7-DoF boxes, vertex batching (bounded number of solve() calls): GPU planner vs CPU oracle
  planner_full           configs 2 and 3 as planner runs on the full 10^6-vertex roadmaps (GPU planner only)
Times are wall clock around synchronous C-ABI calls unless marked device (CUDA events)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from batching_pomp_b200 import capi, workloads  # noqa: E402

PEAK = 6550.4
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
RESULTS = {}


OUT = [None]


def emit(name, d):
    d = dict(section=name, **d)
    RESULTS[name] = d
    line = json.dumps(d)
    print(line, flush=True)
    if OUT[0]:
        with open(OUT[0], "a") as f:
            f.write(line + "\n")


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), sum(ts) / len(ts)


def device_timed(ctx, fn, reps=5, warm=2):
    import torch
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    for _ in range(warm):
        fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    ctx.sync()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def config2(quick):
    import torch
    d = 2
    N = 200000 if quick else 10 ** 6
    size = 4096
    frac = 1e-4
    verts = workloads.halton(N, d)
    r = workloads.halton_radius(N, d)
    L = workloads.segment_length(d, frac)
    ctx = capi.Context(d, L)
    ctx.set_world_image(workloads.image_world(size, seed=2))
    ctx.vertices_append(verts)
    nnz = [0]

    def dens():
        nnz[0] = ctx.neighbors_all(r, fetch=False)
    best, mean = timed(dens)
    algo = N * d * 8 + nnz[0] * 12 + (N + 1) * 8
    emit("config2_neighbors", {"N": N, "dim": d, "radius": r, "nnz": nnz[0], "mean_degree": nnz[0] / N,
                               "seconds_best": best, "seconds_mean": mean, "rows_per_s": N / best,
                               "neighbours_per_s": nnz[0] / best, "algorithmic_bytes": algo,
                               "achieved_GBps": algo / best / 1e9, "hbm_peak_GBps": PEAK,
                               "frac_of_hbm_roofline": algo / best / 1e9 / PEAK,
                               "timing": "wall clock around bpomp_neighbors_all (grid build + count + scan + fill + sort)"})
    rp, col, dist = ctx.neighbors_all(r)
    rows = np.repeat(np.arange(N, dtype=np.uint32), np.diff(rp.astype(np.int64)))
    keep = rows < col
    fi, ti = rows[keep], col[keep]
    E = fi.size
    dev = torch.device("cuda:0")
    tf, tt = torch.from_numpy(fi.astype(np.int32)).to(dev), torch.from_numpy(ti.astype(np.int32)).to(dev)
    tv = torch.empty(E, dtype=torch.uint8, device=dev)
    tfi = torch.empty(E, dtype=torch.int32, device=dev)
    tn = torch.empty(E, dtype=torch.int32, device=dev)
    sec = device_timed(ctx, lambda: ctx.edges_check_indexed_dev(tf.data_ptr(), tt.data_ptr(), E, tv.data_ptr(),
                                                                tfi.data_ptr(), tn.data_ptr()))
    algo_e = E * (8 + 5) + N * d * 8
    emit("config2_edges_image", {"edges": int(E), "blocked_fraction": float((tv == 0).float().mean()),
                                 "mean_nstates": float(tn.float().mean()), "seconds_device": sec,
                                 "edges_per_s": E / sec, "algorithmic_bytes": algo_e,
                                 "achieved_GBps": algo_e / sec / 1e9, "hbm_peak_GBps": PEAK,
                                 "frac_of_hbm_roofline": algo_e / sec / 1e9 / PEAK,
                                 "timing": "CUDA events, indexed endpoints, vertex table + 2 MB bitmap L2-resident"})
    # CPU baseline on a sample
    from oracle import oracle
    m = min(E, 200000)
    pix = workloads.image_world(size, seed=2)
    t0 = time.perf_counter()
    oracle.edges_check_image(verts[fi[:m]], verts[ti[:m]], L, pix, nthreads=oracle.max_threads())
    cpu = m / (time.perf_counter() - t0)
    RESULTS["config2_edges_image"]["cpu_baseline_edges_per_s"] = cpu
    RESULTS["config2_edges_image"]["cpu_cores"] = oracle.max_threads()
    print(json.dumps({"section": "config2_edges_image.cpu", "edges_per_s": cpu, "cores": oracle.max_threads()}))
    ctx.close()


def config3(quick):
    import torch
    d = 7
    N = 200000 if quick else 10 ** 6
    verts = workloads.halton(N, d)
    r = workloads.halton_radius(N, d)
    L = workloads.segment_length(d)
    lo, hi = workloads.box_world(d, 500, seed=3)
    ctx = capi.Context(d, L)
    ctx.set_world_boxes(lo, hi)
    ctx.vertices_append(verts)
    nq = 512 if quick else 2048
    q = (np.arange(nq, dtype=np.uint64) * (N // nq)).astype(np.uint32)
    out = [None]

    def rows():
        out[0] = ctx.neighbors_radius(q, r)

    def rows_dev():
        ctx.neighbors_radius(q, r, fetch=False)
    best_dev, _ = timed(rows_dev, reps=3)
    best, mean = timed(rows, reps=2)
    rp, col, dist = out[0]
    nnz = int(rp[-1])
    algo = N * d * 8 + nnz * 12 + (nq + 1) * 8
    emit("config3_neighbors_7d", {"N": N, "dim": d, "radius": r, "queries": nq, "nnz": nnz, "mean_degree": nnz / nq,
                                  "seconds_best": best_dev, "seconds_with_fetch": best,
                                  "rows_per_s": nq / best_dev, "pair_tests_per_s": nq * N / best_dev,
                                  "algorithmic_bytes": algo, "achieved_GBps": algo / best_dev / 1e9,
                                  "note": "CSR left on the device (seconds_with_fetch adds the D2H); every query scans the whole (L2-resident) table: "
                                          "FP64-bound, %d distance evaluations" % (nq * N)})
    # roadmap edges from those rows
    rows_id = np.repeat(q, np.diff(rp.astype(np.int64)))
    keep = rows_id != col
    fi, ti = rows_id[keep], col[keep]
    reps = max(1, (10 ** 7 if not quick else 10 ** 6) // max(fi.size, 1))
    fi, ti = np.tile(fi, reps), np.tile(ti, reps)
    E = fi.size
    dev = torch.device("cuda:0")
    tf, tt = torch.from_numpy(fi.astype(np.int32)).to(dev), torch.from_numpy(ti.astype(np.int32)).to(dev)
    tv = torch.empty(E, dtype=torch.uint8, device=dev)
    tfi = torch.empty(E, dtype=torch.int32, device=dev)
    tn = torch.empty(E, dtype=torch.int32, device=dev)
    sec = device_timed(ctx, lambda: ctx.edges_check_indexed_dev(tf.data_ptr(), tt.data_ptr(), E, tv.data_ptr(),
                                                                tfi.data_ptr(), tn.data_ptr()))
    algo_e = E * (8 + 5) + N * d * 8
    emit("config3_edges_indexed", {"edges": int(E), "blocked_fraction": float((tv == 0).float().mean()),
                                   "mean_nstates": float(tn.float().mean()), "seconds_device": sec,
                                   "edges_per_s": E / sec, "algorithmic_bytes": algo_e,
                                   "achieved_GBps": algo_e / sec / 1e9, "frac_of_hbm_roofline": algo_e / sec / 1e9 / PEAK})
    ctx.close()


def belief(quick):
    d = 7
    M = 20000 if quick else 200000
    nq = 20000 if quick else 100000
    ctx = capi.Context(d, workloads.segment_length(d))
    pts = workloads.uniform(31, (M, d))
    vals = (workloads.uniform(32, (M,)) < 0.3).astype(np.float64)
    q = workloads.uniform(33, (nq, d))
    ctx.belief_append(pts, vals)
    out = [None]

    def run():
        out[0] = ctx.belief_estimate(q)
    best, mean = timed(run, reps=2)
    algo = nq * d * 8 + M * (d + 1) * 8 + nq * 8
    from oracle import oracle
    m = 200
    t0 = time.perf_counter()
    ref = oracle.belief_estimate(pts, vals, q[:m], nthreads=oracle.max_threads())
    cpu = m / (time.perf_counter() - t0)
    assert np.array_equal(ref, out[0][:m])
    emit("belief_knn", {"points": M, "queries": nq, "k": 15, "dim": d, "seconds_best": best, "queries_per_s": nq / best,
                        "pair_tests_per_s": nq * M / best, "algorithmic_bytes": algo, "achieved_GBps": algo / best / 1e9,
                        "cpu_baseline_queries_per_s": cpu, "cpu_cores": oracle.max_threads(),
                        "note": "brute force: FP64-bound (%d distance evaluations), not HBM-bound" % (nq * M)})
    ctx.close()


def _plan(make, label, max_calls, max_solutions=3, time_cap=60.0):
    from batching_pomp_b200.planner import Planner  # noqa: F401
    gp, op, start, goal, single = make()
    res = {}
    for name, p in (("gpu", gp), ("cpu_oracle", op)):
        if p is None:
            continue
        t0 = time.perf_counter()
        if single:
            p.setup(); p.set_problem(start, goal)
        else:
            p.set_problem(start, goal); p.setup()
        first = None
        costs = []
        st = "UNKNOWN"
        for _ in range(max_calls):
            left = time_cap - (time.perf_counter() - t0)
            if left <= 0:
                st = "BENCH_TIME_CAP"
                break
            st = p.solve(max_seconds=left)
            if len(p.path()) and (not costs or p.best_cost < costs[-1][1]):
                costs.append((time.perf_counter() - t0, p.best_cost))
                if first is None:
                    first = time.perf_counter() - t0
            if st in ("EXACT_SOLUTION", "TIMEOUT") or len(costs) >= max_solutions:
                break
        res[name] = {"time_to_first_path_s": first, "time_to_best_path_s": costs[-1][0] if costs else None,
                     "anytime_curve": [[round(t, 4), c] for t, c in costs],
                     "best_cost": costs[-1][1] if costs else None, "solutions": len(costs), "final_status": st,
                     "total_s": time.perf_counter() - t0, "counters": p.counters(),
                     "phase_seconds": p.times() if hasattr(p, "times") else None,
                     "gpu_launches": p.gpu_launches if hasattr(p, "gpu_launches") else None,
                     "path": [int(v) for v in p.path()]}
    if "cpu_oracle" in res:
        same_work = res["gpu"]["solutions"] == res["cpu_oracle"]["solutions"]
        res["paths_identical"] = (res["gpu"]["path"] == res["cpu_oracle"]["path"] and
                                  res["gpu"]["best_cost"] == res["cpu_oracle"]["best_cost"]) if same_work else None
    for k in ("gpu", "cpu_oracle"):
        if k in res:
            res[k].pop("path")
    emit(label, res)


def planner_config1(quick):
    from batching_pomp_b200.planner import Planner
    from oracle import oracle
    d, N, frac = 2, (3000 if quick else 10 ** 4), 1e-3
    pix = workloads.image_world(1024, seed=1)
    roadmap = workloads.halton(N, d)
    params = {"batching_type": "hybrid", "graph_type": "halton", "increment": 0.2, "prune_threshold": 1.05}

    def make():
        gp = Planner(d, segment_fraction=frac, image=pix)
        op = oracle.OraclePlanner(d, workloads.segment_length(d, frac), workloads.max_extent(d), 1.0)
        op.set_world_image(pix)
        for p in (gp, op):
            p.set_roadmap(roadmap)
            for k, v in params.items():
                p.set_param(k, v)
        return gp, op, [0.1, 0.1], [0.9, 0.9], False
    _plan(make, "planner_config1", 400, max_solutions=(6 if quick else 12), time_cap=(30.0 if quick else 100.0))


def planner_config3(quick):
    from batching_pomp_b200.planner import Planner
    from oracle import oracle
    d, N = 7, (2000 if quick else 20000)
    lo, hi = workloads.box_world(d, 500, seed=3)
    roadmap = workloads.halton(N, d)

    def make():
        gp = Planner(d, boxes=(lo, hi))
        op = oracle.OraclePlanner(d, workloads.segment_length(d), workloads.max_extent(d), 1.0)
        op.set_world_boxes(lo, hi)
        for p in (gp, op):
            p.set_roadmap(roadmap)
            p.set_param("batching_type", "vertex")
        return gp, op, np.full(d, 0.1), np.full(d, 0.9), False
    _plan(make, "planner_config3", 400, max_solutions=(5 if quick else 9), time_cap=(30.0 if quick else 100.0))


def planner_full_scale(quick):
    """Configs 2 and 3 as planner runs at full roadmap size (GPU planner only: the brute-force CPU oracle
    cannot expand vertices of a 10^6-vertex roadmap in reasonable time)."""
    from batching_pomp_b200.planner import Planner
    N = 10 ** 5 if quick else 10 ** 6
    # config 2: 2-D image, edge batching, selector_type=normal
    pix = workloads.image_world(4096, seed=2)
    roadmap = workloads.halton(N, 2)

    def make2():
        gp = Planner(2, segment_fraction=1e-4, image=pix)
        gp.set_roadmap_view(roadmap)
        for k, v in (("batching_type", "edge"), ("graph_type", "halton"), ("selector_type", "normal")):
            gp.set_param(k, v)
        return gp, None, [0.1, 0.1], [0.9, 0.9], False
    _plan(make2, "planner_config2_full", 400, max_solutions=4, time_cap=60.0)
    # config 3: 7-DoF boxes, vertex batching
    lo, hi = workloads.box_world(7, 500, seed=3)
    roadmap7 = workloads.halton(N, 7)

    def make3():
        gp = Planner(7, boxes=(lo, hi))
        gp.set_roadmap_view(roadmap7)
        gp.set_param("batching_type", "vertex")
        return gp, None, np.full(7, 0.1), np.full(7, 0.9), False
    _plan(make3, "planner_config3_full", 400, max_solutions=6, time_cap=60.0)


SECTIONS = {"planner_full": planner_full_scale, "config2": config2, "config3": config3, "belief": belief, "planner_config1": planner_config1,
            "planner_config3": planner_config3}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    names = a.only.split(",") if a.only else list(SECTIONS)
    OUT[0] = a.out
    for n in names:
        SECTIONS[n](a.quick)
