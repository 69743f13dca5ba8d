#!/usr/bin/env python
"""Golden vectors taken from the REFERENCE ITSELF, compiled by oracle/Makefile (target `ref`) from where it
lies under /root/reference into oracle/_ref/libbpomp_ref.so:
  bisect_perm_ref.npz   BisectPerm::get(n) of util/BisectPerm.hpp (oracle/ref_shim.cpp)
  planner_ref_*.npz     whole runs of batching_pomp::BatchingPOMP (src/BatchingPOMP.cpp + every header it includes,
                        against the API stand-ins of oracle/standins; oracle/ref_planner_shim.cpp): status, best
                        cost, alpha, radius, counters and path states after every solve() call of the scenarios in
                        tests/refcases.py
  knn_ref.npz           cspacebelief::KNNModel::estimate (KNNModel.hpp:101-134) on seeded points / queries
Run in the build container (the reference is not on the GPU box); the .npz files travel.

    python tests/golden/make_golden_ref.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

NS = list(range(0, 257)) + [511, 512, 513, 1000, 4097, 10007]


def main():
    if not oracle.ref_available():
        raise SystemExit("oracle/_ref/libbpomp_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    first, second, offs = [], [], [0]
    for n in NS:
        f, s = oracle.ref_bisect_perm(n)
        first.append(f)
        second.append(s)
        offs.append(offs[-1] + f.size)
    out = os.path.join(ROOT, "tests", "golden", "bisect_perm_ref.npz")
    np.savez_compressed(out, n=np.asarray(NS, dtype=np.int32), offsets=np.asarray(offs, dtype=np.int64),
                        first=np.concatenate(first).astype(np.int32), second=np.concatenate(second).astype(np.int32))
    print("wrote", out, os.path.getsize(out), "bytes")
    planner_and_knn()


def planner_and_knn():
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refcases
    from batching_pomp_b200 import workloads as wl
    if not oracle.ref_planner_available():
        raise SystemExit("the reference planner is not in oracle/_ref/libbpomp_ref.so: run `make -C oracle ref`")
    for name, case in refcases.cases().items():
        roadmap = wl.halton(case["N"], case["d"])
        fn = tempfile.mktemp(suffix=".graphml")
        oracle.write_graphml(fn, roadmap, case["edges"])
        rp = oracle.RefPlanner(case["d"], case["fraction"])
        rec = refcases.drive(rp, case, lambda p: p.set_param("roadmap_filename", fn))
        os.unlink(fn)
        out = os.path.join(ROOT, "tests", "golden", "planner_ref_%s.npz" % name)
        np.savez_compressed(out, **refcases.to_npz(rec))
        print("wrote", out, os.path.getsize(out), "bytes;", len(rec["status"]), "solve() calls, final", rec["status"][-1],
              "cost", rec["cost"][-1], "counters", rec["counters"][-1])
    # KNNModel
    recs = {}
    for tag, (d, M, nq, K) in {"a": (7, 3000, 500, 15), "b": (2, 300, 400, 15), "c": (7, 7, 50, 15), "d": (3, 2000, 200, 1),
                               "e": (5, 1000, 200, 32)}.items():
        pts = wl.uniform(131, (M, d))
        vals = (wl.uniform(132, (M,)) < 0.3).astype(np.float64)
        q = wl.uniform(133, (nq, d))
        q[: min(nq, M) // 2] = pts[: min(nq, M) // 2]  # exact hits: the 1e-4 distance clamp
        recs["pts_" + tag], recs["vals_" + tag], recs["q_" + tag] = pts, vals, q
        recs["K_" + tag] = np.int32(K)
        recs["est_" + tag] = oracle.ref_knn_estimate(pts, vals, q, K=K)
    out = os.path.join(ROOT, "tests", "golden", "knn_ref.npz")
    np.savez_compressed(out, **recs)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
