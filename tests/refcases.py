"""Planner scenarios shared by the reference-pinning tests and tests/golden/make_golden_ref.py.  Each case is
small enough for the reference's own single-threaded planner (oracle/_ref) to finish in seconds."""
import numpy as np

from batching_pomp_b200 import workloads as wl


def _single_edges(N, d, r):
    roadmap = wl.halton(N, d)
    diff = roadmap[:, None, :] - roadmap[None, :, :]
    dist = np.sqrt((diff ** 2).sum(-1))
    iu, ju = np.nonzero(np.triu(dist <= r, 1))
    return np.stack([iu, ju], 1).astype(np.uint32)


def cases():
    """name -> dict(d, N, world, params, fraction, start, goal, max_calls, budget, edges, single)"""
    out = {}
    pix1 = ("image", 256, 1)
    pix2 = ("image", 256, 2)
    boxes = ("boxes", 7, 500, 3)
    out["hybrid_image_2d"] = dict(d=2, N=1500, world=pix1, fraction=2e-3, max_calls=14, budget=80,
                                  params={"batching_type": "hybrid", "graph_type": "halton", "increment": 0.2,
                                          "prune_threshold": 1.05})
    out["vertex_boxes_7d"] = dict(d=7, N=1200, world=boxes, fraction=0.01, max_calls=25, budget=120,
                                  params={"batching_type": "vertex"})
    out["edge_rgg_boxes_7d"] = dict(d=7, N=800, world=boxes, fraction=0.01, max_calls=25, budget=70,
                                    params={"batching_type": "edge", "graph_type": "rgg", "increment": 0.5})
    for sel in ("normal", "alternate", "failfast", "maxinf"):
        out["edge_halton_image_2d_" + sel] = dict(d=2, N=1500, world=pix2, fraction=2e-3, max_calls=60, budget=120,
                                                  params={"batching_type": "edge", "graph_type": "halton",
                                                          "selector_type": sel})
    r = 1.2 * wl.halton_radius(300, 2)
    out["single_image_2d"] = dict(d=2, N=300, world=pix1, fraction=4e-3, max_calls=20, budget=60, single=True,
                                  edges=_single_edges(300, 2, r),
                                  params={"batching_type": "single", "start_goal_radius": r})
    for c in out.values():
        c.setdefault("single", False)
        c.setdefault("edges", None)
        c["start"] = np.full(c["d"], 0.1)
        c["goal"] = np.full(c["d"], 0.9)
    return out


def world_of(spec):
    if spec[0] == "image":
        return ("image", wl.image_world(spec[1], seed=spec[2]))
    lo, hi = wl.box_world(spec[1], spec[2], seed=spec[3])
    return ("boxes", lo, hi)


def drive(planner, case, roadmap_setter):
    """Runs setup / setProblemDefinition / solve() as the case says and records every call.  roadmap_setter(planner)
    gives the planner its roadmap (file name or arrays)."""
    w = world_of(case["world"])
    if w[0] == "image":
        planner.set_world_image(w[1])
    else:
        planner.set_world_boxes(w[1], w[2])
    roadmap_setter(planner)
    for k, v in case["params"].items():
        planner.set_param(k, v)
    if case["single"]:
        planner.setup()
        planner.set_problem(case["start"], case["goal"])
    else:
        planner.set_problem(case["start"], case["goal"])
        planner.setup()
    planner.set_ptc_budget(case["budget"])
    rec = {"status": [], "cost": [], "alpha": [], "radius": [], "batches": [], "exhausted": [], "nv": [], "ne": [],
           "counters": [], "path_len": [], "path_states": []}
    for _ in range(case["max_calls"]):
        st = planner.solve()
        c = planner.counters()
        ps = np.asarray(planner.path_states(), dtype=np.float64).reshape(-1, case["d"])
        rec["status"].append(st)
        rec["cost"].append(planner.best_cost)
        rec["alpha"].append(planner.alpha)
        rec["radius"].append(planner.radius)
        rec["batches"].append(planner.num_batches)
        rec["exhausted"].append(int(planner.exhausted))
        rec["nv"].append(planner.num_vertices)
        rec["ne"].append(planner.num_edges)
        rec["counters"].append([c["edge_checks"], c["coll_checks"], c["searches"], c["lookups"]])
        rec["path_len"].append(ps.shape[0])
        rec["path_states"].append(ps)
        if st in ("EXACT_SOLUTION", "TIMEOUT"):
            break
    return rec


def same(a, b):
    """Call-by-call equality of two records (bit-exact doubles)."""
    if a["status"] != b["status"]:
        return "status: %s vs %s" % (a["status"], b["status"])
    for k in ("cost", "alpha", "radius", "batches", "exhausted", "nv", "ne", "counters", "path_len"):
        if not np.array_equal(np.asarray(a[k]), np.asarray(b[k])):
            i = int(np.nonzero(np.any(np.atleast_2d(np.asarray(a[k]).reshape(len(a[k]), -1) !=
                                                    np.asarray(b[k]).reshape(len(b[k]), -1)), axis=1))[0][0])
            return "%s differs at solve() call %d: %s vs %s" % (k, i, a[k][i], b[k][i])
    for i, (pa, pb) in enumerate(zip(a["path_states"], b["path_states"])):
        if not np.array_equal(pa, pb):
            return "path states differ at solve() call %d" % i
    return None


def to_npz(rec):
    return dict(status=np.array(rec["status"]), cost=np.array(rec["cost"]), alpha=np.array(rec["alpha"]),
                radius=np.array(rec["radius"]), batches=np.array(rec["batches"]), exhausted=np.array(rec["exhausted"]),
                nv=np.array(rec["nv"]), ne=np.array(rec["ne"]), counters=np.array(rec["counters"], dtype=np.int64),
                path_len=np.array(rec["path_len"]),
                path_cat=np.concatenate(rec["path_states"]) if rec["path_states"] else np.zeros((0, 1)))


def from_npz(z, d):
    lens = z["path_len"].tolist()
    cat = z["path_cat"].reshape(-1, d) if sum(lens) else np.zeros((0, d))
    ps, o = [], 0
    for n in lens:
        ps.append(cat[o:o + n])
        o += n
    return {"status": [str(s) for s in z["status"]], "cost": z["cost"].tolist(), "alpha": z["alpha"].tolist(),
            "radius": z["radius"].tolist(), "batches": z["batches"].tolist(), "exhausted": z["exhausted"].tolist(),
            "nv": z["nv"].tolist(), "ne": z["ne"].tolist(), "counters": z["counters"].tolist(), "path_len": lens,
            "path_states": ps}
