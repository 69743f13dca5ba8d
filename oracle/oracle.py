"""ctypes bindings for the CPU oracle (oracle/libbpomp_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
``--impl reference`` legs of bench.py.  The product package (batching_pomp_b200) never imports it.
PARITY UNPINNED (except BisectPerm, pinned to the reference's own header through oracle/_ref) — see the
header of bpomp_oracle.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbpomp_oracle.so")
_lib = None

c_dp = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)
c_i32p = C.POINTER(C.c_int32)
c_u32p = C.POINTER(C.c_uint32)
c_u64p = C.POINTER(C.c_uint64)
c_i64p = C.POINTER(C.c_int64)


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("bpomp_oracle.cpp", "planner_oracle.cpp")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libbpomp_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_distance.restype = C.c_double
        _lib.orc_nstates.restype = C.c_uint
        _lib.orc_nstates.argtypes = [C.c_double, C.c_double]
        _lib.orc_edge_states.restype = C.c_uint
        _lib.orc_belief_step.restype = C.c_uint
        _lib.orc_halton_radius.restype = C.c_double
        _lib.orc_rgg_radius.restype = C.c_double
        _lib.orc_rgg_radius.argtypes = [C.c_uint, C.c_uint, C.c_double]
        _lib.orc_unit_nball_measure.restype = C.c_double
        _lib.orcp_create.restype = C.c_void_p
        _lib.orcp_create.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
        _lib.orcp_error.restype = C.c_char_p
        _lib.orcp_best_cost.restype = C.c_double
        _lib.orcp_alpha.restype = C.c_double
        _lib.orcp_radius.restype = C.c_double
        _lib.orcp_num_vertices.restype = C.c_uint64
        _lib.orcp_num_edges.restype = C.c_uint64
        _lib.orcp_num_belief.restype = C.c_uint64
        _lib.orcp_num_batches.restype = C.c_uint
        for f in ("orcp_destroy", "orcp_world_boxes", "orcp_world_image", "orcp_roadmap", "orcp_set_param",
                  "orcp_setup", "orcp_set_problem", "orcp_set_ptc_budget", "orcp_solve", "orcp_error",
                  "orcp_best_cost", "orcp_alpha", "orcp_radius", "orcp_exhausted", "orcp_num_batches",
                  "orcp_path", "orcp_vertex_state", "orcp_counters", "orcp_num_vertices", "orcp_num_edges",
                  "orcp_num_belief"):
            getattr(_lib, f).argtypes = None
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t):
    return a.ctypes.data_as(t)


def max_threads():
    return int(lib().orc_max_threads())


_REF_PATH = os.path.join(_HERE, "_ref", "libbpomp_ref.so")
_ref = None


def ref_available():
    """True when the reference's own BisectPerm.hpp was compiled here (oracle/Makefile target `ref`)."""
    return os.path.exists(_REF_PATH)


def ref_bisect_perm(n):
    """BisectPerm::get(n) of the reference itself (oracle/ref_shim.cpp includes the header from /root/reference)."""
    global _ref
    if _ref is None:
        _ref = C.CDLL(_REF_PATH)
    first = np.zeros(max(n, 1), dtype=np.int32)
    second = np.zeros(max(n, 1), dtype=np.int32)
    m = _ref.ref_bisect_perm(C.c_int(n), _p(first, c_i32p), _p(second, c_i32p))
    return first[:m].copy(), second[:m].copy()


def bisect_perm(n):
    first = np.zeros(max(n, 1), dtype=np.int32)
    second = np.zeros(max(n, 1), dtype=np.int32)
    m = lib().orc_bisect_perm(C.c_int(n), _p(first, c_i32p), _p(second, c_i32p))
    return first[:m].copy(), second[:m].copy()


def distance(a, b):
    a, b = _f64(a), _f64(b)
    return float(lib().orc_distance(_p(a, c_dp), _p(b, c_dp), C.c_int(a.size)))


def interpolate(a, b, t):
    a, b = _f64(a), _f64(b)
    out = np.empty_like(a)
    lib().orc_interpolate(_p(a, c_dp), _p(b, c_dp), C.c_double(t), C.c_int(a.size), _p(out, c_dp))
    return out


def nstates(dist, L):
    return int(lib().orc_nstates(float(dist), float(L)))


def edge_states(a, b, L):
    a, b = _f64(a), _f64(b)
    d = a.size
    n = lib().orc_edge_states(_p(a, c_dp), _p(b, c_dp), C.c_int(d), C.c_double(L), None, C.c_uint(0))
    out = np.empty((n, d), dtype=np.float64)
    lib().orc_edge_states(_p(a, c_dp), _p(b, c_dp), C.c_int(d), C.c_double(L), _p(out, c_dp), C.c_uint(n))
    return out


def belief_step(n):
    return int(lib().orc_belief_step(C.c_uint(n)))


def states_valid_boxes(states, lo, hi):
    states, lo, hi = _f64(states), _f64(lo), _f64(hi)
    cnt, d = states.shape
    out = np.empty(cnt, dtype=np.uint8)
    lib().orc_states_valid_boxes(_p(states, c_dp), C.c_int64(cnt), C.c_int(d), _p(lo, c_dp), _p(hi, c_dp),
                                 C.c_int(lo.shape[0]), _p(out, c_u8p))
    return out


def states_valid_image(states, pix, thr=128):
    states = _f64(states)
    pix = np.ascontiguousarray(pix, dtype=np.uint8)
    H, W = pix.shape
    out = np.empty(states.shape[0], dtype=np.uint8)
    lib().orc_states_valid_image(_p(states, c_dp), C.c_int64(states.shape[0]), _p(pix, c_u8p), C.c_int(W),
                                 C.c_int(H), C.c_int(thr), _p(out, c_u8p))
    return out


def edges_check_boxes(frm, to, L, lo, hi, nthreads=1, refstyle=False):
    frm, to, lo, hi = _f64(frm), _f64(to), _f64(lo), _f64(hi)
    cnt, d = frm.shape
    valid = np.empty(cnt, dtype=np.uint8)
    first = np.empty(cnt, dtype=np.int32)
    ns = np.empty(cnt, dtype=np.uint32)
    if refstyle:
        lib().orc_edges_check_boxes_refstyle(_p(frm, c_dp), _p(to, c_dp), C.c_int64(cnt), C.c_int(d),
                                             C.c_double(L), _p(lo, c_dp), _p(hi, c_dp), C.c_int(lo.shape[0]),
                                             _p(valid, c_u8p), _p(first, c_i32p), _p(ns, c_u32p))
    else:
        lib().orc_edges_check_boxes(_p(frm, c_dp), _p(to, c_dp), C.c_int64(cnt), C.c_int(d), C.c_double(L),
                                    _p(lo, c_dp), _p(hi, c_dp), C.c_int(lo.shape[0]), _p(valid, c_u8p),
                                    _p(first, c_i32p), _p(ns, c_u32p), C.c_int(nthreads))
    return valid, first, ns


def edges_check_image(frm, to, L, pix, thr=128, nthreads=1):
    frm, to = _f64(frm), _f64(to)
    pix = np.ascontiguousarray(pix, dtype=np.uint8)
    H, W = pix.shape
    cnt = frm.shape[0]
    valid = np.empty(cnt, dtype=np.uint8)
    first = np.empty(cnt, dtype=np.int32)
    ns = np.empty(cnt, dtype=np.uint32)
    lib().orc_edges_check_image(_p(frm, c_dp), _p(to, c_dp), C.c_int64(cnt), C.c_double(L), _p(pix, c_u8p),
                                C.c_int(W), C.c_int(H), C.c_int(thr), _p(valid, c_u8p), _p(first, c_i32p),
                                _p(ns, c_u32p), C.c_int(nthreads))
    return valid, first, ns


def neighbors_radius(verts, query_ids, r, active=None, nthreads=1):
    verts = _f64(verts)
    N, d = verts.shape
    q = np.ascontiguousarray(query_ids, dtype=np.uint32)
    act = None if active is None else np.ascontiguousarray(active, dtype=np.uint8)
    actp = None if act is None else _p(act, c_u8p)
    row_ptr = np.zeros(q.size + 1, dtype=np.uint64)
    lib().orc_neighbors_radius(_p(verts, c_dp), actp, C.c_int64(N), C.c_int(d), _p(q, c_u32p),
                               C.c_int64(q.size), C.c_double(r), _p(row_ptr, c_u64p), None, None,
                               C.c_int(nthreads))
    nnz = int(row_ptr[-1])
    col = np.empty(max(nnz, 1), dtype=np.uint32)
    dist = np.empty(max(nnz, 1), dtype=np.float64)
    lib().orc_neighbors_radius(_p(verts, c_dp), actp, C.c_int64(N), C.c_int(d), _p(q, c_u32p),
                               C.c_int64(q.size), C.c_double(r), _p(row_ptr, c_u64p), _p(col, c_u32p),
                               _p(dist, c_dp), C.c_int(nthreads))
    return row_ptr, col[:nnz], dist[:nnz]


def vertices_inadmissible(states, start, goal, best_cost):
    states, start, goal = _f64(states), _f64(start), _f64(goal)
    cnt, d = states.shape
    out = np.empty(cnt, dtype=np.uint8)
    lib().orc_vertices_inadmissible(_p(states, c_dp), C.c_int64(cnt), C.c_int(d), _p(start, c_dp),
                                    _p(goal, c_dp), C.c_double(best_cost), _p(out, c_u8p))
    return out


def belief_estimate(pts, vals, queries, K=15, prior=0.5, prior_weight=0.25, nthreads=1):
    queries = _f64(queries)
    nq, d = queries.shape
    pts = _f64(pts).reshape(-1, d)
    vals = _f64(vals)
    out = np.empty(nq, dtype=np.float64)
    lib().orc_belief_estimate(_p(pts, c_dp), _p(vals, c_dp), C.c_int64(vals.size), C.c_int(d),
                              _p(queries, c_dp), C.c_int64(nq), C.c_int(K), C.c_double(prior),
                              C.c_double(prior_weight), _p(out, c_dp), C.c_int(nthreads))
    return out


def belief_knn_ids(pts, q, K=15):
    pts, q = _f64(pts), _f64(q)
    M, d = pts.shape
    k = min(K, M)
    ids = np.empty(max(k, 1), dtype=np.int64)
    dists = np.empty(max(k, 1), dtype=np.float64)
    lib().orc_belief_knn_ids(_p(pts, c_dp), C.c_int64(M), C.c_int(d), _p(q, c_dp), C.c_int(K), _p(ids, c_i64p),
                             _p(dists, c_dp))
    return ids[:k], dists[:k]


def belief_edge_pfree(frm, to, L, pts, vals, K=15, prior=0.5, prior_weight=0.25, nthreads=1):
    frm, to = _f64(frm), _f64(to)
    cnt, d = frm.shape
    pts = _f64(pts).reshape(-1, d)
    vals = _f64(vals)
    pf = np.empty(cnt, dtype=np.float64)
    nl = np.empty(cnt, dtype=np.uint32)
    lib().orc_belief_edge_pfree(_p(frm, c_dp), _p(to, c_dp), C.c_int64(cnt), C.c_int(d), C.c_double(L),
                                _p(pts, c_dp), _p(vals, c_dp), C.c_int64(vals.size), C.c_int(K),
                                C.c_double(prior), C.c_double(prior_weight), _p(pf, c_dp), _p(nl, c_u32p),
                                C.c_int(nthreads))
    return pf, nl


def halton_radius(n, d):
    return float(lib().orc_halton_radius(C.c_uint(n), C.c_uint(d)))


def rgg_radius(n, d, measure=1.0):
    return float(lib().orc_rgg_radius(C.c_uint(n), C.c_uint(d), C.c_double(measure)))


def halton(first_index, count, d):
    out = np.empty((count, d), dtype=np.float64)
    lib().orc_halton(C.c_int64(first_index), C.c_int64(count), C.c_int(d), _p(out, c_dp))
    return out


STATUS = {0: "UNKNOWN", 1: "INVALID_START", 2: "INVALID_GOAL", 4: "TIMEOUT", 5: "APPROXIMATE_SOLUTION",
          6: "EXACT_SOLUTION", 8: "ABORT"}


class OraclePlanner:
    """Literal CPU restatement of batching_pomp::BatchingPOMP (planner_oracle.cpp)."""

    def __init__(self, d, L, max_extent, space_measure=1.0):
        self._l = lib()
        self.d = d
        self._h = C.c_void_p(self._l.orcp_create(C.c_int(d), C.c_double(L), C.c_double(max_extent),
                                                 C.c_double(space_measure)))

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.orcp_destroy(self._h)
            self._h = None

    def set_world_boxes(self, lo, hi):
        lo, hi = _f64(lo), _f64(hi)
        self._l.orcp_world_boxes(self._h, _p(lo, c_dp), _p(hi, c_dp), C.c_int(lo.shape[0]))

    def set_world_image(self, pix, thr=128):
        pix = np.ascontiguousarray(pix, dtype=np.uint8)
        self._l.orcp_world_image(self._h, _p(pix, c_u8p), C.c_int(pix.shape[1]), C.c_int(pix.shape[0]),
                                 C.c_int(thr))

    def set_roadmap(self, states, edges=None):
        states = _f64(states)
        if edges is None or len(edges) == 0:
            self._l.orcp_roadmap(self._h, _p(states, c_dp), C.c_uint32(states.shape[0]), None, C.c_uint32(0))
        else:
            e = np.ascontiguousarray(edges, dtype=np.uint32)
            self._l.orcp_roadmap(self._h, _p(states, c_dp), C.c_uint32(states.shape[0]), _p(e, c_u32p),
                                 C.c_uint32(e.shape[0]))

    def set_param(self, key, val):
        r = self._l.orcp_set_param(self._h, key.encode(), str(val).encode())
        if r != 0:
            raise KeyError(key)

    def setup(self):
        r = self._l.orcp_setup(self._h)
        if r != 0:
            raise RuntimeError(self._l.orcp_error(self._h).decode())

    def set_problem(self, start, goal):
        start, goal = _f64(start), _f64(goal)
        r = self._l.orcp_set_problem(self._h, _p(start, c_dp), _p(goal, c_dp))
        if r != 0:
            raise RuntimeError(self._l.orcp_error(self._h).decode())

    def set_ptc_budget(self, n):
        self._l.orcp_set_ptc_budget(self._h, C.c_long(n))

    def set_time_limit(self, seconds):
        self._l.orcp_set_time_limit(self._h, C.c_double(seconds))

    def solve(self, max_seconds=0.0):
        if max_seconds:
            self.set_time_limit(max_seconds)
        return STATUS[int(self._l.orcp_solve(self._h))]

    @property
    def best_cost(self):
        return float(self._l.orcp_best_cost(self._h))

    @property
    def alpha(self):
        return float(self._l.orcp_alpha(self._h))

    @property
    def radius(self):
        return float(self._l.orcp_radius(self._h))

    @property
    def exhausted(self):
        return bool(self._l.orcp_exhausted(self._h))

    @property
    def num_batches(self):
        return int(self._l.orcp_num_batches(self._h))

    def path(self):
        buf = np.zeros(1 << 16, dtype=np.uint32)
        n = self._l.orcp_path(self._h, _p(buf, c_u32p), C.c_int(buf.size))
        return buf[:n].copy()

    def path_states(self):
        out = []
        for v in self.path():
            s = np.empty(self.d, dtype=np.float64)
            self._l.orcp_vertex_state(self._h, C.c_uint32(int(v)), _p(s, c_dp))
            out.append(s)
        return np.array(out).reshape(-1, self.d)

    def counters(self):
        c = np.zeros(4, dtype=np.uint32)
        self._l.orcp_counters(self._h, _p(c, c_u32p))
        return {"edge_checks": int(c[0]), "coll_checks": int(c[1]), "searches": int(c[2]), "lookups": int(c[3])}

    @property
    def num_vertices(self):
        return int(self._l.orcp_num_vertices(self._h))

    @property
    def num_edges(self):
        return int(self._l.orcp_num_edges(self._h))

    @property
    def num_belief(self):
        return int(self._l.orcp_num_belief(self._h))


# ---------------------------------------------------------------------------------------------------
# The REFERENCE itself (oracle/_ref/libbpomp_ref.so: /root/reference/src/BatchingPOMP.cpp and the headers it
# includes, compiled against the API stand-ins of oracle/standins — oracle/Makefile target `ref`,
# oracle/ref_planner_shim.cpp).  Only used to pin the oracle above; never timed, never shipped.
# ---------------------------------------------------------------------------------------------------
def _reflib():
    global _ref
    if _ref is None:
        _ref = C.CDLL(_REF_PATH)
    if not getattr(_ref, "_bound", False):
        if hasattr(_ref, "refp_create"):
            _ref.refp_create.restype = C.c_void_p
            _ref.refp_create.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
            _ref.refp_error.restype = C.c_char_p
            for f in ("refp_best_cost", "refp_alpha", "refp_radius", "refp_segment_length"):
                getattr(_ref, f).restype = C.c_double
            for f in ("refp_num_vertices", "refp_num_edges", "refp_solution_count"):
                getattr(_ref, f).restype = C.c_uint64
            _ref.refp_num_batches.restype = C.c_uint
        _ref._bound = True
    return _ref


def ref_planner_available():
    """True when the reference planner was compiled here (oracle/Makefile target `ref`)."""
    return ref_available() and hasattr(_reflib(), "refp_create")


def write_graphml(filename, states, edges=None):
    """GraphML as the reference reads it (util/RoadmapFromFile.hpp:82-95,136-147): node attribute `state` =
    space-separated doubles (repr round-trips exactly through `ss >> double`); optional edges."""
    states = _f64(states)
    with open(filename, "w") as f:
        f.write('<?xml version="1.0" encoding="UTF-8"?>\n<graphml xmlns="http://graphml.graphdrawing.org/xmlns">\n')
        f.write('<key id="d0" for="node" attr.name="state" attr.type="string"/>\n<graph edgedefault="undirected">\n')
        for i, s in enumerate(states):
            f.write('<node id="n%d"><data key="d0">%s</data></node>\n' % (i, " ".join(repr(float(x)) for x in s)))
        if edges is not None:
            for a, b in np.asarray(edges).reshape(-1, 2):
                f.write('<edge source="n%d" target="n%d"/>\n' % (int(a), int(b)))
        f.write('</graph>\n</graphml>\n')


class RefPlanner:
    """batching_pomp::BatchingPOMP — the reference's own class — on the unit hypercube [lo, hi]^d."""

    def __init__(self, d, fraction, lo=0.0, hi=1.0):
        self._l = _reflib()
        self.d = d
        self._h = C.c_void_p(self._l.refp_create(C.c_int(d), C.c_double(lo), C.c_double(hi), C.c_double(fraction)))

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.refp_destroy(self._h)
            self._h = None

    @property
    def segment_length(self):
        return float(self._l.refp_segment_length(self._h))

    def set_world_boxes(self, lo, hi):
        lo, hi = _f64(lo), _f64(hi)
        self._l.refp_world_boxes(self._h, _p(lo, c_dp), _p(hi, c_dp), C.c_int(lo.shape[0]))

    def set_world_image(self, pix, thr=128):
        pix = np.ascontiguousarray(pix, dtype=np.uint8)
        self._l.refp_world_image(self._h, _p(pix, c_u8p), C.c_int(pix.shape[1]), C.c_int(pix.shape[0]), C.c_int(thr))

    def set_param(self, key, val):
        if self._l.refp_set_param(self._h, key.encode(), str(val).encode()) != 0:
            raise KeyError(key)

    def setup(self):
        r = self._l.refp_setup(self._h)
        if r != 0:
            raise RuntimeError(self._l.refp_error(self._h).decode())

    def set_problem(self, start, goal):
        start, goal = _f64(start), _f64(goal)
        if self._l.refp_set_problem(self._h, _p(start, c_dp), _p(goal, c_dp)) != 0:
            raise RuntimeError(self._l.refp_error(self._h).decode())

    def set_ptc_budget(self, n):
        self._l.refp_set_ptc_budget(self._h, C.c_long(n))

    def solve(self):
        st = int(self._l.refp_solve(self._h))
        if st < 0:
            raise RuntimeError(self._l.refp_error(self._h).decode())
        return STATUS[st]

    best_cost = property(lambda self: float(self._l.refp_best_cost(self._h)))
    alpha = property(lambda self: float(self._l.refp_alpha(self._h)))
    radius = property(lambda self: float(self._l.refp_radius(self._h)))
    exhausted = property(lambda self: bool(self._l.refp_exhausted(self._h)))
    num_batches = property(lambda self: int(self._l.refp_num_batches(self._h)))
    num_vertices = property(lambda self: int(self._l.refp_num_vertices(self._h)))
    num_edges = property(lambda self: int(self._l.refp_num_edges(self._h)))
    solution_count = property(lambda self: int(self._l.refp_solution_count(self._h)))

    def path_states(self):
        out = np.zeros((4096, self.d), dtype=np.float64)
        n = int(self._l.refp_path_states(self._h, _p(out, c_dp), C.c_int(4096)))
        return out[:n].copy()

    def counters(self):
        c = np.zeros(4, dtype=np.uint64)
        self._l.refp_counters(self._h, _p(c, c_u64p))
        return {"edge_checks": int(c[0]), "coll_checks": int(c[1]), "searches": int(c[2]), "lookups": int(c[3])}


def ref_knn_estimate(pts, vals, queries, K=15, prior=0.5, prior_weight=0.25):
    """cspacebelief::KNNModel::addPoint / estimate of the reference (KNNModel.hpp:89-134)."""
    pts, vals, queries = _f64(pts), _f64(vals), _f64(queries)
    d = queries.shape[1]
    out = np.empty(queries.shape[0], dtype=np.float64)
    rc = _reflib().ref_knn_estimate(_p(pts, c_dp), _p(vals, c_dp), C.c_int64(vals.shape[0]), C.c_int(d),
                                    _p(queries, c_dp), C.c_int64(queries.shape[0]), C.c_int(K), C.c_double(prior),
                                    C.c_double(prior_weight), _p(out, c_dp))
    if rc != 0:
        raise RuntimeError("ref_knn_estimate failed")
    return out


def ref_edges_available():
    return ref_available() and hasattr(_reflib(), "ref_edges_check_boxes")


def ref_edges_check_boxes(frm, to, fraction, lo, hi, nthreads=1):
    """The reference's own initializeEdgePoints + checkAndSetEdgeBlocked (src/BatchingPOMP.cpp:435-461, 496-544) on
    arbitrary edges of the unit hypercube against a box world: (valid, first_invalid_slot, nstates).  `fraction` is
    the state space's longestValidSegmentFraction (the planner derives L = maxExtent * fraction itself)."""
    frm, to, lo, hi = _f64(frm), _f64(to), _f64(lo), _f64(hi)
    n, d = frm.shape
    valid = np.empty(n, dtype=np.uint8)
    first = np.empty(n, dtype=np.int32)
    ns = np.empty(n, dtype=np.uint32)
    rc = _reflib().ref_edges_check_boxes(C.c_int(d), C.c_double(fraction), _p(lo, c_dp), _p(hi, c_dp),
                                         C.c_int(lo.shape[0]), _p(frm, c_dp), _p(to, c_dp), C.c_int64(n),
                                         valid.ctypes.data_as(C.c_void_p), first.ctypes.data_as(C.c_void_p),
                                         ns.ctypes.data_as(C.c_void_p), C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError("ref_edges_check_boxes failed")
    return valid, first, ns


def ref_selector(sel_type, prob_free):
    """util::Selector<Graph>::selectEdges of the reference: evaluation order of a path's edges (positions)."""
    p = _f64(prob_free)
    order = np.zeros(max(p.size, 1), dtype=np.int32)
    n = int(_reflib().ref_selector(sel_type.encode(), _p(p, c_dp), C.c_int(p.size), _p(order, c_i32p)))
    if n < 0:
        raise ValueError("Invalid selector type specified - %s!" % sel_type)
    return order[:n].copy()
