"""ctypes binding of the host planner (include/bpomp_planner.h -> libbpomp_host.so).

`Planner` mirrors batching_pomp::BatchingPOMP: the seven OMPL parameters by name
(/root/reference/src/BatchingPOMP.cpp:313-319), setup(), set_problem() (= setProblemDefinition) and
solve() returning the OMPL PlannerStatus name.  The hot operators run on the GPU through
libbpomp_cuda.so; there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# env override: tuning builds, and the CPU test-suite's copy of the library that sits next to a test double of
# libbpomp_cuda.so (tests/mock) — the product never sets it
LIB_PATH = os.environ.get("BPOMP_HOST_LIB") or os.path.join(_HERE, "libbpomp_host.so")
vp = C.c_void_p

SIGNATURES = {
    "bpomp_planner_create_boxes": (vp, [C.c_int, C.c_int, vp, vp, C.c_double, vp, vp, C.c_int]),
    "bpomp_planner_create_image": (vp, [C.c_int, vp, vp, C.c_double, vp, C.c_int, C.c_int, C.c_int]),
    "bpomp_planner_destroy": (None, [vp]),
    "bpomp_planner_error": (C.c_char_p, [vp]),
    "bpomp_planner_set_param": (C.c_int, [vp, C.c_char_p, C.c_char_p]),
    "bpomp_planner_set_roadmap": (C.c_int, [vp, vp, C.c_uint32, vp, C.c_uint32]),
    "bpomp_planner_set_roadmap_view": (C.c_int, [vp, vp, C.c_uint32]),
    "bpomp_planner_setup": (C.c_int, [vp]),
    "bpomp_planner_set_problem": (C.c_int, [vp, vp, vp]),
    "bpomp_planner_solve": (C.c_int, [vp, C.c_long, C.c_double, C.POINTER(C.c_int)]),
    "bpomp_planner_best_cost": (C.c_double, [vp]),
    "bpomp_planner_alpha": (C.c_double, [vp]),
    "bpomp_planner_radius": (C.c_double, [vp]),
    "bpomp_planner_exhausted": (C.c_int, [vp]),
    "bpomp_planner_num_batches": (C.c_uint32, [vp]),
    "bpomp_planner_path": (C.c_int, [vp, vp, C.c_int]),
    "bpomp_planner_vertex_state": (C.c_int, [vp, C.c_uint32, vp]),
    "bpomp_planner_counters": (None, [vp, vp, vp]),
    "bpomp_planner_device_seconds": (C.c_double, [vp]),
    "bpomp_planner_num_vertices": (C.c_uint64, [vp]),
    "bpomp_planner_num_edges": (C.c_uint64, [vp]),
    "bpomp_planner_gpu_launches": (C.c_uint64, [vp]),
    "bpomp_graphml_write": (C.c_int, [C.c_char_p, C.c_int, vp, C.c_uint32, vp, C.c_uint32]),
    "bpomp_graphml_read": (C.c_int, [C.c_char_p, C.c_int, vp, C.POINTER(C.c_uint32), vp, C.POINTER(C.c_uint32)]),
    "bpomp_graphml_read_cached": (C.c_int, [C.c_char_p, C.c_int, vp, C.POINTER(C.c_uint32), vp, C.POINTER(C.c_uint32),
                                            C.POINTER(C.c_int)]),
}

STATUS = {0: "UNKNOWN", 1: "INVALID_START", 2: "INVALID_GOAL", 3: "UNRECOGNIZED_GOAL_TYPE", 4: "TIMEOUT",
          5: "APPROXIMATE_SOLUTION", 6: "EXACT_SOLUTION", 7: "CRASH", 8: "ABORT"}

_lib = None


class PlannerError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s not found: run __graft_entry__.build()" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Planner:
    def __init__(self, dim, low=None, high=None, segment_fraction=0.01, boxes=None, image=None, threshold=128,
                 device=0):
        self.lib = load()
        self.dim = dim
        low = _f64(np.zeros(dim) if low is None else low)
        high = _f64(np.ones(dim) if high is None else high)
        if boxes is not None:
            lo, hi = _f64(boxes[0]).reshape(-1, dim), _f64(boxes[1]).reshape(-1, dim)
            h = self.lib.bpomp_planner_create_boxes(device, dim, _p(low), _p(high), float(segment_fraction), _p(lo),
                                                    _p(hi), lo.shape[0])
        elif image is not None:
            pix = np.ascontiguousarray(image, dtype=np.uint8)
            h = self.lib.bpomp_planner_create_image(device, _p(low), _p(high), float(segment_fraction), _p(pix),
                                                    pix.shape[1], pix.shape[0], int(threshold))
        else:
            raise ValueError("a collision world (boxes= or image=) is required")
        if not h:
            raise PlannerError(-1, self.lib.bpomp_planner_error(None).decode())
        self.h = vp(h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.bpomp_planner_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise PlannerError(rc, self.lib.bpomp_planner_error(self.h).decode())

    def set_param(self, name, value):
        self._ck(self.lib.bpomp_planner_set_param(self.h, name.encode(), str(value).encode()))

    def set_roadmap(self, states, edges=None):
        s = _f64(states).reshape(-1, self.dim)
        if edges is None or len(edges) == 0:
            self._ck(self.lib.bpomp_planner_set_roadmap(self.h, _p(s), s.shape[0], None, 0))
        else:
            e = np.ascontiguousarray(edges, dtype=np.uint32).reshape(-1, 2)
            self._ck(self.lib.bpomp_planner_set_roadmap(self.h, _p(s), s.shape[0], _p(e), e.shape[0]))

    def set_roadmap_view(self, states):
        """Non-owning: the (C-contiguous float64) array is kept alive by this object and not copied."""
        assert states.dtype == np.float64 and states.flags["C_CONTIGUOUS"]
        self._roadmap_ref = states
        self._ck(self.lib.bpomp_planner_set_roadmap_view(self.h, _p(states), states.shape[0]))

    def setup(self):
        self._ck(self.lib.bpomp_planner_setup(self.h))

    def set_problem(self, start, goal):
        s, g = _f64(start), _f64(goal)
        self._ck(self.lib.bpomp_planner_set_problem(self.h, _p(s), _p(g)))

    def solve(self, max_iterations=-1, max_seconds=0.0):
        st = C.c_int(0)
        self._ck(self.lib.bpomp_planner_solve(self.h, int(max_iterations), float(max_seconds), C.byref(st)))
        return STATUS[st.value]

    @property
    def best_cost(self):
        return float(self.lib.bpomp_planner_best_cost(self.h))

    @property
    def alpha(self):
        return float(self.lib.bpomp_planner_alpha(self.h))

    @property
    def radius(self):
        return float(self.lib.bpomp_planner_radius(self.h))

    @property
    def exhausted(self):
        return bool(self.lib.bpomp_planner_exhausted(self.h))

    @property
    def num_batches(self):
        return int(self.lib.bpomp_planner_num_batches(self.h))

    def path(self):
        buf = np.zeros(1 << 16, dtype=np.uint32)
        n = self.lib.bpomp_planner_path(self.h, _p(buf), buf.size)
        return buf[:n].copy()

    def path_states(self):
        out = []
        for v in self.path():
            s = np.empty(self.dim, dtype=np.float64)
            self.lib.bpomp_planner_vertex_state(self.h, int(v), _p(s))
            out.append(s)
        return np.array(out).reshape(-1, self.dim)

    def counters(self):
        c = np.zeros(4, dtype=np.uint32)
        t = np.zeros(3, dtype=np.float64)
        self.lib.bpomp_planner_counters(self.h, _p(c), _p(t))
        return {"edge_checks": int(c[0]), "coll_checks": int(c[1]), "searches": int(c[2]), "lookups": int(c[3])}

    def times(self):
        c = np.zeros(4, dtype=np.uint32)
        t = np.zeros(3, dtype=np.float64)
        self.lib.bpomp_planner_counters(self.h, _p(c), _p(t))
        return {"lookup_s": float(t[0]), "search_s": float(t[1]), "collcheck_s": float(t[2]),
                "device_s": float(self.lib.bpomp_planner_device_seconds(self.h))}

    @property
    def num_vertices(self):
        return int(self.lib.bpomp_planner_num_vertices(self.h))

    @property
    def num_edges(self):
        return int(self.lib.bpomp_planner_num_edges(self.h))

    @property
    def gpu_launches(self):
        return int(self.lib.bpomp_planner_gpu_launches(self.h))


def write_graphml(filename, states, edges=None):
    s = _f64(states)
    n, d = s.shape
    if edges is None or len(edges) == 0:
        rc = load().bpomp_graphml_write(filename.encode(), d, _p(s), n, None, 0)
    else:
        e = np.ascontiguousarray(edges, dtype=np.uint32).reshape(-1, 2)
        rc = load().bpomp_graphml_write(filename.encode(), d, _p(s), n, _p(e), e.shape[0])
    if rc != 0:
        raise PlannerError(rc, load().bpomp_planner_error(None).decode())


def read_graphml(filename, dim):
    lib = load()
    n, ne = C.c_uint32(0), C.c_uint32(0)
    rc = lib.bpomp_graphml_read(filename.encode(), dim, None, C.byref(n), None, C.byref(ne))
    if rc != 0:
        raise PlannerError(rc, lib.bpomp_planner_error(None).decode())
    s = np.zeros((n.value, dim), dtype=np.float64)
    e = np.zeros((max(ne.value, 1), 2), dtype=np.uint32)
    rc = lib.bpomp_graphml_read(filename.encode(), dim, _p(s), C.byref(n), _p(e), C.byref(ne))
    if rc != 0:
        raise PlannerError(rc, lib.bpomp_planner_error(None).decode())
    return s, e[:ne.value]


def read_graphml_cached(filename, dim):
    """read_graphml through the binary cache beside the file (the path the planner's setup() takes).
    Returns (states, edges, from_cache)."""
    lib = load()
    n, ne, hit = C.c_uint32(0), C.c_uint32(0), C.c_int(0)
    rc = lib.bpomp_graphml_read_cached(filename.encode(), dim, None, C.byref(n), None, C.byref(ne), C.byref(hit))
    if rc != 0:
        raise PlannerError(rc, lib.bpomp_planner_error(None).decode())
    s = np.zeros((n.value, dim), dtype=np.float64)
    e = np.zeros((max(ne.value, 1), 2), dtype=np.uint32)
    ne.value = max(ne.value, 1)
    rc = lib.bpomp_graphml_read_cached(filename.encode(), dim, _p(s), C.byref(n), _p(e), C.byref(ne), C.byref(hit))
    if rc != 0:
        raise PlannerError(rc, lib.bpomp_planner_error(None).decode())
    return s, e[:ne.value], bool(hit.value)
