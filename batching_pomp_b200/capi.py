"""ctypes binding of the C-ABI CUDA library (include/bpomp_cuda.h -> libbpomp_cuda.so).

This is the harness-side binding used by the tests and bench.py; the reference-side binding a
maintainer would write is shown in INTEGRATION.md.  There is no CPU fallback: if the shared
library is missing or no CUDA device is present, loading / Context() raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BPOMP_CUDA_LIB") or os.path.join(_HERE, "libbpomp_cuda.so")  # env: tuning builds only

c_dp = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)
c_i32p = C.POINTER(C.c_int32)
c_u32p = C.POINTER(C.c_uint32)
c_u64p = C.POINTER(C.c_uint64)
vp = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/bpomp_cuda.h declares.
SIGNATURES = {
    "bpomp_create": (C.c_int, [C.c_int, C.c_int, C.c_double, C.POINTER(vp)]),
    "bpomp_destroy": (None, [vp]),
    "bpomp_last_error": (C.c_char_p, [vp]),
    "bpomp_sync": (C.c_int, [vp]),
    "bpomp_set_stream": (C.c_int, [vp, vp]),
    "bpomp_get_stream": (vp, [vp]),
    "bpomp_launch_count": (C.c_uint64, [vp]),
    "bpomp_perm_reserve": (C.c_int, [vp, C.c_uint32]),
    "bpomp_bisect_perm": (C.c_int, [C.c_uint32, c_i32p]),
    "bpomp_world_set_image": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int]),
    "bpomp_world_set_boxes": (C.c_int, [vp, vp, vp, C.c_int]),
    "bpomp_world_set_bounds": (C.c_int, [vp, vp, vp]),
    "bpomp_set_option": (C.c_int, [vp, C.c_char_p, C.c_int64]),
    "bpomp_get_counter": (C.c_int, [vp, C.c_char_p, C.POINTER(C.c_int64)]),
    "bpomp_states_valid": (C.c_int, [vp, vp, C.c_int64, vp]),
    "bpomp_states_valid_dev": (C.c_int, [vp, vp, C.c_int64, vp]),
    "bpomp_states_prune": (C.c_int, [vp, vp, C.c_int64, vp, vp, C.c_double, C.c_int, vp]),
    "bpomp_vertices_prune": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_double, C.c_int, vp]),
    "bpomp_roadmap_generate": (C.c_int, [vp, C.c_int, C.c_uint64, C.c_uint32, vp]),
    "bpomp_vertices_append": (C.c_int, [vp, vp, C.c_uint32, c_u32p]),
    "bpomp_vertices_set_active": (C.c_int, [vp, vp, C.c_uint32, C.c_int]),
    "bpomp_vertices_count": (C.c_int, [vp, c_u32p]),
    "bpomp_vertices_clear": (C.c_int, [vp]),
    "bpomp_neighbors_radius": (C.c_int, [vp, vp, C.c_uint32, C.c_double, c_u64p]),
    "bpomp_neighbors_fetch": (C.c_int, [vp, vp, vp, vp]),
    "bpomp_neighbors_device": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
    "bpomp_neighbors_all": (C.c_int, [vp, C.c_double, c_u64p]),
    "bpomp_edges_check": (C.c_int, [vp, vp, vp, C.c_int64, vp, vp, vp]),
    "bpomp_edges_check_indexed": (C.c_int, [vp, vp, vp, C.c_int64, vp, vp, vp]),
    "bpomp_edges_check_dev": (C.c_int, [vp, vp, vp, C.c_int64, vp, vp, vp]),
    "bpomp_edges_check_indexed_dev": (C.c_int, [vp, vp, vp, C.c_int64, vp, vp, vp]),
    "bpomp_edge_status_pack_dev": (C.c_int, [vp, vp, C.c_int64, vp]),
    "bpomp_edge_status_unpack_dev": (C.c_int, [vp, vp, C.c_int64, vp]),
    "bpomp_edge_states": (C.c_int, [vp, vp, vp, vp, C.c_uint32, c_u32p]),
    "bpomp_belief_config": (C.c_int, [vp, C.c_int, C.c_double, C.c_double]),
    "bpomp_belief_append": (C.c_int, [vp, vp, vp, C.c_int64]),
    "bpomp_belief_append_checked": (C.c_int, [vp, vp, vp, vp, C.c_int64]),
    "bpomp_belief_append_checked_indexed": (C.c_int, [vp, vp, vp, vp, vp, C.c_int64]),
    "bpomp_belief_size": (C.c_int, [vp, c_u64p]),
    "bpomp_belief_clear": (C.c_int, [vp]),
    "bpomp_belief_estimate": (C.c_int, [vp, vp, C.c_int64, vp]),
    "bpomp_belief_edge_pfree": (C.c_int, [vp, vp, vp, C.c_int64, vp, vp]),
    "bpomp_belief_edge_pfree_indexed": (C.c_int, [vp, vp, vp, C.c_int64, vp, vp]),
    "bpomp_belief_estimate_dev": (C.c_int, [vp, vp, C.c_int64, vp]),
    "bpomp_shard_range": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "bpomp_comm_unique_id": (C.c_int, [vp, C.c_size_t]),
    "bpomp_comm_init_rank": (C.c_int, [vp, C.c_int, C.c_int, vp]),
    "bpomp_comm_init": (C.c_int, [C.POINTER(vp), C.c_int]),
    "bpomp_comm_destroy": (C.c_int, [vp]),
    "bpomp_comm_rank": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "bpomp_edges_check_sharded": (C.c_int, [vp, vp, vp, C.c_int64, vp, vp, vp]),
    "bpomp_edges_check_sharded_dev": (C.c_int, [vp, vp, vp, C.c_int64, vp, vp, vp, vp, vp]),
    "bpomp_comm_wait": (C.c_int, [vp]),
    "bpomp_neighbors_radius_sharded": (C.c_int, [vp, vp, C.c_uint32, C.c_double, c_u64p]),
    "bpomp_belief_estimate_sharded": (C.c_int, [vp, vp, C.c_int64, vp]),
}

_lib = None


class BpompError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("bpomp error %d: %s" % (code, msg))
        self.code = code


def load():
    """Loads libbpomp_cuda.so and binds every declared symbol.  Raises if the library is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


class Context:
    """One bpomp_ctx (one device).  Thin, one method per C entry point."""

    def __init__(self, dim, segment_length, device=0):
        self.lib = load()
        self.dim = int(dim)
        h = vp()
        rc = self.lib.bpomp_create(int(device), int(dim), float(segment_length), C.byref(h))
        if rc != 0:
            raise BpompError(rc, self.lib.bpomp_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.bpomp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise BpompError(rc, self.lib.bpomp_last_error(self.h).decode())

    # ---- context -------------------------------------------------------------------------
    def sync(self):
        self._ck(self.lib.bpomp_sync(self.h))

    def set_stream(self, stream_ptr):
        self._ck(self.lib.bpomp_set_stream(self.h, C.c_void_p(stream_ptr) if stream_ptr else None))

    @property
    def launches(self):
        return int(self.lib.bpomp_launch_count(self.h))

    def perm_reserve(self, nmax):
        self._ck(self.lib.bpomp_perm_reserve(self.h, int(nmax)))

    # ---- world ---------------------------------------------------------------------------
    def set_world_image(self, pix, threshold=128):
        pix = np.ascontiguousarray(pix, dtype=np.uint8)
        self._ck(self.lib.bpomp_world_set_image(self.h, _ptr(pix), pix.shape[1], pix.shape[0], int(threshold)))

    def set_world_boxes(self, lo, hi):
        lo, hi = _f64(lo).reshape(-1, self.dim), _f64(hi).reshape(-1, self.dim)
        self._ck(self.lib.bpomp_world_set_boxes(self.h, _ptr(lo), _ptr(hi), lo.shape[0]))

    def set_bounds(self, lo, hi):
        """State-space bounds hint (scalars broadcast to dim); None, None removes it."""
        if lo is None:
            self._ck(self.lib.bpomp_world_set_bounds(self.h, None, None))
            return
        lo = _f64(np.broadcast_to(np.asarray(lo, dtype=np.float64), (self.dim,)))
        hi = _f64(np.broadcast_to(np.asarray(hi, dtype=np.float64), (self.dim,)))
        self._ck(self.lib.bpomp_world_set_bounds(self.h, _ptr(lo), _ptr(hi)))

    def set_option(self, name, value):
        self._ck(self.lib.bpomp_set_option(self.h, name.encode(), int(value)))

    def counter(self, name):
        v = C.c_int64(0)
        self._ck(self.lib.bpomp_get_counter(self.h, name.encode(), C.byref(v)))
        return int(v.value)

    def states_valid(self, states):
        s = _f64(states).reshape(-1, self.dim)
        out = np.empty(s.shape[0], dtype=np.uint8)
        self._ck(self.lib.bpomp_states_valid(self.h, _ptr(s), s.shape[0], _ptr(out)))
        return out

    def states_prune(self, states, start, goal, best_cost, check_validity=True):
        s = _f64(states).reshape(-1, self.dim)
        st, go = _f64(start), _f64(goal)
        out = np.empty(s.shape[0], dtype=np.uint8)
        self._ck(self.lib.bpomp_states_prune(self.h, _ptr(s), s.shape[0], _ptr(st), _ptr(go), float(best_cost),
                                             1 if check_validity else 0, _ptr(out)))
        return out

    def vertices_prune(self, start_id, goal_id, best_cost, check_validity=False):
        out = np.empty(self.vertices_count(), dtype=np.uint8)
        self._ck(self.lib.bpomp_vertices_prune(self.h, int(start_id), int(goal_id), float(best_cost),
                                               1 if check_validity else 0, _ptr(out)))
        return out

    # ---- vertices / neighbours -----------------------------------------------------------
    def roadmap_generate(self, kind, first, count):
        """Roadmap states generated on the device: kind "halton" (first = 1-based index of the first point) or
        "uniform" (first = splitmix64 seed); bit-equal to workloads.halton / workloads.uniform."""
        out = np.empty((int(count), self.dim), dtype=np.float64)
        k = {"halton": 0, "uniform": 1}[kind]
        self._ck(self.lib.bpomp_roadmap_generate(self.h, k, int(first), int(count), _ptr(out)))
        return out

    def vertices_append(self, states):
        s = _f64(states).reshape(-1, self.dim)
        first = C.c_uint32(0)
        self._ck(self.lib.bpomp_vertices_append(self.h, _ptr(s), s.shape[0], C.byref(first)))
        return int(first.value)

    def vertices_set_active(self, ids, flag):
        ids = _u32(ids)
        self._ck(self.lib.bpomp_vertices_set_active(self.h, _ptr(ids), ids.size, 1 if flag else 0))

    def vertices_count(self):
        n = C.c_uint32(0)
        self._ck(self.lib.bpomp_vertices_count(self.h, C.byref(n)))
        return int(n.value)

    def vertices_clear(self):
        self._ck(self.lib.bpomp_vertices_clear(self.h))

    def _fetch(self, nq, nnz):
        row_ptr = np.empty(nq + 1, dtype=np.uint64)
        col = np.empty(max(nnz, 1), dtype=np.uint32)
        dist = np.empty(max(nnz, 1), dtype=np.float64)
        self._ck(self.lib.bpomp_neighbors_fetch(self.h, _ptr(row_ptr), _ptr(col), _ptr(dist)))
        return row_ptr, col[:nnz], dist[:nnz]

    def neighbors_radius(self, query_ids, r, fetch=True):
        q = _u32(query_ids)
        nnz = C.c_uint64(0)
        self._ck(self.lib.bpomp_neighbors_radius(self.h, _ptr(q), q.size, float(r), C.byref(nnz)))
        if not fetch:  # the CSR stays on the device (bpomp_neighbors_device / a later fetch)
            return int(nnz.value)
        return self._fetch(q.size, int(nnz.value))

    def neighbors_all(self, r, fetch=True):
        nnz = C.c_uint64(0)
        self._ck(self.lib.bpomp_neighbors_all(self.h, float(r), C.byref(nnz)))
        if not fetch:
            return int(nnz.value)
        return self._fetch(self.vertices_count(), int(nnz.value))

    # ---- edges ---------------------------------------------------------------------------
    def edges_check(self, frm, to):
        f, t = _f64(frm).reshape(-1, self.dim), _f64(to).reshape(-1, self.dim)
        n = f.shape[0]
        valid = np.empty(n, dtype=np.uint8)
        first = np.empty(n, dtype=np.int32)
        ns = np.empty(n, dtype=np.uint32)
        self._ck(self.lib.bpomp_edges_check(self.h, _ptr(f), _ptr(t), n, _ptr(valid), _ptr(first), _ptr(ns)))
        return valid, first, ns

    def edges_check_into(self, frm, to, valid, first, ns=None):
        """Host-buffer call with caller-provided (e.g. pinned) numpy arrays; no allocation."""
        self._ck(self.lib.bpomp_edges_check(self.h, _ptr(frm), _ptr(to), frm.shape[0], _ptr(valid), _ptr(first),
                                            _ptr(ns)))

    def edges_check_indexed(self, from_ids, to_ids):
        f, t = _u32(from_ids), _u32(to_ids)
        n = f.size
        valid = np.empty(n, dtype=np.uint8)
        first = np.empty(n, dtype=np.int32)
        ns = np.empty(n, dtype=np.uint32)
        self._ck(self.lib.bpomp_edges_check_indexed(self.h, _ptr(f), _ptr(t), n, _ptr(valid), _ptr(first), _ptr(ns)))
        return valid, first, ns

    def edges_check_dev(self, d_from, d_to, count, d_valid, d_first, d_nstates=0):
        """Device pointers (ints, e.g. torch tensor.data_ptr()); asynchronous on the ctx stream."""
        self._ck(self.lib.bpomp_edges_check_dev(self.h, _ptr(d_from), _ptr(d_to), int(count), _ptr(d_valid),
                                                _ptr(d_first), _ptr(d_nstates) if d_nstates else None))

    def edges_check_indexed_dev(self, d_from_ids, d_to_ids, count, d_valid, d_first, d_nstates=0):
        self._ck(self.lib.bpomp_edges_check_indexed_dev(self.h, _ptr(d_from_ids), _ptr(d_to_ids), int(count),
                                                        _ptr(d_valid), _ptr(d_first),
                                                        _ptr(d_nstates) if d_nstates else None))

    def edge_status_pack_dev(self, d_valid, count, d_packed):
        """Status bytes -> 2 bits per edge ((count+15)//16 uint32 words); device pointers, ctx stream."""
        self._ck(self.lib.bpomp_edge_status_pack_dev(self.h, _ptr(d_valid), int(count), _ptr(d_packed)))

    def edge_status_unpack_dev(self, d_packed, count, d_valid):
        self._ck(self.lib.bpomp_edge_status_unpack_dev(self.h, _ptr(d_packed), int(count), _ptr(d_valid)))

    def edge_states(self, frm, to):
        f, t = _f64(frm), _f64(to)
        n = C.c_uint32(0)
        self._ck(self.lib.bpomp_edge_states(self.h, _ptr(f), _ptr(t), None, 0, C.byref(n)))
        out = np.empty((n.value, self.dim), dtype=np.float64)
        self._ck(self.lib.bpomp_edge_states(self.h, _ptr(f), _ptr(t), _ptr(out), n.value, C.byref(n)))
        return out

    # ---- multi-GPU ----------------------------------------------------------------------
    def comm_init_rank(self, nranks, rank, unique_id):
        buf = C.create_string_buffer(unique_id, 128)
        self._ck(self.lib.bpomp_comm_init_rank(self.h, int(nranks), int(rank), buf))

    def comm_rank(self):
        r, n = C.c_int(0), C.c_int(1)
        self._ck(self.lib.bpomp_comm_rank(self.h, C.byref(r), C.byref(n)))
        return int(r.value), int(n.value)

    def comm_wait(self):
        self._ck(self.lib.bpomp_comm_wait(self.h))

    def edges_check_sharded(self, frm, to):
        f, t = _f64(frm).reshape(-1, self.dim), _f64(to).reshape(-1, self.dim)
        n = f.shape[0]
        valid = np.empty(n, dtype=np.uint8)
        first = np.empty(n, dtype=np.int32)
        ns = np.empty(n, dtype=np.uint32)
        self._ck(self.lib.bpomp_edges_check_sharded(self.h, _ptr(f), _ptr(t), n, _ptr(valid), _ptr(first), _ptr(ns)))
        return valid, first, ns

    def edges_check_sharded_dev(self, d_from, d_to, shard_count, d_valid, d_first, d_nstates, d_packed, d_packed_all):
        self._ck(self.lib.bpomp_edges_check_sharded_dev(self.h, _ptr(d_from), _ptr(d_to), int(shard_count), _ptr(d_valid),
                                                        _ptr(d_first), _ptr(d_nstates) if d_nstates else None,
                                                        _ptr(d_packed), _ptr(d_packed_all)))

    def neighbors_radius_sharded(self, query_ids, r):
        q = _u32(query_ids)
        nnz = C.c_uint64(0)
        self._ck(self.lib.bpomp_neighbors_radius_sharded(self.h, _ptr(q), q.size, float(r), C.byref(nnz)))
        return self._fetch(q.size, int(nnz.value))

    def belief_estimate_sharded(self, queries):
        q = _f64(queries).reshape(-1, self.dim)
        out = np.empty(q.shape[0], dtype=np.float64)
        self._ck(self.lib.bpomp_belief_estimate_sharded(self.h, _ptr(q), q.shape[0], _ptr(out)))
        return out

    # ---- belief --------------------------------------------------------------------------
    def belief_config(self, k=15, prior=0.5, prior_weight=0.25):
        self._ck(self.lib.bpomp_belief_config(self.h, int(k), float(prior), float(prior_weight)))

    def belief_append(self, states, values):
        s = _f64(states).reshape(-1, self.dim)
        v = _f64(values)
        self._ck(self.lib.bpomp_belief_append(self.h, _ptr(s), _ptr(v), s.shape[0]))

    def belief_append_checked(self, frm, to, first_invalid):
        f, t = _f64(frm).reshape(-1, self.dim), _f64(to).reshape(-1, self.dim)
        fi = np.ascontiguousarray(first_invalid, dtype=np.int32)
        self._ck(self.lib.bpomp_belief_append_checked(self.h, _ptr(f), _ptr(t), _ptr(fi), f.shape[0]))

    def belief_append_checked_indexed(self, from_ids, to_ids, first_invalid, nstates):
        f, t = _u32(from_ids), _u32(to_ids)
        fi = np.ascontiguousarray(first_invalid, dtype=np.int32)
        ns = _u32(nstates)
        self._ck(self.lib.bpomp_belief_append_checked_indexed(self.h, _ptr(f), _ptr(t), _ptr(fi), _ptr(ns), f.size))

    def belief_size(self):
        n = C.c_uint64(0)
        self._ck(self.lib.bpomp_belief_size(self.h, C.byref(n)))
        return int(n.value)

    def belief_clear(self):
        self._ck(self.lib.bpomp_belief_clear(self.h))

    def belief_estimate(self, queries):
        q = _f64(queries).reshape(-1, self.dim)
        out = np.empty(q.shape[0], dtype=np.float64)
        self._ck(self.lib.bpomp_belief_estimate(self.h, _ptr(q), q.shape[0], _ptr(out)))
        return out

    def belief_edge_pfree(self, frm, to):
        f, t = _f64(frm).reshape(-1, self.dim), _f64(to).reshape(-1, self.dim)
        n = f.shape[0]
        pf = np.empty(n, dtype=np.float64)
        nl = np.empty(n, dtype=np.uint32)
        self._ck(self.lib.bpomp_belief_edge_pfree(self.h, _ptr(f), _ptr(t), n, _ptr(pf), _ptr(nl)))
        return pf, nl

    def belief_edge_pfree_indexed(self, from_ids, to_ids):
        f, t = _u32(from_ids), _u32(to_ids)
        n = f.size
        pf = np.empty(n, dtype=np.float64)
        nl = np.empty(n, dtype=np.uint32)
        self._ck(self.lib.bpomp_belief_edge_pfree_indexed(self.h, _ptr(f), _ptr(t), n, _ptr(pf), _ptr(nl)))
        return pf, nl


def shard_range(count, nranks, rank):
    b, e = C.c_int64(0), C.c_int64(0)
    rc = load().bpomp_shard_range(int(count), int(nranks), int(rank), C.byref(b), C.byref(e))
    if rc != 0:
        raise BpompError(rc, "bpomp_shard_range")
    return int(b.value), int(e.value)


def comm_unique_id():
    """128-byte NCCL id made by rank 0 (pass it to every rank's Context.comm_init_rank)."""
    buf = (C.c_char * 128)()
    rc = load().bpomp_comm_unique_id(buf, 128)
    if rc != 0:
        raise BpompError(rc, load().bpomp_last_error(None).decode())
    return bytes(buf)


def comm_init_all(contexts):
    """One process driving several contexts (contexts[i] becomes rank i)."""
    arr = (vp * len(contexts))(*[c.h for c in contexts])
    rc = load().bpomp_comm_init(arr, len(contexts))
    if rc != 0:
        raise BpompError(rc, load().bpomp_last_error(contexts[0].h).decode())


def bisect_perm(n):
    out = np.empty(max(n, 1), dtype=np.int32)
    rc = load().bpomp_bisect_perm(int(n), out.ctypes.data_as(c_i32p))
    if rc != 0:
        raise BpompError(rc, "bpomp_bisect_perm")
    return out[:n]
