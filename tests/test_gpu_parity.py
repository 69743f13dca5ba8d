"""GPU parity tests proper: the sm_100a kernels, called through the C ABI, against the CPU oracle on
the same seeded inputs.  Bit-exact for validity flags, first-invalid slots, nStates, neighbour rows and
distances, prune masks; belief values bit-equal to the oracle (1e-12 relative is the bar against the
reference's Eigen summation order)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from batching_pomp_b200 import capi as c
    c.load()
    return c


@pytest.fixture(scope="module")
def wl():
    from batching_pomp_b200 import workloads
    return workloads


def _ctx_boxes(capi, wl, d, nboxes, seed, fraction=0.01):
    L = wl.segment_length(d, fraction)
    lo, hi = wl.box_world(d, nboxes, seed)
    ctx = capi.Context(d, L)
    ctx.set_world_boxes(lo, hi)
    return ctx, L, lo, hi


@pytest.mark.parametrize("d,nboxes,count", [(7, 500, 200000), (2, 40, 50000), (3, 100, 50000), (6, 300, 50000),
                                            (8, 200, 30000), (1, 5, 5000), (7, 33, 20000), (7, 1, 20000),
                                            # tables too large for shared memory -> global-memory path
                                            (8, 500, 30000), (7, 1200, 30000), (4, 700, 30000)])
def test_edge_check_boxes_parity(capi, orc, wl, d, nboxes, count):
    ctx, L, lo, hi = _ctx_boxes(capi, wl, d, nboxes, seed=3)
    frm, to = wl.random_edges(count, d, seed=4, max_len=wl.halton_radius(10 ** 6, 7))
    v, f, n = ctx.edges_check(frm, to)
    ov, of, on = orc.edges_check_boxes(frm, to, L, lo, hi, nthreads=orc.max_threads())
    np.testing.assert_array_equal(n, on)
    np.testing.assert_array_equal(v, ov)
    np.testing.assert_array_equal(f, of)
    assert 0 < int((v == 0).sum()) < count or nboxes <= 1
    ctx.close()


def test_edge_check_long_edges_and_deferred_orders(capi, orc, wl):
    """Edges across the whole cube (n up to ~100 at the default resolution, and up to ~2600 at a
    finer one) exercise sample orders beyond the 256 rows resident at create."""
    d = 7
    lo, hi = wl.box_world(d, 500, 3)
    frm = wl.uniform(11, (20000, d))
    to = wl.uniform(12, (20000, d))
    for fraction in (0.01, 0.0005):
        L = wl.segment_length(d, fraction)
        ctx = capi.Context(d, L)
        ctx.set_world_boxes(lo, hi)
        m = 20000 if fraction == 0.01 else 3000
        v, f, n = ctx.edges_check(frm[:m], to[:m])
        ov, of, on = orc.edges_check_boxes(frm[:m], to[:m], L, lo, hi, nthreads=orc.max_threads())
        np.testing.assert_array_equal(n, on)
        np.testing.assert_array_equal(v, ov)
        np.testing.assert_array_equal(f, of)
        if fraction != 0.01:
            assert n.max() > 256
        ctx.close()


def test_edge_check_pipelined_host_path_with_deferred_orders(capi, orc, wl):
    """> 2M edges take the chunked three-stream host path; long edges at a fine resolution need sample
    orders that are not resident, which makes that path fall back and upload them."""
    pix = wl.image_world(512, seed=1)
    L = wl.segment_length(2, 0.002)
    ctx = capi.Context(2, L)
    ctx.set_world_image(pix)
    count = 2200000
    frm = wl.uniform(61, (count, 2))
    to = wl.uniform(62, (count, 2))
    v, f, n = ctx.edges_check(frm, to)
    assert n.max() > 256
    ov, of, on = orc.edges_check_image(frm, to, L, pix, nthreads=orc.max_threads())
    np.testing.assert_array_equal(n, on)
    np.testing.assert_array_equal(v, ov)
    np.testing.assert_array_equal(f, of)
    # second call: every order is resident now, the pipelined path completes on its own
    v2, f2, n2 = ctx.edges_check(frm, to)
    np.testing.assert_array_equal(f2, of)
    ctx.close()


def test_edge_check_dense_world_candidate_overflow(capi, orc, wl):
    """Many overlapping boxes around every edge: more than EC_MAXCAND candidates -> brute path."""
    d = 3
    rng = np.random.default_rng(5)
    c = rng.random((400, d)) * 0.2 + 0.4
    h = rng.random((400, d)) * 0.02 + 0.001
    lo, hi = c - h, c + h
    L = wl.segment_length(d, 0.005)
    ctx = capi.Context(d, L)
    ctx.set_world_boxes(lo, hi)
    frm = rng.random((20000, d)) * 0.3 + 0.35
    to = rng.random((20000, d)) * 0.3 + 0.35
    v, f, n = ctx.edges_check(frm, to)
    ov, of, on = orc.edges_check_boxes(frm, to, L, lo, hi, nthreads=orc.max_threads())
    np.testing.assert_array_equal(n, on)
    np.testing.assert_array_equal(v, ov)
    np.testing.assert_array_equal(f, of)
    ctx.close()


def test_edge_check_edge_cases(capi, orc, wl):
    d = 7
    ctx, L, lo, hi = _ctx_boxes(capi, wl, d, 500, seed=3)
    # empty batch
    v, f, n = ctx.edges_check(np.zeros((0, d)), np.zeros((0, d)))
    assert v.size == 0
    # zero-length and sub-resolution edges: n = 2, FREE unchecked even inside an obstacle
    inside = (lo[0] + hi[0]) / 2
    frm = np.stack([inside, inside, np.zeros(d)])
    to = np.stack([inside, inside + L / 10, np.full(d, L)])
    v, f, n = ctx.edges_check(frm, to)
    ov, of, on = orc.edges_check_boxes(frm, to, L, lo, hi)
    assert n.tolist() == on.tolist() and n[0] == 2 and n[1] == 2
    np.testing.assert_array_equal(v, ov)
    np.testing.assert_array_equal(f, of)
    # orientation matters: reversed edges are evaluated independently and must match the oracle too
    a, b = wl.random_edges(30000, d, seed=9, max_len=0.6)
    v1, f1, n1 = ctx.edges_check(b, a)
    ov1, of1, on1 = orc.edges_check_boxes(b, a, L, lo, hi, nthreads=orc.max_threads())
    np.testing.assert_array_equal(v1, ov1)
    np.testing.assert_array_equal(f1, of1)
    # no world with zero boxes: everything FREE
    ctx.set_world_boxes(np.zeros((0, d)), np.zeros((0, d)))
    v, f, n = ctx.edges_check(a[:100], b[:100])
    assert v.tolist() == [1] * 100 and f.tolist() == [-1] * 100
    ctx.close()


def test_edge_check_too_many_states_is_a_range_error(capi, wl):
    """nStates above BPOMP_NSTATES_MAX (65535): the call fails with BPOMP_ERR_RANGE instead of truncating."""
    ctx = capi.Context(2, 1e-6)
    ctx.set_world_boxes(np.array([[0.4, 0.4]]), np.array([[0.6, 0.6]]))
    with pytest.raises(capi.BpompError) as ei:
        ctx.edges_check(np.array([[0.0, 0.0]]), np.array([[1.0, 1.0]]))
    assert ei.value.code == -4
    # the context stays usable
    v, f, n = ctx.edges_check(np.array([[0.0, 0.0]]), np.array([[0.01, 0.01]]))
    assert n[0] == 14142 and v[0] == 1
    with pytest.raises(capi.BpompError):
        ctx.belief_edge_pfree(np.array([[0.0, 0.0]]), np.array([[1.0, 1.0]]))
    ctx.close()


def test_empty_and_degenerate_calls(capi, wl):
    ctx = capi.Context(3, 0.01)
    ctx.set_world_boxes(np.zeros((0, 3)), np.zeros((0, 3)))
    assert ctx.states_valid(np.zeros((0, 3))).size == 0
    assert ctx.belief_estimate(np.zeros((0, 3))).size == 0
    pf, nl = ctx.belief_edge_pfree(np.zeros((0, 3)), np.zeros((0, 3)))
    assert pf.size == 0
    rp, col, dist = ctx.neighbors_radius(np.zeros(0, dtype=np.uint32), 0.1)
    assert rp.tolist() == [0] and col.size == 0
    assert ctx.neighbors_all(0.1, fetch=False) == 0
    ctx.vertices_append(wl.halton(10, 3))
    ctx.vertices_set_active(np.arange(10, dtype=np.uint32), 0)  # everything deactivated
    rp, col, dist = ctx.neighbors_radius(np.array([3], dtype=np.uint32), 10.0)
    assert rp.tolist() == [0, 0]
    with pytest.raises(capi.BpompError):
        ctx.neighbors_radius(np.array([10], dtype=np.uint32), 1.0)  # id out of range
    ctx.belief_config(15, 0.5, 0.25)
    ctx.belief_append(wl.halton(3, 3), np.array([1.0, 0.0, 1.0]))  # fewer points than k
    import ctypes
    got = ctx.belief_estimate(wl.halton(5, 3, first_index=20))
    assert np.all((got > 0) & (got < 1))
    ctx.close()


def test_edge_check_no_world_is_an_error(capi):
    ctx = capi.Context(2, 0.01)
    with pytest.raises(capi.BpompError) as ei:
        ctx.edges_check(np.zeros((1, 2)), np.ones((1, 2)))
    assert ei.value.code == -3
    ctx.close()


@pytest.mark.parametrize("size,fraction,count", [(1024, 1e-3, 200000), (4096, 1e-4, 200000), (37, 1e-3, 20000)])
def test_edge_check_image_parity(capi, orc, wl, size, fraction, count):
    d = 2
    pix = wl.image_world(size, seed=1) if size >= 64 else (np.random.default_rng(1).random((size, size + 5)) * 255).astype(np.uint8)
    L = wl.segment_length(d, fraction)
    ctx = capi.Context(d, L)
    ctx.set_world_image(pix)
    r = wl.halton_radius(10 ** 4 if fraction == 1e-3 else 10 ** 6, 2)
    frm, to = wl.random_edges(count, d, seed=4, max_len=2 * r)
    v, f, n = ctx.edges_check(frm, to)
    ov, of, on = orc.edges_check_image(frm, to, L, pix, nthreads=orc.max_threads())
    np.testing.assert_array_equal(n, on)
    np.testing.assert_array_equal(v, ov)
    np.testing.assert_array_equal(f, of)
    assert 0 < int((v == 0).sum()) < count
    ctx.close()


def test_states_valid_and_prune_parity(capi, orc, wl):
    d = 7
    ctx, L, lo, hi = _ctx_boxes(capi, wl, d, 500, seed=3)
    st = wl.uniform(21, (100000, d)) * 1.2 - 0.1  # some states outside the unit cube
    np.testing.assert_array_equal(ctx.states_valid(st), orc.states_valid_boxes(st, lo, hi))
    # points exactly on box faces are inside (closed boxes)
    on_face = np.concatenate([lo[:50], hi[:50]])
    assert ctx.states_valid(on_face).tolist() == [0] * 100
    start, goal = np.full(d, 0.1), np.full(d, 0.9)
    for best in (np.finfo(np.float64).max, 2.2, 2.1166, 3.0):
        got = ctx.states_prune(st, start, goal, best, True)
        inad = orc.vertices_inadmissible(st, start, goal, best)
        want = (inad | (orc.states_valid_boxes(st, lo, hi) == 0)).astype(np.uint8)
        np.testing.assert_array_equal(got, want)
        np.testing.assert_array_equal(ctx.states_prune(st, start, goal, best, False), inad)
    ctx.close()
    # image world
    pix = wl.image_world(512, seed=1)
    ctx = capi.Context(2, 0.001)
    ctx.set_world_image(pix)
    st = wl.uniform(22, (100000, 2)) * 1.02 - 0.01
    st[:4] = [[0.0, 0.0], [1.0, 1.0], [1.0, 0.0], [0.0, 1.0]]
    np.testing.assert_array_equal(ctx.states_valid(st), orc.states_valid_image(st, pix))
    ctx.close()


@pytest.mark.parametrize("d,N,nq,rfun", [(2, 20000, 20000, "halton"), (7, 20000, 300, "halton"), (3, 5000, 5000, "halton"),
                                         (2, 3000, 50, "max"), (7, 2000, 20, "max"), (2, 1, 1, "halton"),
                                         # many queries in the brute-force regime -> tiled all-pairs kernel
                                         (7, 20000, 2000, "halton"), (2, 3000, 1000, "max"), (5, 4000, 4000, "max"),
                                         # finite radius, dim >= 4: the packed-FP32 prefilter form of that kernel
                                         (4, 30000, 3000, "halton"), (8, 5000, 1500, "halton"), (6, 777, 777, "halton")])
def test_neighbors_parity(capi, orc, wl, d, N, nq, rfun):
    verts = wl.halton(N, d)
    r = wl.halton_radius(N, d) if rfun == "halton" else np.finfo(np.float64).max
    ctx = capi.Context(d, wl.segment_length(d))
    assert ctx.vertices_append(verts[: N // 2]) == 0
    assert ctx.vertices_append(verts[N // 2:]) == N // 2
    rng = np.random.default_rng(0)
    q = rng.choice(N, size=nq, replace=False).astype(np.uint32) if nq < N else np.arange(N, dtype=np.uint32)
    rp, col, dist = ctx.neighbors_radius(q, r)
    orp, ocol, odist = orc.neighbors_radius(verts, q, r, nthreads=orc.max_threads())
    np.testing.assert_array_equal(rp, orp)
    np.testing.assert_array_equal(col, ocol)
    np.testing.assert_array_equal(dist, odist)
    # deactivate a third of the vertices (BatchingManager::pruneVertices -> NN.remove)
    off = rng.choice(N, size=N // 3, replace=False).astype(np.uint32)
    ctx.vertices_set_active(off, 0)
    active = np.ones(N, dtype=np.uint8)
    active[off] = 0
    rp, col, dist = ctx.neighbors_radius(q, r)
    orp, ocol, odist = orc.neighbors_radius(verts, q, r, active=active, nthreads=orc.max_threads())
    np.testing.assert_array_equal(rp, orp)
    np.testing.assert_array_equal(col, ocol)
    np.testing.assert_array_equal(dist, odist)
    ctx.close()


def test_neighbors_huge_rows_global_sort(capi, orc, wl):
    """Rows longer than the shared-memory sort handles (vertex batching's r = DBL_MAX on many vertices)."""
    d, N = 2, 9000
    verts = wl.halton(N, d)
    ctx = capi.Context(d, wl.segment_length(d))
    ctx.vertices_append(verts)
    q = np.array([0, 4321, 8999], dtype=np.uint32)
    rp, col, dist = ctx.neighbors_radius(q, np.finfo(np.float64).max)
    orp, ocol, odist = orc.neighbors_radius(verts, q, np.finfo(np.float64).max, nthreads=orc.max_threads())
    np.testing.assert_array_equal(rp, orp)
    np.testing.assert_array_equal(col, ocol)
    np.testing.assert_array_equal(dist, odist)
    ctx.close()


def test_neighbors_all_matches_per_query(capi, orc, wl):
    d, N = 2, 30000
    verts = wl.halton(N, d)
    r = wl.halton_radius(N, d)
    ctx = capi.Context(d, wl.segment_length(d))
    ctx.vertices_append(verts)
    off = np.arange(0, N, 7, dtype=np.uint32)
    ctx.vertices_set_active(off, 0)
    rp, col, dist = ctx.neighbors_all(r)
    active = np.ones(N, dtype=np.uint8)
    active[off] = 0
    orp, ocol, odist = orc.neighbors_radius(verts, np.arange(N), r, active=active, nthreads=orc.max_threads())
    # inactive query vertices get empty rows in the whole-roadmap call
    lens = np.diff(orp.astype(np.int64))
    lens[off] = 0
    want_rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    np.testing.assert_array_equal(rp, want_rp)
    keep = np.repeat(active.astype(bool), np.diff(orp.astype(np.int64)))
    np.testing.assert_array_equal(col, ocol[keep])
    np.testing.assert_array_equal(dist, odist[keep])
    # symmetry: u in row(v) <=> v in row(u)
    rows = np.repeat(np.arange(N), np.diff(rp.astype(np.int64)))
    a = set(zip(rows.tolist(), col.tolist()))
    assert all((c, r_) in a for r_, c in list(a)[:5000])
    ctx.close()


def test_neighbors_ties_and_duplicates(capi, orc):
    verts = np.array([[0.0, 0.0], [0.3, 0.0], [0.0, 0.3], [0.5, 0.0], [0.3, 0.0], [0.0, 0.0]])
    ctx = capi.Context(2, 0.01)
    ctx.vertices_append(verts)
    rp, col, dist = ctx.neighbors_radius([0, 3], 0.3)
    orp, ocol, odist = orc.neighbors_radius(verts, [0, 3], 0.3)
    np.testing.assert_array_equal(rp, orp)
    np.testing.assert_array_equal(col, ocol)
    np.testing.assert_array_equal(dist, odist)
    assert col[:rp[1]].tolist() == [0, 5, 1, 2, 4]
    ctx.close()


@pytest.mark.parametrize("d,M,nq,K", [(7, 5000, 3000, 15), (2, 300, 1000, 15), (7, 7, 100, 15), (3, 40000, 500, 15),
                                      (7, 2000, 200, 1), (5, 1000, 300, 32)])
def test_belief_estimate_parity(capi, orc, wl, d, M, nq, K):
    pts = wl.uniform(31, (M, d))
    vals = (wl.uniform(32, (M,)) < 0.3).astype(np.float64)
    q = wl.uniform(33, (nq, d))
    q[: min(nq, M) // 2] = pts[: min(nq, M) // 2]  # exact hits -> the 1e-4 distance clamp
    ctx = capi.Context(d, wl.segment_length(d))
    ctx.belief_config(K, 0.5, 0.25)
    assert np.all(ctx.belief_estimate(q[:10]) == 0.5)  # empty model -> prior
    ctx.belief_append(pts[: M // 2], vals[: M // 2])
    ctx.belief_append(pts[M // 2:], vals[M // 2:])
    assert ctx.belief_size() == M
    got = ctx.belief_estimate(q)
    want = orc.belief_estimate(pts, vals, q, K=K, nthreads=orc.max_threads())
    np.testing.assert_array_equal(got, want)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)
    ctx.close()


def test_belief_duplicate_points_tie_break(capi, orc, wl):
    d = 2
    base = wl.uniform(41, (10, d))
    pts = np.concatenate([base, base, base])  # every distance appears three times
    vals = np.concatenate([np.zeros(10), np.ones(10), np.zeros(10)])
    q = wl.uniform(42, (200, d))
    ctx = capi.Context(d, 0.01)
    ctx.belief_config(4, 0.5, 0.25)
    ctx.belief_append(pts, vals)
    np.testing.assert_array_equal(ctx.belief_estimate(q), orc.belief_estimate(pts, vals, q, K=4))
    ctx.close()


def test_belief_edge_pfree_and_append_checked(capi, orc, wl):
    d = 7
    ctx, L, lo, hi = _ctx_boxes(capi, wl, d, 500, seed=3)
    frm, to = wl.random_edges(3000, d, seed=6, max_len=1.2)
    # empty model: pfree = 0.5^(#query slots)
    pf, nl = ctx.belief_edge_pfree(frm, to)
    opf, onl = orc.belief_edge_pfree(frm, to, L, np.zeros((0, d)), np.zeros(0))
    np.testing.assert_array_equal(nl, onl)
    np.testing.assert_array_equal(pf, opf)
    # replay the belief inserts of checkAndSetEdgeBlocked on the device and compare with a host replay
    v, f, n = ctx.edges_check(frm, to)
    ctx.belief_append_checked(frm, to, f)
    pts, vals = [], []
    for e in range(frm.shape[0]):
        st = orc.edge_states(frm[e], to[e], L)
        last = f[e] if f[e] >= 1 else n[e] - 2
        for i in range(1, last + 1):
            pts.append(st[i])
            vals.append(1.0 if (f[e] >= 1 and i == f[e]) else 0.0)
    pts, vals = np.array(pts), np.array(vals)
    assert ctx.belief_size() == len(vals)
    assert vals.sum() == (v == 0).sum()
    pf, nl = ctx.belief_edge_pfree(frm[:800], to[:800])
    opf, onl = orc.belief_edge_pfree(frm[:800], to[:800], L, pts, vals, nthreads=orc.max_threads())
    np.testing.assert_array_equal(nl, onl)
    np.testing.assert_array_equal(pf, opf)
    ctx.close()


@pytest.mark.parametrize("d,K,M,L", [(7, 15, 20000, None), (2, 15, 5000, 0.004), (3, 1, 3000, None), (5, 32, 9000, None),
                                     (7, 15, 9, None), (2, 15, 600, 0.0002)])
def test_belief_edge_pfree_one_launch_path(capi, orc, wl, d, K, M, L):
    """Small batches run computeAndSetEdgeFreeProbability (BP.cpp:464-493) as ONE launch; that path, the
    tiled multi-launch path and the oracle must agree bit for bit (ties, M < K, absent sample orders)."""
    L = L or wl.segment_length(d)
    base = wl.uniform(71, (M, d))
    base[M // 2:] = base[: M - M // 2]  # duplicates: equal distances, the earlier index must win
    vals = (wl.uniform(72, (M,)) < 0.4).astype(np.float64)
    frm, to = wl.random_edges(600, d, seed=73, max_len=1.0)
    z = min(50, M)
    frm[:z] = base[:z]  # zero-distance queries -> the 1e-4 clamp
    resident = np.floor(np.sqrt(((frm - to) ** 2).sum(1)) / L).max() <= 256  # sample orders uploaded at create
    want_pf, want_nl = orc.belief_edge_pfree(frm, to, L, base, vals, K=K, nthreads=orc.max_threads())
    for fused in (1, 0):
        ctx = capi.Context(d, L)
        ctx.belief_config(K, 0.5, 0.25)
        ctx.set_option("fused_pfree", fused)
        ctx.belief_append(base, vals)
        ctx.belief_estimate(frm[:1])  # dim <= 3: builds the model's spatial index (its own launches)
        n0 = ctx.launches
        pf, nl = ctx.belief_edge_pfree(frm, to)
        if fused and resident:
            assert ctx.launches - n0 == 1
        np.testing.assert_array_equal(nl, want_nl)
        np.testing.assert_array_equal(pf, want_pf)
        ctx.close()


def test_indexed_edges_reject_bad_vertex_ids(capi, orc, wl):
    """Vertex ids of a host-buffer call are validated on the device (no host loop over the ids): one bad id fails
    the call with BPOMP_ERR_RANGE — small batch and pipelined batch — and the context stays usable."""
    d, N = 7, 1000
    ctx, L, lo, hi = _ctx_boxes(capi, wl, d, 500, seed=3)
    verts = wl.halton(N, d)
    ctx.vertices_append(verts)
    rng = np.random.default_rng(3)
    for E in (5000, (1 << 21) + 12345):
        fi = rng.integers(0, N, E).astype(np.uint32)
        ti = rng.integers(0, N, E).astype(np.uint32)
        good = ctx.edges_check_indexed(fi, ti)
        for arr, pos in ((fi, E // 2), (ti, E - 1)):
            keep = arr[pos]
            arr[pos] = N  # first id beyond the table
            with pytest.raises(capi.BpompError) as ei:
                ctx.edges_check_indexed(fi, ti)
            assert "out of range" in str(ei.value)
            arr[pos] = keep
        again = ctx.edges_check_indexed(fi, ti)
        for a, b in zip(good, again):
            np.testing.assert_array_equal(a, b)
    m = 20000
    ov, of, on = orc.edges_check_boxes(verts[fi[:m]], verts[ti[:m]], L, lo, hi, nthreads=orc.max_threads())
    np.testing.assert_array_equal(again[0][:m], ov)
    np.testing.assert_array_equal(again[1][:m], of)
    ctx.close()


@pytest.mark.parametrize("d,K,M,nq", [(7, 15, 30000, 2500), (4, 32, 9000, 700), (8, 1, 20000, 1300), (5, 15, 300, 600)])
def test_fp32_prefilter_paths_equal_fp64_and_oracle(capi, orc, wl, d, K, M, nq):
    """dim >= 4: the packed-FP32 prefilter forms of the tiled kNN and neighbour kernels against their FP64-prefilter
    forms and the oracle — bit for bit; duplicates (ties by insertion index), queries on points, ragged tiles."""
    L = wl.segment_length(d)
    pts = wl.uniform(81, (M, d))
    pts[M // 2: M // 2 + M // 8] = pts[: M // 8]
    vals = (wl.uniform(82, (M,)) < 0.35).astype(np.float64)
    q = wl.uniform(83, (nq, d))
    q[:100] = pts[:100]
    want = orc.belief_estimate(pts, vals, q, K=K, nthreads=orc.max_threads())
    r = 0.9 * wl.halton_radius(M, d)
    qid = np.arange(0, M, max(1, M // 1100), dtype=np.uint32)
    rows = None
    for f32 in (1, 0):
        ctx = capi.Context(d, L)
        ctx.set_option("fp32_prefilter", f32)
        ctx.belief_config(K, 0.5, 0.25)
        ctx.belief_append(pts, vals)
        np.testing.assert_array_equal(ctx.belief_estimate(q), want)
        ctx.vertices_append(pts)
        got = ctx.neighbors_radius(qid, r)
        if rows is None:
            rows = got
            orp, ocol, odist = orc.neighbors_radius(pts, qid, r, nthreads=orc.max_threads())
            np.testing.assert_array_equal(got[0], orp)
            np.testing.assert_array_equal(got[1], ocol)
            np.testing.assert_array_equal(got[2], odist)
        else:
            for a, b in zip(rows, got):
                np.testing.assert_array_equal(a, b)
        ctx.close()


@pytest.mark.parametrize("d,n,first", [(7, 100000, 1), (2, 1000, 1), (3, 5000, 123457), (8, 300, 1)])
def test_roadmap_generated_on_device_equals_the_file_sequences(capi, orc, wl, d, n, first):
    """f3: bpomp_roadmap_generate against the two sample sequences as the roadmap files hold them (oracle Halton,
    splitmix64 uniform): bit for bit."""
    ctx = capi.Context(d, wl.segment_length(d))
    np.testing.assert_array_equal(ctx.roadmap_generate("halton", first, n), orc.halton(first, n, d))
    np.testing.assert_array_equal(ctx.roadmap_generate("halton", first, n), wl.halton(n, d, first_index=first))
    np.testing.assert_array_equal(ctx.roadmap_generate("uniform", 42 + first, n), wl.uniform(42 + first, (n, d)))
    assert ctx.roadmap_generate("uniform", 1, 0).shape == (0, d)
    ctx.close()


def _clustered_points(wl, d, M, seed):
    """Belief-like point sets: states strung along segments (checked edges), plus duplicates and a far outlier."""
    rng = np.random.default_rng(seed)
    nseg = max(1, M // 12)
    a = rng.random((nseg, d)) * 0.6 + 0.2
    b = a + (rng.random((nseg, d)) - 0.5) * 0.1
    t = rng.random((M, 1))
    k = rng.integers(0, nseg, M)
    pts = a[k] + (b[k] - a[k]) * t
    pts[M // 3: M // 3 + M // 10] = pts[: M // 10]  # duplicates: equal distances, the earlier index must win
    pts[-1] = 3.0  # far outside everything else: stretches the bounding box
    return np.ascontiguousarray(pts)


@pytest.mark.parametrize("d,K,M", [(2, 15, 6000), (2, 15, 40000), (1, 15, 3000), (3, 15, 20000), (2, 1, 5000), (3, 32, 4000),
                                   (2, 15, 600)])
def test_belief_grid_index_equals_scan_and_oracle(capi, orc, wl, d, K, M):
    """f4: the belief model's spatial index (KNNModel's GNAT, KNNModel.hpp:65, :108-111) for dim <= 3.  Estimates
    through the grid, through the plain scan and from the oracle are bit-equal: clustered points, duplicates,
    queries on points (distance 0), inside clusters, in empty regions and outside the bounding box; the model grows
    between queries (tail of unindexed points, rebuilds)."""
    L = wl.segment_length(d)
    pts = _clustered_points(wl, d, M, seed=11 + d)
    vals = (wl.uniform(12, (M,)) < 0.4).astype(np.float64)
    rng = np.random.default_rng(5)
    q = np.concatenate([pts[rng.integers(0, M, 300)],                                     # on points
                        pts[rng.integers(0, M, 600)] + (rng.random((600, d)) - 0.5) * 0.01,  # near clusters
                        rng.random((600, d)) * 1.4 - 0.2,                                   # anywhere, also outside
                        np.full((3, d), -5.0), np.full((3, d), 9.0)])
    ctxs = []
    for grid in (1, 0):
        ctx = capi.Context(d, L)
        ctx.belief_config(K, 0.5, 0.25)
        ctx.set_option("belief_grid", grid)
        ctxs.append(ctx)
    steps = [M // 2, M // 2 + 100, M // 2 + 700, M]  # first build, a short tail, a longer tail, a rebuild
    prev = 0
    for m in steps:
        for ctx in ctxs:
            ctx.belief_append(pts[prev:m], vals[prev:m])
        prev = m
        want = orc.belief_estimate(pts[:m], vals[:m], q, K=K, nthreads=orc.max_threads())
        got = [ctx.belief_estimate(q) for ctx in ctxs]
        np.testing.assert_array_equal(got[1], want)
        np.testing.assert_array_equal(got[0], want)
    if M >= 1024:
        assert ctxs[0].counter("belief_grid_builds") >= 2 and ctxs[0].counter("belief_grid_points") > 0
    assert ctxs[1].counter("belief_grid_builds") == 0
    # edge free-probabilities (the planner's call) through the index
    frm, to = wl.random_edges(500, d, seed=21, max_len=0.3)
    want_pf, want_nl = orc.belief_edge_pfree(frm, to, L, pts, vals, K=K, nthreads=orc.max_threads())
    for ctx in ctxs:
        pf, nl = ctx.belief_edge_pfree(frm, to)
        np.testing.assert_array_equal(nl, want_nl)
        np.testing.assert_array_equal(pf, want_pf)
    # a cleared model forgets its index
    ctxs[0].belief_clear()
    ctxs[0].belief_append(pts[:700], vals[:700])
    np.testing.assert_array_equal(ctxs[0].belief_estimate(q), orc.belief_estimate(pts[:700], vals[:700], q, K=K))
    for ctx in ctxs:
        ctx.close()


def test_belief_append_checked_indexed_matches_explicit(capi, orc, wl):
    d, N = 7, 4000
    ctx, L, lo, hi = _ctx_boxes(capi, wl, d, 500, seed=3)
    ctx2, _, _, _ = _ctx_boxes(capi, wl, d, 500, seed=3)
    verts = wl.halton(N, d)
    ctx.vertices_append(verts)
    rng = np.random.default_rng(5)
    fi = rng.integers(0, N, 3000).astype(np.uint32)
    ti = rng.integers(0, N, 3000).astype(np.uint32)
    ti[:10] = fi[:10]  # zero-length edges: two states, nothing to add
    v, f, n = ctx.edges_check_indexed(fi, ti)
    ctx.belief_append_checked_indexed(fi, ti, f, n)
    ctx2.belief_append_checked(verts[fi], verts[ti], f)
    assert ctx.belief_size() == ctx2.belief_size() > 0
    q = wl.uniform(9, (2000, d))
    np.testing.assert_array_equal(ctx.belief_estimate(q), ctx2.belief_estimate(q))
    fr2, to2 = wl.random_edges(500, d, seed=10, max_len=0.8)
    a, b = ctx.belief_edge_pfree(fr2, to2), ctx2.belief_edge_pfree(fr2, to2)
    np.testing.assert_array_equal(a[0], b[0])
    with pytest.raises(capi.BpompError):
        ctx.belief_append_checked_indexed(fi[:5], ti[:5], f[:5], np.ones(5, dtype=np.uint32))
    with pytest.raises(capi.BpompError):
        ctx.belief_append_checked_indexed(fi[10:12], ti[10:12], np.array([10**6, 1], dtype=np.int32), n[10:12])
    ctx.close()
    ctx2.close()


def test_indexed_variants_match_explicit(capi, orc, wl):
    d, N = 7, 5000
    ctx, L, lo, hi = _ctx_boxes(capi, wl, d, 500, seed=3)
    verts = wl.halton(N, d)
    ctx.vertices_append(verts)
    rng = np.random.default_rng(3)
    fi = rng.integers(0, N, 40000).astype(np.uint32)
    ti = rng.integers(0, N, 40000).astype(np.uint32)
    v, f, n = ctx.edges_check_indexed(fi, ti)
    ov, of, on = orc.edges_check_boxes(verts[fi], verts[ti], L, lo, hi, nthreads=orc.max_threads())
    np.testing.assert_array_equal(n, on)
    np.testing.assert_array_equal(v, ov)
    np.testing.assert_array_equal(f, of)
    ctx.belief_append(verts[:500], (np.arange(500) % 3 == 0).astype(np.float64))
    pf, nl = ctx.belief_edge_pfree_indexed(fi[:500], ti[:500])
    pf2, nl2 = ctx.belief_edge_pfree(verts[fi[:500]], verts[ti[:500]])
    np.testing.assert_array_equal(pf, pf2)
    np.testing.assert_array_equal(nl, nl2)
    with pytest.raises(capi.BpompError):
        ctx.edges_check_indexed(np.array([N], dtype=np.uint32), np.array([0], dtype=np.uint32))
    ctx.close()


def test_edge_states_match_oracle(capi, orc, wl):
    d = 7
    L = wl.segment_length(d)
    ctx = capi.Context(d, L)
    a, b = wl.uniform(51, (d,)), wl.uniform(52, (d,))
    np.testing.assert_array_equal(ctx.edge_states(a, b), orc.edge_states(a, b, L))
    ctx.close()


def test_device_pointer_api_with_torch(capi, orc, wl):
    import torch
    d = 7
    ctx, L, lo, hi = _ctx_boxes(capi, wl, d, 500, seed=3)
    frm, to = wl.random_edges(100000, d, seed=8, max_len=0.35)
    dev = torch.device("cuda:0")
    tf = torch.from_numpy(frm).to(dev)
    tt = torch.from_numpy(to).to(dev)
    tv = torch.empty(frm.shape[0], dtype=torch.uint8, device=dev)
    tfi = torch.empty(frm.shape[0], dtype=torch.int32, device=dev)
    tn = torch.empty(frm.shape[0], dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx.edges_check_dev(tf.data_ptr(), tt.data_ptr(), frm.shape[0], tv.data_ptr(), tfi.data_ptr(), tn.data_ptr())
    ctx.sync()
    ov, of, on = orc.edges_check_boxes(frm, to, L, lo, hi, nthreads=orc.max_threads())
    np.testing.assert_array_equal(tv.cpu().numpy(), ov)
    np.testing.assert_array_equal(tfi.cpu().numpy(), of)
    np.testing.assert_array_equal(tn.cpu().numpy().astype(np.uint32), on)
    ctx.close()


def _pack_status_numpy(valid):
    """The layout include/bpomp_cuda.h states: edge e -> word e//16, bits 2*(e%16)..+1, unused bits 0."""
    n = valid.size
    pad = np.zeros((n + 15) // 16 * 16, dtype=np.uint32)
    pad[:n] = valid & 3
    return (pad.reshape(-1, 16) << (2 * np.arange(16, dtype=np.uint32))).sum(axis=1, dtype=np.uint64).astype(np.uint32)


@pytest.mark.parametrize("count,shift", [(1, 0), (15, 0), (16, 0), (17, 0), (4096, 0), (1000003, 0), (100001, 1), (64, 3)])
def test_edge_status_pack_unpack(capi, count, shift):
    """Multi-GPU merge helpers: 2-bit packing of the status bytes and its inverse, aligned and unaligned."""
    import torch
    ctx = capi.Context(2, 0.01)
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(count)
    valid = rng.integers(0, 3, size=count, dtype=np.uint8)  # BLOCKED / FREE / DEFERRED
    buf = torch.zeros(count + 16, dtype=torch.uint8, device=dev)
    tv = buf[shift:shift + count]
    tv.copy_(torch.from_numpy(valid))
    words = (count + 15) // 16
    tp = torch.full((words,), -1, dtype=torch.int32, device=dev)
    out = torch.full((count + 16,), 7, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ctx.edge_status_pack_dev(tv.data_ptr(), count, tp.data_ptr())
    ctx.edge_status_unpack_dev(tp.data_ptr(), count, out[shift:].data_ptr())
    ctx.sync()
    np.testing.assert_array_equal(tp.cpu().numpy().view(np.uint32), _pack_status_numpy(valid))
    o = out.cpu().numpy()
    np.testing.assert_array_equal(o[shift:shift + count], valid)
    assert (o[:shift] == 7).all() and (o[shift + count:] == 7).all()  # nothing written outside [0, count)
    ctx.edge_status_pack_dev(0, 0, 0)  # empty input is a no-op
    ctx.close()


def test_full_size_config4_against_oracle(capi, orc, wl):
    """BASELINE config 4 at full size (10^7 edges, 7-DoF, 500 boxes), the workload bench.py times:
    every edge's nStates, validity flag and first-invalid slot bit-equal to the CPU oracle (the oracle needs a
    few seconds on all host cores), plus size-independent properties:
    (a) a chunked evaluation equals the one-shot evaluation; (b) an edge is BLOCKED iff its first
    invalid slot is in 1..n-2; (c) the state at the reported slot is invalid and every earlier slot
    state is valid (checked through the independent batched isValid kernel on a sample)."""
    d = 7
    count = 10 ** 7
    ctx, L, lo, hi = _ctx_boxes(capi, wl, d, 500, seed=3)
    frm, to = wl.random_edges(count, d, seed=4, max_len=wl.halton_radius(10 ** 6, d))
    v, f, n = ctx.edges_check(frm, to)
    ov, of, on = orc.edges_check_boxes(frm, to, L, lo, hi, nthreads=orc.max_threads())
    np.testing.assert_array_equal(n, on)
    np.testing.assert_array_equal(v, ov)
    np.testing.assert_array_equal(f, of)
    # the device-pointer entry point bench.py's `value` times, on the same batch
    import torch
    dev = torch.device("cuda:0")
    tf, tt = torch.from_numpy(frm).to(dev), torch.from_numpy(to).to(dev)
    tv = torch.empty(count, dtype=torch.uint8, device=dev)
    tfi = torch.empty(count, dtype=torch.int32, device=dev)
    tn = torch.empty(count, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx.edges_check_dev(tf.data_ptr(), tt.data_ptr(), count, tv.data_ptr(), tfi.data_ptr(), tn.data_ptr())
    ctx.sync()
    np.testing.assert_array_equal(tv.cpu().numpy(), ov)
    np.testing.assert_array_equal(tfi.cpu().numpy(), of)
    np.testing.assert_array_equal(tn.cpu().numpy().view(np.uint32), on)
    del tf, tt, tv, tfi, tn
    assert np.all((v == 0) == (f >= 1))
    assert np.all(f[v == 0] <= n[v == 0].astype(np.int64) - 2)
    assert np.all(n >= 2) and n.max() <= 13
    vc = np.concatenate([ctx.edges_check(frm[i:i + 2500000], to[i:i + 2500000])[0] for i in range(0, count, 2500000)])
    np.testing.assert_array_equal(v, vc)
    from batching_pomp_b200.capi import bisect_perm
    idx = np.nonzero(v == 0)[0][:20000]
    slot_states, prev_states = [], []
    for e in idx:
        p = bisect_perm(int(n[e]))
        t = (1 + p[f[e]]) / (n[e] + 1.0)
        slot_states.append(frm[e] + (to[e] - frm[e]) * t)
        for i in range(1, f[e]):
            prev_states.append(frm[e] + (to[e] - frm[e]) * ((1 + p[i]) / (n[e] + 1.0)))
    assert ctx.states_valid(np.array(slot_states)).sum() == 0
    if prev_states:
        assert ctx.states_valid(np.array(prev_states)).all()
    ctx.close()


def test_full_size_config3_csr_edges_indexed_against_oracle(capi, orc, wl):
    """The planner's real edge path at BASELINE config 3 size: r-disk rows of the 10^6-vertex 7-D Halton
    roadmap (r = haltonRadius) turned into indexed edges (source = the expanded vertex, BP.cpp:162-165),
    checked against the 500-box world through the indexed entry point, bit-equal to the oracle on the
    gathered endpoints."""
    d, N = 7, 10 ** 6
    verts = wl.halton(N, d)
    r = wl.halton_radius(N, d)
    ctx, L, lo, hi = _ctx_boxes(capi, wl, d, 500, seed=3)
    ctx.vertices_append(verts)
    nq = 1024
    q = (np.arange(nq, dtype=np.uint64) * (N // nq)).astype(np.uint32)
    rp, col, dist = ctx.neighbors_radius(q, r)
    lens = np.diff(rp.astype(np.int64))
    fi = np.repeat(q, lens)
    keep = fi != col
    fi, ti = fi[keep], col[keep]
    assert fi.size > 10 ** 6
    v, f, n = ctx.edges_check_indexed(fi, ti)
    ov, of, on = orc.edges_check_boxes(verts[fi], verts[ti], L, lo, hi, nthreads=orc.max_threads())
    np.testing.assert_array_equal(n, on)
    np.testing.assert_array_equal(v, ov)
    np.testing.assert_array_equal(f, of)
    # reversed orientation is a different edge (interpolate is not symmetric in floating point)
    v2, f2, n2 = ctx.edges_check_indexed(ti[:300000], fi[:300000])
    ov2, of2, on2 = orc.edges_check_boxes(verts[ti[:300000]], verts[fi[:300000]], L, lo, hi, nthreads=orc.max_threads())
    np.testing.assert_array_equal(v2, ov2)
    np.testing.assert_array_equal(f2, of2)
    ctx.close()


def test_more_than_65536_boxes(capi, orc, wl):
    """Box ids beyond 16 bits: candidates must not wrap."""
    d = 2
    rng = np.random.default_rng(7)
    nb = 70000
    c = rng.random((nb, d))
    h = rng.random((nb, d)) * 0.002 + 0.0005
    lo, hi = c - h, c + h
    L = wl.segment_length(d, 0.002)
    ctx = capi.Context(d, L)
    ctx.set_world_boxes(lo, hi)
    frm, to = wl.random_edges(4000, d, seed=12, max_len=0.05)
    v, f, n = ctx.edges_check(frm, to)
    ov, of, on = orc.edges_check_boxes(frm, to, L, lo, hi, nthreads=orc.max_threads())
    np.testing.assert_array_equal(n, on)
    np.testing.assert_array_equal(v, ov)
    np.testing.assert_array_equal(f, of)
    # a box that only a high id can explain: one isolated obstacle appended last
    lo2 = np.concatenate([np.full((nb, d), 5.0), [[0.45, 0.45]]])
    hi2 = np.concatenate([np.full((nb, d), 5.001), [[0.55, 0.55]]])
    ctx.set_world_boxes(lo2, hi2)
    v, f, n = ctx.edges_check(np.array([[0.3, 0.5]]), np.array([[0.7, 0.5]]))
    assert v[0] == 0 and f[0] >= 1
    st = wl.uniform(5, (20000, d))
    np.testing.assert_array_equal(ctx.states_valid(st), orc.states_valid_boxes(st, lo2, hi2))
    ctx.close()
