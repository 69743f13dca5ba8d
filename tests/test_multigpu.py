"""Multi-GPU form of the batch operators through the C ABI (include/bpomp_cuda.h: bpomp_comm_*, *_sharded).
CPU: the shard arithmetic.  GPU (needs two devices; skipped otherwise — run with `gpurun --gpus 2`): one process
drives two contexts, one thread per context, results compared with the CPU oracle and with the single-GPU calls."""
import threading

import numpy as np
import pytest


def test_shard_range_covers_and_is_contiguous():
    from batching_pomp_b200 import capi
    capi.load()
    for count in (0, 1, 7, 16, 1000003, 10 ** 7):
        for n in (1, 2, 3, 8):
            prev = 0
            for r in range(n):
                b, e = capi.shard_range(count, n, r)
                assert b == prev and e >= b and e - b in (count // n, count // n + 1)
                prev = e
            assert prev == count
    with pytest.raises(capi.BpompError):
        capi.shard_range(10, 2, 2)


def _two_contexts(capi, dim, L):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctxs = [capi.Context(dim, L, device=0), capi.Context(dim, L, device=1)]
    return ctxs


def _in_threads(fns):
    out, err = [None] * len(fns), [None] * len(fns)

    def run(i):
        try:
            out[i] = fns[i]()
        except Exception as e:  # noqa: BLE001
            err[i] = e
    ts = [threading.Thread(target=run, args=(i,)) for i in range(len(fns))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(300)
    for e in err:
        if e is not None:
            raise e
    return out


@pytest.mark.gpu
def test_sharded_operators_two_gpus(orc):
    from batching_pomp_b200 import capi, workloads as wl
    d = 7
    L = wl.segment_length(d)
    lo, hi = wl.box_world(d, 500, seed=3)
    ctxs = _two_contexts(capi, d, L)
    capi.comm_init_all(ctxs)
    assert [c.comm_rank() for c in ctxs] == [(0, 2), (1, 2)]
    verts = wl.halton(20000, d)
    pts = wl.uniform(31, (4000, d))
    vals = (wl.uniform(32, (4000,)) < 0.3).astype(np.float64)
    for c in ctxs:  # replicated state
        c.set_bounds(0.0, 1.0)
        c.set_world_boxes(lo, hi)
        c.vertices_append(verts)
        c.belief_append(pts, vals)
    # edges, ragged shards (odd count), long edges at a fine resolution exercise the deferred sample orders too
    frm, to = wl.random_edges(200001, d, seed=4, max_len=0.35)
    res = _in_threads([lambda c=c: c.edges_check_sharded(frm, to) for c in ctxs])
    ov, of, on = orc.edges_check_boxes(frm, to, L, lo, hi, nthreads=orc.max_threads())
    for v, f, n in res:
        np.testing.assert_array_equal(v, ov)
        np.testing.assert_array_equal(f, of)
        np.testing.assert_array_equal(n, on)
    # neighbours
    q = np.arange(0, 20000, 37, dtype=np.uint32)
    r = wl.halton_radius(20000, d)
    res = _in_threads([lambda c=c: c.neighbors_radius_sharded(q, r) for c in ctxs])
    orp, ocol, odist = orc.neighbors_radius(verts, q, r, nthreads=orc.max_threads())
    for rp, col, dist in res:
        np.testing.assert_array_equal(rp, orp)
        np.testing.assert_array_equal(col, ocol)
        np.testing.assert_array_equal(dist, odist)
    # belief
    qs = wl.uniform(33, (3001, d))
    res = _in_threads([lambda c=c: c.belief_estimate_sharded(qs) for c in ctxs])
    want = orc.belief_estimate(pts, vals, qs, nthreads=orc.max_threads())
    for got in res:
        np.testing.assert_array_equal(got, want)
    # empty batches are fine
    res = _in_threads([lambda c=c: c.neighbors_radius_sharded(np.zeros(0, dtype=np.uint32), r) for c in ctxs])
    assert res[0][0].tolist() == [0]
    for c in ctxs:
        c.close()


@pytest.mark.gpu
def test_sharded_device_edge_path_two_gpus(orc):
    """The device-resident form bench.py times at N > 1: equal shards, packed status merged with one ncclAllGather
    on the communicator's stream."""
    import torch
    from batching_pomp_b200 import capi, workloads as wl
    d = 7
    L = wl.segment_length(d)
    lo, hi = wl.box_world(d, 500, seed=3)
    ctxs = _two_contexts(capi, d, L)
    capi.comm_init_all(ctxs)
    E = 100000
    shards = [wl.random_edges(E, d, seed=4 + 100 * g, max_len=0.35) for g in range(2)]
    words = (E + 15) // 16
    bufs = []
    for g, c in enumerate(ctxs):
        c.set_world_boxes(lo, hi)
        dev = torch.device("cuda", g)
        tf, tt = torch.from_numpy(shards[g][0]).to(dev), torch.from_numpy(shards[g][1]).to(dev)
        bufs.append(dict(tf=tf, tt=tt, v=torch.empty(E, dtype=torch.uint8, device=dev),
                         f=torch.empty(E, dtype=torch.int32, device=dev), n=torch.empty(E, dtype=torch.int32, device=dev),
                         p=torch.empty(words, dtype=torch.int32, device=dev),
                         pa=torch.empty(2 * words, dtype=torch.int32, device=dev),
                         out=torch.empty(2 * E, dtype=torch.uint8, device=dev)))
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)

    def step(g):
        c, b = ctxs[g], bufs[g]
        c.edges_check_sharded_dev(b["tf"].data_ptr(), b["tt"].data_ptr(), E, b["v"].data_ptr(), b["f"].data_ptr(),
                                  b["n"].data_ptr(), b["p"].data_ptr(), b["pa"].data_ptr())
        c.comm_wait()
        for k in range(2):
            c.edge_status_unpack_dev(b["pa"][k * words:].data_ptr(), E, b["out"][k * E:].data_ptr())
        c.sync()
        return b["out"].cpu().numpy(), b["f"].cpu().numpy()
    res = _in_threads([lambda g=g: step(g) for g in range(2)])
    want = [orc.edges_check_boxes(s[0], s[1], L, lo, hi, nthreads=orc.max_threads()) for s in shards]
    merged = np.concatenate([want[0][0], want[1][0]])
    for g in range(2):
        np.testing.assert_array_equal(res[g][0], merged)
        np.testing.assert_array_equal(res[g][1], want[g][1])
    for c in ctxs:
        c.close()


@pytest.mark.gpu
def test_sharded_device_edge_path_one_rank(orc):
    """The sharded device path on a one-rank communicator (runs on a single GPU): fast-kernel edges and edges left
    to the general kernel, a ragged batch that ends in a partial status word, and both buffer disciplines
    (alternating sets, the same set again) give the packed form of exactly the status bytes — which equal the
    oracle's."""
    import torch
    from batching_pomp_b200 import capi, workloads as wl
    d = 7
    L = wl.segment_length(d)
    lo, hi = wl.box_world(d, 500, seed=3)
    ctx = capi.Context(d, L, device=0)
    ctx.set_bounds(0.0, 1.0)
    ctx.set_world_boxes(lo, hi)
    capi.comm_init_all([ctx])
    E = 100003  # not a multiple of 32: partial tile, partial word
    frm, to = wl.random_edges(E, d, seed=4, max_len=0.35)
    frm[::50], to[::50] = wl.random_edges(len(frm[::50]), d, seed=5, max_len=2.5)  # long edges: left to the general kernel
    frm[7] = 1.5  # an endpoint outside the quantised zone: also the general kernel's
    words = (E + 15) // 16
    dev = torch.device("cuda", 0)
    tf, tt = torch.from_numpy(frm).to(dev), torch.from_numpy(to).to(dev)
    v = torch.empty(E, dtype=torch.uint8, device=dev)
    f = torch.empty(E, dtype=torch.int32, device=dev)
    n = torch.empty(E, dtype=torch.int32, device=dev)
    sets = [(torch.full((words,), -1, dtype=torch.int32, device=dev), torch.full((words,), -1, dtype=torch.int32, device=dev))
            for _ in range(2)]
    want = orc.edges_check_boxes(frm, to, L, lo, hi, nthreads=orc.max_threads())
    assert ctx.counter("fast_world") == 1
    for k in (0, 1, 0, 0, 1):  # alternating sets, then the same set twice in a row
        p, pa = sets[k]
        ctx.edges_check_sharded_dev(tf.data_ptr(), tt.data_ptr(), E, v.data_ptr(), f.data_ptr(), n.data_ptr(),
                                    p.data_ptr(), pa.data_ptr())
        ctx.comm_wait()
        out = torch.empty(E, dtype=torch.uint8, device=dev)
        ctx.edge_status_unpack_dev(pa.data_ptr(), E, out.data_ptr())
        ref = torch.empty(words, dtype=torch.int32, device=dev)
        ctx.edge_status_pack_dev(v.data_ptr(), E, ref.data_ptr())
        ctx.sync()
        assert ctx.counter("slow_edges") > 1000
        np.testing.assert_array_equal(v.cpu().numpy(), want[0])
        np.testing.assert_array_equal(f.cpu().numpy(), want[1])
        np.testing.assert_array_equal(out.cpu().numpy(), want[0])
        assert torch.equal(pa, ref) and torch.equal(p, ref)  # word for word what the separate pack pass writes
    ctx.close()
