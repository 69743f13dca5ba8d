"""world_size-2 gloo test (CPU) of the N>1 host logic: shard the edge batch, evaluate each shard
independently (the CPU oracle stands in for the device here), merge with the all-gather helpers and
compare with the unsharded evaluation."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, count, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from batching_pomp_b200 import sharding, workloads
    from oracle import oracle
    d = 7
    L = workloads.segment_length(d)
    lo, hi = workloads.box_world(d, 500, seed=3)
    frm, to = workloads.random_edges(count, d, seed=4, max_len=0.35)
    b = sharding.shard_bounds(count, world)
    v, f, n = oracle.edges_check_boxes(frm[b[rank]:b[rank + 1]], to[b[rank]:b[rank + 1]], L, lo, hi)
    merged = sharding.allgather_ragged(torch.from_numpy(f.copy()))
    if count % world == 0:
        eq, _ = sharding.allgather_equal(torch.from_numpy(f.copy()))
        assert torch.equal(eq, merged)
    if rank == 0:
        q.put(merged.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _pack2(valid):
    """2 bits per status, 16 per word: the layout of bpomp_edge_status_pack_dev (include/bpomp_cuda.h)."""
    n = valid.size
    pad = np.zeros((n + 15) // 16 * 16, dtype=np.uint32)
    pad[:n] = valid & 3
    return (pad.reshape(-1, 16) << (2 * np.arange(16, dtype=np.uint32))).sum(axis=1, dtype=np.uint64).astype(np.uint32)


def _unpack2(words, n):
    return ((words[:, None] >> (2 * np.arange(16, dtype=np.uint32))) & 3).astype(np.uint8).reshape(-1)[:n]


def _packed_worker(rank, world, port, per_rank, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from batching_pomp_b200 import sharding
    valid = np.random.default_rng(100 + rank).integers(0, 3, size=per_rank, dtype=np.uint8)
    words = (per_rank + 15) // 16
    merged, _ = sharding.allgather_equal(torch.from_numpy(_pack2(valid).view(np.int32).copy()))
    merged = merged.numpy().view(np.uint32)
    # bench.py's finish(): one unpack over world*E when E % 16 == 0, else one per rank
    if per_rank % 16 == 0:
        out = _unpack2(merged, world * per_rank)
    else:
        out = np.concatenate([_unpack2(merged[g * words:(g + 1) * words], per_rank) for g in range(world)])
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("per_rank", [4096, 4099])
def test_packed_status_merge_layout(per_rank):
    """The N>1 merge of bench.py on gloo: every rank's 2-bit packed flags gathered and expanded back to the
    concatenation of the ranks' status bytes (the CUDA pack/unpack kernels are checked in test_gpu_parity)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_packed_worker, args=(r, 2, port, per_rank, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=180)
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    want = np.concatenate([np.random.default_rng(100 + r).integers(0, 3, size=per_rank, dtype=np.uint8) for r in range(2)])
    np.testing.assert_array_equal(out, want)


class _OracleCtx:
    """Stand-in for capi.Context in the gloo tests: the same method names and return shapes, answered by the
    CPU oracle (there is no GPU here; on a GPU box the sharded helpers are given a real Context)."""

    def __init__(self, orc, d, L, lo, hi, verts, bel_pts, bel_vals):
        self.orc, self.d, self.L, self.lo, self.hi = orc, d, L, lo, hi
        self.verts, self.bel_pts, self.bel_vals = verts, bel_pts, bel_vals

    def neighbors_radius(self, query_ids, r):
        return self.orc.neighbors_radius(self.verts, np.asarray(query_ids, dtype=np.uint32), r)

    def belief_estimate(self, queries):
        return self.orc.belief_estimate(self.bel_pts, self.bel_vals, queries)

    def edges_check(self, frm, to):
        return self.orc.edges_check_boxes(frm, to, self.L, self.lo, self.hi)


def _sharded_ops_inputs():
    from batching_pomp_b200 import workloads
    d = 3
    L = workloads.segment_length(d)
    lo, hi = workloads.box_world(d, 40, seed=3)
    verts = workloads.halton(700, d)
    bel_pts = workloads.uniform(21, (900, d))
    bel_vals = (workloads.uniform(22, (900,)) < 0.3).astype(np.float64)
    queries = workloads.uniform(23, (333, d))
    qids = np.arange(0, 700, 3, dtype=np.uint32)  # 234 queries: not divisible by 4
    frm, to = workloads.random_edges(1001, d, seed=24, max_len=0.4)
    return d, L, lo, hi, verts, bel_pts, bel_vals, queries, qids, frm, to


def _sharded_ops_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from batching_pomp_b200 import sharding
    from oracle import oracle
    d, L, lo, hi, verts, bel_pts, bel_vals, queries, qids, frm, to = _sharded_ops_inputs()
    ctx = _OracleCtx(oracle, d, L, lo, hi, verts, bel_pts, bel_vals)
    rp, col, dst = sharding.neighbors_radius_sharded(ctx, qids, 0.2)
    est = sharding.belief_estimate_sharded(ctx, queries)
    v, fi, ns = sharding.edges_check_sharded(ctx, frm, to)
    empty = sharding.belief_estimate_sharded(ctx, queries[:1])  # fewer units than ranks: empty shards
    if rank == world - 1:  # any rank holds the complete result
        q.put((rp, col, dst, est, v, fi, ns, empty))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_neighbours_belief_edges_equal_unsharded(orc, world):
    """query vertices, belief queries and edges sharded over the ranks, results all-gathered;
    every rank must end with exactly the unsharded result (world sizes 2 and 3, ragged shards)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_ops_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    rp, col, dst, est, v, fi, ns, empty = q.get(timeout=240)
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    d, L, lo, hi, verts, bel_pts, bel_vals, queries, qids, frm, to = _sharded_ops_inputs()
    rp0, col0, dst0 = orc.neighbors_radius(verts, qids, 0.2)
    np.testing.assert_array_equal(rp, rp0)
    np.testing.assert_array_equal(col, col0)
    np.testing.assert_array_equal(dst, dst0)
    np.testing.assert_array_equal(est, orc.belief_estimate(bel_pts, bel_vals, queries))
    v0, fi0, ns0 = orc.edges_check_boxes(frm, to, L, lo, hi)
    np.testing.assert_array_equal(v, v0)
    np.testing.assert_array_equal(fi, fi0)
    np.testing.assert_array_equal(ns, ns0)
    np.testing.assert_array_equal(empty, orc.belief_estimate(bel_pts, bel_vals, queries[:1]))


@pytest.mark.parametrize("count", [4000, 4001])
def test_sharded_edge_check_merges_to_unsharded(orc, count):
    from batching_pomp_b200 import workloads
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, count, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=180)
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    d = 7
    lo, hi = workloads.box_world(d, 500, seed=3)
    frm, to = workloads.random_edges(count, d, seed=4, max_len=0.35)
    v, f, n = orc.edges_check_boxes(frm, to, workloads.segment_length(d), lo, hi)
    np.testing.assert_array_equal(merged, f)


def test_shard_bounds_cover_everything():
    from batching_pomp_b200 import sharding
    for count in (0, 1, 7, 10 ** 7, 10 ** 7 + 3):
        for world in (1, 2, 4, 8):
            b = sharding.shard_bounds(count, world)
            assert b[0] == 0 and b[-1] == count and all(b[i] <= b[i + 1] for i in range(world))
            assert max(b[i + 1] - b[i] for i in range(world)) - min(b[i + 1] - b[i] for i in range(world)) <= 1
