"""Multi-GPU plumbing for the sharded edge-evaluation path: one process per GPU,
torch.distributed for the rendezvous and the only collective the path has — the all-gather that merges
the per-shard first-invalid-slot results.  The collision world and the vertex table are replicated."""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(count, world):
    """Contiguous shards [bounds[g], bounds[g+1]) of `count` units over `world` ranks."""
    return [(count * g) // world for g in range(world + 1)]


def chunk_bounds(count, nchunk):
    return [(count * c) // nchunk for c in range(nchunk + 1)]


def allgather_equal(local, out=None, async_op=False):
    """All ranks contribute tensors of the same length; returns ([world*len] tensor, work)."""
    world = dist.get_world_size()
    if out is None:
        out = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
    work = dist.all_gather_into_tensor(out, local.contiguous(), async_op=async_op)
    return out, work


def allgather_ragged(local):
    """Shards of different lengths (count not divisible by world): gather the lengths, pad, gather, trim.
    Returns the concatenation in rank order."""
    world = dist.get_world_size()
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    lens = torch.empty(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(lens, n)
    m = int(lens.max())
    pad = torch.zeros(m, dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local
    out = torch.empty(world * m, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[g * m:g * m + int(lens[g])] for g in range(world)])


# ---------------------------------------------------------------------------
# Sharded operators: the collision world, the vertex table and the belief points are
# replicated on every rank's device context; the UNITS (query vertices, belief queries, edges) are
# split into contiguous slices, every rank evaluates its slice on its own GPU through `ctx`
# (a batching_pomp_b200.capi.Context) and the results are merged with all-gathers, so every rank ends
# with the complete result in the unsharded order.  No collective touches the data path itself.
# ---------------------------------------------------------------------------
def _comm_tensor(a):
    """numpy -> tensor on the device the process group communicates over (CUDA for nccl, CPU for gloo)."""
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dist.get_backend() == "nccl":
        t = t.cuda()
    return t


def _gather_np(a):
    return allgather_ragged(_comm_tensor(a)).cpu().numpy()


def neighbors_radius_sharded(ctx, query_ids, r):
    """bpomp_neighbors_radius with the query vertices sharded over the ranks.  Every rank passes the same
    `query_ids`; returns (row_ptr[nq+1], col[nnz], dist[nnz]) complete on every rank, rows in query order."""
    q = np.ascontiguousarray(query_ids, dtype=np.uint32)
    world, rank = dist.get_world_size(), dist.get_rank()
    b = shard_bounds(q.size, world)
    rp, col, dst = ctx.neighbors_radius(q[b[rank]:b[rank + 1]], r)
    counts = _gather_np(np.diff(np.asarray(rp, dtype=np.uint64)).astype(np.int64))
    row_ptr = np.zeros(q.size + 1, dtype=np.uint64)
    np.cumsum(counts, out=row_ptr[1:])
    return row_ptr, _gather_np(np.asarray(col, dtype=np.int64)).astype(np.uint32), _gather_np(dst)


def belief_estimate_sharded(ctx, queries):
    """bpomp_belief_estimate with the queries sharded over the ranks (the belief points are replicated:
    every rank appended the same points in the same order)."""
    qs = np.ascontiguousarray(queries, dtype=np.float64)
    world, rank = dist.get_world_size(), dist.get_rank()
    b = shard_bounds(qs.shape[0], world)
    return _gather_np(ctx.belief_estimate(qs[b[rank]:b[rank + 1]]))


def edges_check_sharded(ctx, frm, to):
    """bpomp_edges_check with the edges sharded over the ranks; returns (valid, first_invalid_slot, nstates)
    for all edges on every rank."""
    f = np.ascontiguousarray(frm, dtype=np.float64)
    t = np.ascontiguousarray(to, dtype=np.float64)
    world, rank = dist.get_world_size(), dist.get_rank()
    b = shard_bounds(f.shape[0], world)
    v, fi, ns = ctx.edges_check(f[b[rank]:b[rank + 1]], t[b[rank]:b[rank + 1]])
    return (_gather_np(v), _gather_np(fi), _gather_np(np.asarray(ns, dtype=np.int64)).astype(np.uint32))
