"""Multi-GPU plumbing for the sharded edge-evaluation path: one process per GPU,
torch.distributed for the rendezvous and the only collective the path has — the all-gather that merges
the per-shard first-invalid-slot results.  The collision world and the vertex table are replicated."""
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
