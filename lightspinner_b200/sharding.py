"""Column sharding across the GPUs of one node (one process per GPU, torch.distributed for the plumbing).

Atmosphere columns are independent -- no arithmetic crosses columns anywhere on the MALI path (response_fn.py:23-39
rebuilds everything per column) -- so the path shards with NO data-path collective: rank r owns a contiguous block of
columns, runs the same kernels on it, and the only exchange is the final gather of results (emergent intensities,
optionally populations).  Results are therefore bit-identical for 1, 2, 4 or 8 GPUs.
"""
import torch
import torch.distributed as dist


def shard_range(ncol_total, world, rank):
    """Contiguous block of columns owned by `rank`: (first column, count); the remainder goes to the low ranks."""
    base, rem = divmod(int(ncol_total), int(world))
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def gather_columns(local, ncol_total, group=None):
    """All-gather per-column results.  `local` is [count_of_this_rank, ...]; returns [ncol_total, ...] on every
    rank, columns in global order.  NCCL over NVLink for CUDA tensors, gloo for CPU tensors (tests)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = [shard_range(ncol_total, world, r)[1] for r in range(world)]
    if local.shape[0] != counts[rank]:
        raise ValueError('rank %d holds %d columns, expected %d' % (rank, local.shape[0], counts[rank]))
    tail = tuple(local.shape[1:])
    if len(set(counts)) == 1:
        out = torch.empty((ncol_total,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # uneven split: pad every shard to the largest one, gather, drop the padding
    cmax = max(counts)
    padded = torch.zeros((cmax,) + tail, dtype=local.dtype, device=local.device)
    padded[:counts[rank]] = local
    out = torch.empty((world * cmax,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    out = out.view((world, cmax) + tail)
    return torch.cat([out[r, :counts[r]] for r in range(world)], dim=0)
