"""Multi-GPU plumbing: images are independent units, so a batch shards by image
with no collective on the data path.  torch.distributed (NCCL on GPUs, gloo in
the CPU tests) is used only to gather per-image stream lengths and, when asked,
the variable-length streams themselves.
"""
from typing import List, Sequence, Tuple


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous block [lo, hi) of a batch of n_items owned by `rank`"""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_by_cost(costs: Sequence[float], world: int) -> List[List[int]]:
    """size-balanced assignment for mixed-size batches: greedy longest-processing-time
    on cost (e.g. H*W + max_bits); returns the item indices of every rank."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    loads = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda q: loads[q])
        out[r].append(i)
        loads[r] += costs[i]
    for r in range(world):
        out[r].sort()
    return out


def gather_lengths(nbits, max_n, group=None):
    """all-gather of per-image (nbits, max_n); every rank holds an equal-sized shard.
    Returns tensors of shape (world * B_local,)"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    packed = torch.stack([nbits.to(torch.int64), max_n.to(torch.int64)], dim=1).contiguous()
    out = torch.empty((world * packed.shape[0], 2), dtype=torch.int64, device=packed.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    return out[:, 0], out[:, 1]


def gather_streams(streams, nbits, dst=0, group=None):
    """gather the variable-length streams of every rank on `dst` (padded rows are
    exchanged at the largest used length, not the full row stride).
    Returns a list of bytes objects on dst (rank order, then image order), None elsewhere."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    nbytes = (nbits.to(torch.int64) + 7) // 8
    local_max = nbytes.max().reshape(1)
    dist.all_reduce(local_max, op=dist.ReduceOp.MAX, group=group)
    width = int(local_max.item())
    rows = streams[:, :width].contiguous()
    all_rows = torch.empty((world * rows.shape[0], width), dtype=rows.dtype, device=rows.device)
    all_nbytes = torch.empty((world * nbytes.shape[0],), dtype=torch.int64, device=rows.device)
    dist.all_gather_into_tensor(all_rows, rows, group=group)
    dist.all_gather_into_tensor(all_nbytes, nbytes.contiguous(), group=group)
    if rank != dst:
        return None
    host = all_rows.cpu().numpy()
    lens = all_nbytes.cpu().numpy()
    return [host[i, :int(lens[i])].tobytes() for i in range(host.shape[0])]
