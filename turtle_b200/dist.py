"""Multi-GPU plumbing: rays are independent, so the path shards with NO data-path
collective. One process per GPU (torch.distributed); the DEM is replicated by each
rank's own turtle_stepper_freeze; the only exchange is the final gather of the fixed-size
result records (96 bytes per ray) on one rank.
"""
import torch
import torch.distributed as dist

RECORD_BYTES = 96


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, rank, world_size):
    """Contiguous shard [first, last) of n rays for `rank`; sizes differ by at most 1."""
    base, extra = divmod(n, world_size)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def gather_records(local, dst=0):
    """Gather per-rank record tensors (uint8 [n_r, 96], n_r may differ by rank) on `dst`.
    Returns the concatenated uint8 [sum n_r, 96] tensor on dst, None elsewhere."""
    rank, size = world()
    if size == 1:
        return local
    counts = torch.zeros(size, dtype=torch.int64, device=local.device)
    counts[rank] = local.shape[0]
    dist.all_reduce(counts)
    most = int(counts.max().item())
    padded = local
    if local.shape[0] < most:
        padded = torch.zeros((most, RECORD_BYTES), dtype=torch.uint8, device=local.device)
        padded[:local.shape[0]] = local
    out = [torch.empty_like(padded) for _ in range(size)] if rank == dst else None
    dist.gather(padded.contiguous(), out, dst=dst)
    if rank != dst:
        return None
    return torch.cat([out[r][:int(counts[r].item())] for r in range(size)], 0)
