"""Multi-GPU plumbing: rays are independent, so the path shards with NO data-path
collective. One process per GPU (torch.distributed); the DEM is replicated by each
rank's own turtle_stepper_freeze; the only exchange is the delivery of the fixed-size
result records (96 bytes per ray) to one rank:

  * `PeerRecords`: the consumer rank owns the whole result array; every other rank maps
    it (CUDA IPC, turtle_b200_peer_*) and its trace kernel stores each record over
    NVLink the moment the ray ends -- the exchange overlaps the stepping, no gather;
  * `gather_records`: a plain gather after the kernel (the fallback when the GPUs
    cannot reach each other, and the gloo path of the CPU tests).
"""
import ctypes as C

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


class PeerRecords:
    """Result records of `counts[r]` rays per rank, all in the memory of rank `dst`.

    ``view`` (on every rank) is a raw device pointer to this rank's slice of the array:
    pass it as the `results` of Plan.trace_device. ``tensor`` (on dst only) is the whole
    uint8 [sum counts, 96] array. Collective: construct it, `ready()` it and `close()` it
    on all ranks."""

    def __init__(self, counts, dst=0):
        from ._lib import lib
        from .api import _check
        self._lib, self._check = lib, _check
        self.rank, self.size = world()
        self.dst = dst
        self.counts = [int(c) for c in counts]
        self.first = sum(self.counts[:self.rank])
        total = sum(self.counts)
        self._base = C.c_void_p()
        self._owner = self.rank == dst
        handle = torch.zeros(64, dtype=torch.uint8)
        if self._owner:
            _check(lib.turtle_b200_peer_alloc(max(total, 1) * RECORD_BYTES, C.byref(self._base)))
            if self.size > 1:
                buf = (C.c_ubyte * 64)()
                _check(lib.turtle_b200_peer_export(self._base, buf))
                handle = torch.tensor(list(buf), dtype=torch.uint8)
        if self.size > 1:
            dev = torch.device("cuda", torch.cuda.current_device())
            h = handle.to(dev)
            dist.broadcast(h, src=dst)
            if not self._owner:
                buf = (C.c_ubyte * 64)(*h.cpu().tolist())
                _check(lib.turtle_b200_peer_open(buf, C.byref(self._base)))
        self.view = self._base.value + self.first * RECORD_BYTES
        self.tensor = None
        if self._owner:
            self.tensor = _as_tensor(self._base.value, total * RECORD_BYTES).view(total, RECORD_BYTES)

    def ready(self):
        """All records of all ranks have landed: every rank's stream is drained (a
        finished kernel's peer stores are visible), then the ranks meet."""
        torch.cuda.synchronize()
        if self.size > 1:
            dist.barrier()

    def close(self):
        if self._base.value is None:
            return
        torch.cuda.synchronize()
        if self.size > 1:
            dist.barrier()
        if self._owner:
            self.tensor = None
            if self.size > 1:
                dist.barrier()  # peers close before the owner frees
            self._check(self._lib.turtle_b200_peer_free(self._base))
        else:
            self._check(self._lib.turtle_b200_peer_close(self._base))
            dist.barrier()
        self._base = C.c_void_p()


class _RawCuda:
    """__cuda_array_interface__ over a raw device pointer (no ownership)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
            "strides": None}


def _as_tensor(ptr, nbytes):
    if nbytes == 0:
        return torch.empty((0,), dtype=torch.uint8, device="cuda")
    return torch.as_tensor(_RawCuda(ptr, nbytes), device=torch.device("cuda", torch.cuda.current_device()))
