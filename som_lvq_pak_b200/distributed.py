"""Data-parallel batch search over the GPUs of one node.

One process per GPU (torch.distributed).  Samples are independent and the codebook is
read-only, so the rows are cut into contiguous shards (which keeps the data order for the
host replay), the codebook is replicated by one broadcast, every rank searches its shard
with no data-path collective, and only the small statistics vector -- qerror sum, found
count, BMU histogram, confusion counts -- is combined by ONE all-reduce.  torch.distributed
is plumbing here; all compute is in libbmu_b200.so.  Online training does not shard
(step t+1 reads the codebook of step t): replicas only."""
import numpy as np

TILE = 128


def shard_bounds(n_rows, rank, world):
    """contiguous, balanced row shard [lo, hi) of rank; boundaries on 128-row tiles"""
    tiles = (n_rows + TILE - 1) // TILE
    base, rem = divmod(tiles, world)
    lo_t = rank * base + min(rank, rem)
    hi_t = lo_t + base + (1 if rank < rem else 0)
    return min(lo_t * TILE, n_rows), min(hi_t * TILE, n_rows)


def pack_stats(qsum, n_found, hist=None, confusion=None):
    """one float64 vector for the all-reduce; integer counts < 2^53 stay exact"""
    parts = [np.array([qsum, n_found], np.float64)]
    if hist is not None:
        parts.append(np.asarray(hist, np.float64).ravel())
    if confusion is not None:
        parts.append(np.asarray(confusion, np.float64).ravel())
    return np.concatenate(parts)


def unpack_stats(vec, M=0, L=0):
    vec = np.asarray(vec)
    out = {"qsum": float(vec[0]), "n_found": int(round(vec[1]))}
    off = 2
    if M:
        out["hist"] = np.rint(vec[off:off + M]).astype(np.int64)
        off += M
    if L:
        out["confusion"] = np.rint(vec[off:off + L * L]).astype(np.int64).reshape(L, L)
    return out


def allreduce_stats(vec, group=None):
    """sum the packed statistics over all ranks (NCCL on GPUs, gloo in the CPU tests)"""
    import torch
    import torch.distributed as dist
    t = vec if isinstance(vec, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(vec, np.float64))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, group=group)
    return t


def gather_rows(local, n_rows, group=None):
    """concatenate per-rank row results on rank 0 in shard (= data) order for the host replay"""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [hi - lo for lo, hi in (shard_bounds(n_rows, r, world) for r in range(world))]
    big = max(sizes)
    # gather needs equal shapes: pad every shard to the largest one, trim on rank 0
    padded = local
    if local.shape[0] < big:
        pad = torch.zeros((big - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                          device=local.device)
        padded = torch.cat([local, pad])
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
    dist.gather(padded.contiguous(), bufs, dst=0, group=group)
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)]) if rank == 0 else None
