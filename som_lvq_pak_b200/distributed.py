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


# ---------------------------------------------------------------------------- vfind
def trial_numbers(trials, rank, world):
    """trials are numbered trials..1 and run in that order (vfind.c:250-306); rank r takes every
    world-th one, so that the union over ranks is the reference's sequence"""
    return [n for k, n in enumerate(range(trials, 0, -1)) if k % world == rank]


def select_best(results):
    """results: (qerror float32, trial number) pairs from all ranks.  The reference keeps a map only
    on a strictly smaller error while counting the trial number DOWN (vfind.c:288), so among equal
    errors the LARGEST trial number wins."""
    best = None
    for q, n in sorted(results, key=lambda t: -t[1]):
        if best is None or q < best[0]:
            best = (q, n)
    return best


def vfind(data, test, xdim, ydim, topol, neigh, trials, length1, alpha1, radius1, length2, alpha2, radius2,
          alpha_type=1, qetype=0, group=None):
    """Multi-trial map search, one trial stream per GPU: every rank trains its share of the trials
    (randinit with seed = trial number, two training phases, quantization error on `test`), the
    (error, trial) pairs are all-gathered and the owner of the best map broadcasts it.  Training
    itself does not shard; this is the one place more GPUs help it.
    Returns (codes, qerror_sum, trial)."""
    import torch
    import torch.distributed as dist
    from . import engine as E
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = None
    for n in trial_numbers(trials, rank, world):
        codes = E.randinit_codes(data, xdim, ydim, n)
        codes = E.som_training(codes, data, xdim, ydim, topol, neigh, length1, alpha1, radius1, alpha_type)
        codes = E.som_training(codes, data, xdim, ydim, topol, neigh, length2, alpha2, radius2, alpha_type)
        q = E.find_qerror2(codes, test, xdim, ydim, topol, neigh, radius2)[0] if qetype else E.find_qerror(codes, test)
        if mine is None or q < mine[0]:
            mine = (np.float32(q), n, codes)
    if world == 1:
        return mine[2], mine[0], mine[1]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    pair = torch.tensor([float(mine[0]) if mine else float("inf"), float(mine[1]) if mine else 0.0],
                        dtype=torch.float64, device=dev)
    allp = [torch.empty_like(pair) for _ in range(world)]
    dist.all_gather(allp, pair, group=group)
    cands = [(np.float32(p[0].item()), int(p[1].item())) for p in allp if p[1].item() > 0]
    q, n = select_best(cands)
    owner = [r for r in range(world) if n in trial_numbers(trials, r, world)][0]
    M, D = xdim * ydim, np.asarray(data).shape[1]
    buf = torch.from_numpy(mine[2].copy() if rank == owner else np.empty((M, D), np.float32)).to(dev)
    dist.broadcast(buf, owner, group=group)
    return buf.cpu().numpy(), q, n
