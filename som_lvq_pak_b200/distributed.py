"""Process-per-GPU plumbing for the data-parallel batch search.

The split itself -- contiguous row shards, replicated codebook, per-shard statistics, ONE grouped
NCCL all-reduce of {double sum} + {int64 counts} -- lives in libbmu_b200.so (csrc/bmu_multi.cu):
`bmu_multi_*` when one process drives all GPUs (the C hosts), `bmu_comm_*` when a launcher starts
one process per GPU (torchrun; bench.py --gpus N).  What is left here is the launcher's part of
the second mode: torch.distributed moves the 128-byte NCCL unique id from rank 0 to the other
ranks (gloo in the CPU tests, NCCL on the GPUs), and `ShardedSearch` strings the library calls
together on device buffers.  No arithmetic happens in this module.  Online training does not
shard (step t+1 reads the codebook of step t): replicas only (see `vfind`)."""
import ctypes as C

import numpy as np



def shard_bounds(n_rows, rank, world):
    """rows [lo, hi) of a rank: the library's own rule (bmu_multi_shard_bounds: contiguous,
    balanced, cut at multiples of 512 rows), so that both modes shard identically"""
    from . import _lib
    lo, hi = C.c_long(), C.c_long()
    _lib.load().bmu_multi_shard_bounds(n_rows, world, rank, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def exchange_unique_id(make_id, group=None):
    """rank 0 calls make_id() -> 128 bytes; every rank returns those bytes (one broadcast)"""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return bytes(make_id())
    rank = dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        t = torch.frombuffer(bytearray(make_id()), dtype=torch.uint8).clone()
    t = t.to(dev)
    dist.broadcast(t, 0, group=group)
    return bytes(t.cpu().numpy().tobytes())


def comm_init(group=None):
    """bind this rank's bmu_init device to a library-owned NCCL communicator spanning the ranks of
    the torch.distributed group; returns (rank, world)"""
    import torch.distributed as dist
    from . import _lib
    lib = _lib.load()
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0, 1
    rank, world = dist.get_rank(group), dist.get_world_size(group)

    def make_id():
        buf = (C.c_ubyte * 128)()
        _lib.check(lib.bmu_comm_unique_id(buf))
        return bytes(buf)
    uid = exchange_unique_id(make_id, group)
    buf = (C.c_ubyte * 128).from_buffer_copy(uid)
    _lib.check(lib.bmu_comm_init_rank(world, rank, buf))
    return rank, world


class ShardedSearch:
    """One rank's part of a sharded search on DEVICE buffers (torch tensors are only the allocator):
    search the local rows, reduce the shard's statistics on the device, combine them over the ranks
    with the library's grouped NCCL all-reduce.  `stats` is ONE int64 buffer: [0] the bits of the
    double sum of sqrt(diff), [1] n_found, [2:2+M] BMU hits per code vector."""

    def __init__(self, cb_handle, M, rows, k, device):
        import torch
        self.cb, self.M, self.rows, self.k = cb_handle, M, rows, k
        self.idx = torch.empty((rows, k), dtype=torch.int32, device=device)
        self.diff = torch.empty((rows, k), dtype=torch.float32, device=device)
        self.nf = torch.empty(rows, dtype=torch.int32, device=device)
        self.stats = torch.zeros(2 + M, dtype=torch.int64, device=device)
        self.stream = torch.cuda.current_stream(device).cuda_stream

    def step(self, d_data, allreduce=True):
        from . import _lib
        lib = _lib.load()
        p = self.stats.data_ptr()
        self.stats.zero_()
        _lib.check(lib.bmu_search_dev(self.cb, d_data, None, self.rows, self.k, self.idx.data_ptr(),
                                      self.diff.data_ptr(), self.nf.data_ptr(), self.stream))
        _lib.check(lib.bmu_search_stats_dev(self.idx.data_ptr(), self.diff.data_ptr(), self.nf.data_ptr(), self.rows,
                                            self.k, self.M, p, p + 8, p + 16, None, None, 0, None, self.stream))
        if allreduce:
            _lib.check(lib.bmu_comm_allreduce_stats_dev(p, 1, p + 8, 1 + self.M, self.stream))

    def totals(self):
        """(sum of sqrt(diff) as float, n_found, hist) -- synchronises"""
        h = self.stats.cpu().numpy()
        return float(h[:1].view(np.float64)[0]), int(h[1]), h[2:]


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
