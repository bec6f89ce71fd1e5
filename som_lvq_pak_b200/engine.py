"""Host-side mirror (Python) of the reference's plugin interface for the BMU path.

The reference hands its hot path around as function pointers in struct teach_params
(reference lvq_pak.h:186-204): `winner` (find_winner_euc / find_winner_knn), `vector_adapt`,
`neigh_adapt`, `alpha_func`.  Here the same operations are exposed at the granularity a GPU
can be fed at -- whole data sets -- with the same names, argument meaning and error
behaviour, all of them calling the CUDA library through the C ABI of include/bmu.h.
numpy arrays stand in for `struct entries`; nothing in this module computes on the CPU.
"""
import ctypes as C

import numpy as np

from . import _lib

TOPOL_HEXA, TOPOL_RECT = 3, 4
NEIGH_BUBBLE, NEIGH_GAUSSIAN = 1, 2
ALPHA_LINEAR, ALPHA_INVERSE_T = 1, 2
LVQ1, LVQ2, LVQ3, OLVQ1 = 1, 2, 3, 4
PATH_AUTO, PATH_EXACT, PATH_FILTER = 0, 1, 2
KMAX = 16


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _opt(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def init(device=0):
    _lib.check(_lib.load().bmu_init(device))


def device_info():
    sm, ma, mi, sh = C.c_int32(), C.c_int32(), C.c_int32(), C.c_size_t()
    _lib.check(_lib.load().bmu_device_info(C.byref(sm), C.byref(ma), C.byref(mi), C.byref(sh)))
    return {"sm_count": sm.value, "cc": (ma.value, mi.value), "smem_optin": sh.value}


def set_search_path(path):
    _lib.check(_lib.load().bmu_set_search_path(path))


def launch_count():
    return _lib.load().bmu_launch_count()


def last_search_breakdown():
    out = (C.c_long * 5)()
    _lib.check(_lib.load().bmu_last_search_breakdown(out))
    return dict(zip(("rows", "warp_rows", "seq_rows", "k2_certified", "k2_failed"), [int(v) for v in out]))


def last_search_kernel_ms():
    out = (C.c_float * 8)()
    _lib.check(_lib.load().bmu_last_search_kernel_ms(out))
    names = ("k1_data_prep", "k1_fast", "k1_warp", "k1_seq", "k2_row_prep", "k2_gemm", "k2_rerank", "k2_lists")
    return dict(zip(names, [float(v) for v in out]))


class Codebook:
    """A codebook resident on the GPU (reference: struct entries *codes)."""

    def __init__(self, codes):
        codes = _f32(codes)
        if codes.ndim != 2:
            raise ValueError("codes must be M x D")
        self.M, self.D = codes.shape
        self._h = _lib.load().bmu_codebook_create(_ptr(codes), self.M, self.D)
        if not self._h:
            raise RuntimeError("bmu_codebook_create: " + _lib.load().bmu_last_error().decode())

    def update(self, codes):
        codes = _f32(codes)
        assert codes.shape == (self.M, self.D)
        _lib.check(_lib.load().bmu_codebook_update(self._h, _ptr(codes)))

    def close(self):
        if self._h:
            _lib.load().bmu_codebook_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- WINNER_FUNCTION over a whole data set (lvq_pak.c:41-94 / 152-221)
    def find_winners(self, data, knn=1, mask=None):
        """returns (index N x knn int32, diff N x knn float32 [squared], nfound N int32)"""
        data = _f32(data)
        if data.ndim != 2 or data.shape[1] != self.D:
            raise ValueError("data must be N x %d" % self.D)
        mask = _opt(mask, np.uint8)
        N = data.shape[0]
        idx = np.empty((N, knn), np.int32)
        diff = np.empty((N, knn), np.float32)
        nf = np.empty(N, np.int32)
        _lib.check(_lib.load().bmu_search(self._h, _ptr(data), _ptr(mask), N, knn, _ptr(idx),
                                          _ptr(diff), _ptr(nf)))
        return idx, diff, nf

    def search_dev(self, d_data, N, knn, d_idx, d_diff, d_nfound, d_mask=None, stream=None):
        """device-pointer variant: arguments are integer device addresses (e.g. tensor.data_ptr())"""
        _lib.check(_lib.load().bmu_search_dev(self._h, d_data, d_mask, N, knn, d_idx, d_diff,
                                              d_nfound, stream))


class MultiCodebook:
    """The codebook replicated over every shard of the data-parallel search (bmu_multi_*,
    SURVEY.md 8e): rows are split into contiguous slices, one per GPU (or per logical shard),
    per-row results land in the caller's arrays in data order, statistics are combined by one
    grouped NCCL all-reduce."""

    def __init__(self, codes, nshards=0, code_label=None):
        lib = _lib.load()
        _lib.check(lib.bmu_multi_init(nshards))
        codes = _f32(codes)
        self.M, self.D = codes.shape
        self._h = lib.bmu_mcodebook_create(_ptr(codes), self.M, self.D)
        if not self._h:
            raise RuntimeError("bmu_mcodebook_create: " + lib.bmu_last_error().decode())
        if code_label is not None:
            cl = np.ascontiguousarray(code_label, np.int32)
            _lib.check(lib.bmu_mcodebook_set_labels(self._h, _ptr(cl)))

    @staticmethod
    def shards():
        return _lib.load().bmu_multi_shards()

    def update(self, codes):
        codes = _f32(codes)
        _lib.check(_lib.load().bmu_mcodebook_update(self._h, _ptr(codes)))

    def find_winners(self, data, knn=1, mask=None, stats=False, hist=False, sample_label=None, n_labels=0):
        """(idx, diff, nfound[, stats dict]) -- stats: sum_sqrt (double), n_found, hist[M],
        confusion[n_labels, n_labels] totals over all shards"""
        data = _f32(data)
        mask = _opt(mask, np.uint8)
        N = data.shape[0]
        idx = np.empty((N, knn), np.int32)
        diff = np.empty((N, knn), np.float32)
        nf = np.empty(N, np.int32)
        st = None
        keep = []
        if stats:
            st = _lib.Stats()
            if hist:
                h = np.zeros(self.M, np.int64)
                keep.append(h)
                st.hist = h.ctypes.data
            if sample_label is not None:
                sl = np.ascontiguousarray(sample_label, np.int32)
                cf = np.zeros((n_labels, n_labels), np.int64)
                keep += [sl, cf]
                st.sample_label = sl.ctypes.data
                st.confusion = cf.ctypes.data
                st.n_labels = n_labels
        _lib.check(_lib.load().bmu_multi_search(self._h, _ptr(data), _ptr(mask), N, knn, _ptr(idx), _ptr(diff),
                                                _ptr(nf), C.byref(st) if st is not None else None))
        if not stats:
            return idx, diff, nf
        out = {"sum_sqrt": float(st.sum_sqrt), "n_found": int(st.n_found)}
        if hist:
            out["hist"] = keep[0]
        if sample_label is not None:
            out["confusion"] = keep[-1]
        return idx, diff, nf, out

    def close(self):
        if self._h:
            _lib.load().bmu_mcodebook_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def search_stats_dev(d_idx, d_diff, d_nfound, N, k, M, d_sum, d_nfound_total, d_hist=None, d_sample_label=None,
                     d_code_label=None, n_labels=0, d_confusion=None, stream=None):
    """bmu_search_stats_dev: per-shard sums of a finished search, device addresses in and out"""
    _lib.check(_lib.load().bmu_search_stats_dev(d_idx, d_diff, d_nfound, N, k, M, d_sum, d_nfound_total, d_hist,
                                                d_sample_label, d_code_label, n_labels, d_confusion, stream))


def find_winner_euc(codes, data, mask=None):
    cb = Codebook(codes)
    try:
        return cb.find_winners(data, 1, mask)
    finally:
        cb.close()


def find_winner_knn(codes, data, knn, mask=None):
    cb = Codebook(codes)
    try:
        return cb.find_winners(data, knn, mask)
    finally:
        cb.close()


def find_qerror(codes, data, mask=None):
    """find_qerror (som_rout.c:678-731): sum over the samples of sqrt(diff), accumulated in one
    float in data order (bmu_replay_qerror); samples without a winner are skipped."""
    idx, diff, nf = find_winner_knn(codes, data, 1, mask)
    lib = _lib.load()
    lib.bmu_replay_qerror.restype = C.c_float
    return np.float32(lib.bmu_replay_qerror(_ptr(diff), _ptr(nf), C.c_long(diff.shape[0]), 1))


def find_qerror2(codes, data, xdim, ydim, topol, neigh, radius, mask=None):
    """find_qerror2 (som_rout.c:823-891, `qerror -qetype 1`): neighbourhood-weighted error around
    each sample's winner.  The per-sample values come from the GPU (bmu_qerror2); they are added
    here in data order in one float, as som_rout.c:872 does."""
    codes, data = _f32(codes), _f32(data)
    mask = _opt(mask, np.uint8)
    N = data.shape[0]
    out = np.empty(N, np.float32)
    cb = Codebook(codes)
    try:
        _lib.check(_lib.load().bmu_qerror2(cb._h, xdim, ydim, topol, neigh, C.c_float(radius), _ptr(data),
                                           _ptr(mask), N, _ptr(out)))
    finally:
        cb.close()
    q = np.float32(0.0)
    if N:
        q = np.cumsum(out, dtype=np.float32)[-1]      # cumsum is strictly sequential: float q += e
    return np.float32(q), out


def class_nearest(codes, labels, mask=None):
    """dissf of min_distances / med_distances (lvq_rout.c:325-350): per code vector the distance
    (vector_dist_euc) to the nearest LATER code vector with the same label; (dist, found)."""
    codes = _f32(codes)
    labels = np.ascontiguousarray(labels, np.int32)
    mask = _opt(mask, np.uint8)
    M, D = codes.shape
    dist = np.empty(M, np.float32)
    found = np.empty(M, np.int32)
    _lib.check(_lib.load().bmu_class_nearest(_ptr(codes), _ptr(mask), _ptr(labels), M, D, _ptr(dist), _ptr(found)))
    return dist, found


def hitlist_order(labels):
    """(label, count) pairs in the order add_hit (labels.c:370-410) leaves them after the labels were
    added one by one: a count that grows past its predecessor's moves in front of it, ties keep
    their place."""
    lab, freq = [], []
    for l in labels:
        l = int(l)
        if l in lab:
            i = lab.index(l)
            freq[i] += 1
            while i > 0 and freq[i - 1] < freq[i]:
                lab[i - 1], lab[i] = lab[i], lab[i - 1]
                freq[i - 1], freq[i] = freq[i], freq[i - 1]
                i -= 1
        else:
            lab.append(l)
            freq.append(1)
    return lab, freq


def class_distances(codes, labels, median=True, mask=None):
    """med_distances (median=True, lvq_rout.c:375-473) or min_distances (mean, lvq_rout.c:280-361):
    one value per class in hitlist order; returns (class labels, entries per class, dists)."""
    dist, found = class_nearest(codes, labels, mask)
    labels = np.asarray(labels)
    cls, noe = hitlist_order(labels)
    out = np.zeros(len(cls), np.float32)
    for c, l in enumerate(cls):
        d = dist[(labels == l) & (found != 0)]
        if d.size == 0:
            continue
        if median:
            out[c] = np.sort(d)[d.size // 2]                       # meds[not/2] after qsort
        else:
            out[c] = np.cumsum(d, dtype=np.float32)[-1] / np.float32(d.size)   # dists[i] += dissf; /= note
    return np.array(cls, np.int32), np.array(noe, np.int32), out


def identical_pairs(codes, mask=None, cap=1 << 20):
    """pairs (i < j) of code vectors at vector_dist_euc == 0.0, sorted (remove_identicals, sammon.c:83-127)"""
    codes = _f32(codes)
    mask = _opt(mask, np.uint8)
    M, D = codes.shape
    pairs = np.empty((cap, 2), np.int32)
    n = C.c_long(0)
    _lib.check(_lib.load().bmu_identical_pairs(_ptr(codes), _ptr(mask), M, D, _ptr(pairs), cap, C.byref(n)))
    return pairs[:n.value].copy()


def remove_identicals(codes, mask=None):
    """indices of the entries remove_identicals (sammon.c:83-127) keeps: walking the list, every later
    entry still present at distance 0 from the current one is dropped"""
    M = np.asarray(codes).shape[0]
    alive = np.ones(M, bool)
    for i, j in identical_pairs(codes, mask):          # sorted by (i, j): the order of the reference's walk
        if alive[i] and alive[j]:
            alive[j] = False
    return np.nonzero(alive)[0]


def sammon_init(M, seed):
    """initial positions of sammon_iterate (sammon.c:159-162) after init_random(seed)"""
    state = int(seed)
    x = np.empty(M, np.float32)
    for i in range(M):
        state = (state * 23) % 100000001               # orand, lvq_pak.c:470-473
        x[i] = np.float32(state % 32767 % M) / np.float32(M)
    y = np.arange(M, dtype=np.float32) / np.float32(M)
    return x, y


def sammon(codes, length, x, y, mask=None, errors=False):
    """sammon_iterate (sammon.c:129-262) from the initial positions (x, y); returns the final ones
    (and the per-sweep mapping errors when errors=True)"""
    codes = _f32(codes)
    mask = _opt(mask, np.uint8)
    M, D = codes.shape
    x = np.ascontiguousarray(x, np.float32).copy()
    y = np.ascontiguousarray(y, np.float32).copy()
    err = np.empty(length, np.float32) if errors else None
    _lib.check(_lib.load().bmu_sammon(_ptr(codes), _ptr(mask), M, D, length, _ptr(x), _ptr(y), _ptr(err)))
    return (x, y, err) if errors else (x, y)


# ---------------------------------------------------------------------------- host helpers
def rand_order(n, seed):
    """list order after `-rand seed` (datafile.c:1152-1188 driven by lvq_pak.c:459-473)"""
    order = np.empty(n, np.int32)
    _lib.load().bmu_rand_order(n, seed, _ptr(order))
    return order


def randinit_codes(data, xdim, ydim, seed, mask=None):
    """randinit_codes (som_rout.c:34-157) after init_random(seed): the map `randinit -rand seed` writes"""
    data = _f32(data)
    mask = _opt(mask, np.uint8)
    N, D = data.shape
    codes = np.empty((xdim * ydim, D), np.float32)
    _lib.load().bmu_randinit_codes(_ptr(data), _ptr(mask), N, D, xdim * ydim, int(seed), _ptr(codes))
    return codes


def som_schedule(le0, le1, length, alpha, radius, alpha_type, N, order=None, weight=None):
    n = le1 - le0
    sample = np.empty(n, np.int32)
    talp = np.empty(n, np.float32)
    trad = np.empty(n, np.float32)
    order = _opt(order, np.int32)
    weight = _opt(weight, np.int16)
    _lib.load().bmu_som_schedule(le0, le1, length, alpha, radius, alpha_type, N, _ptr(order),
                                 _ptr(weight), _ptr(sample), _ptr(talp), _ptr(trad))
    return sample, talp, trad


def lvq_schedule(le0, le1, length, alpha, alpha_type, N, order=None):
    n = le1 - le0
    sample = np.empty(n, np.int32)
    talp = np.empty(n, np.float32)
    order = _opt(order, np.int32)
    _lib.load().bmu_lvq_schedule(le0, le1, length, alpha, alpha_type, N, _ptr(order), _ptr(sample),
                                 _ptr(talp))
    return sample, talp


# ---------------------------------------------------------------------------- training
class Trainer:
    """Device-resident codebook + data for chunked training (snapshot boundaries)."""

    def __init__(self, codes, data, mask=None):
        codes, data = _f32(codes), _f32(data)
        mask = _opt(mask, np.uint8)
        self.M, self.D = codes.shape
        self.N = data.shape[0]
        self._h = _lib.load().bmu_trainer_create(_ptr(codes), self.M, self.D, _ptr(data),
                                                 _ptr(mask), self.N)
        if not self._h:
            raise RuntimeError("bmu_trainer_create: " + _lib.load().bmu_last_error().decode())

    def set_som(self, xdim, ydim, topol, neigh, fixed_xy=None):
        fixed_xy = _opt(fixed_xy, np.int16)
        _lib.check(_lib.load().bmu_trainer_set_som(self._h, xdim, ydim, topol, neigh, _ptr(fixed_xy)))

    def set_lvq(self, algo, code_label, data_label, win_thr=0.0, epsilon=0.0, alpha_cap=0.0,
                unit_alpha=None):
        cl = np.ascontiguousarray(code_label, np.int32)
        dl = np.ascontiguousarray(data_label, np.int32)
        ua = _opt(unit_alpha, np.float32)
        _lib.check(_lib.load().bmu_trainer_set_lvq(self._h, algo, _ptr(cl), _ptr(dl), win_thr,
                                                   epsilon, alpha_cap, _ptr(ua)))

    def steps(self, sample, talp=None, trad=None):
        sample = np.ascontiguousarray(sample, np.int32)
        talp, trad = _opt(talp, np.float32), _opt(trad, np.float32)
        _lib.check(_lib.load().bmu_trainer_steps(self._h, _ptr(sample), _ptr(talp), _ptr(trad),
                                                 sample.shape[0]))

    def codes(self):
        out = np.empty((self.M, self.D), np.float32)
        _lib.check(_lib.load().bmu_trainer_get_codes(self._h, _ptr(out)))
        return out

    def unit_alpha(self):
        out = np.empty(self.M, np.float32)
        _lib.check(_lib.load().bmu_trainer_get_unit_alpha(self._h, _ptr(out)))
        return out

    def last_ms(self):
        return float(_lib.load().bmu_trainer_last_ms(self._h))

    def close(self):
        if self._h:
            _lib.load().bmu_trainer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def som_training(codes, data, xdim, ydim, topol, neigh, length, alpha, radius,
                 alpha_type=ALPHA_LINEAR, rand_seed=None, mask=None, weight=None, fixed_xy=None,
                 snapshot_interval=0, snapshot_cb=None):
    """som_training (som_rout.c:556-671).  rand_seed None = list order (no -rand).  With a
    snapshot interval the run is cut at the reference's snapshot steps (`le % interval == 0
    and le > 0`, evaluated after step le: som_rout.c:650) and snapshot_cb(le, codes) is called."""
    data = _f32(data)
    N = data.shape[0]
    order = None if rand_seed is None else rand_order(N, rand_seed)
    tr = Trainer(codes, data, mask)
    try:
        tr.set_som(xdim, ydim, topol, neigh, fixed_xy)
        snaps = []
        if snapshot_interval and snapshot_cb:
            snaps = list(range(snapshot_interval, length, snapshot_interval))
        le0 = 0
        for le in snaps + [None]:
            le1 = length if le is None else le + 1
            if le1 > le0:
                s, ta, tr_ = som_schedule(le0, le1, length, alpha, radius, alpha_type, N, order, weight)
                tr.steps(s, ta, tr_)
            if le is not None:
                snapshot_cb(le, tr.codes())
            le0 = le1
        return tr.codes()
    finally:
        tr.close()


def lvq_training(algo, codes, code_label, data, data_label, length, alpha,
                 alpha_type=ALPHA_LINEAR, winlen=0.3, epsilon=0.1, rand_seed=None, mask=None,
                 unit_alpha=None):
    """lvq1/olvq1/lvq2/lvq3_training (lvq_rout.c:498-916).  Returns codes (and the per-unit
    rates for OLVQ1).  For OLVQ1 `alpha` is both the initial per-unit rate (when unit_alpha is
    None) and the cap, as in lvq_rout.c:614-627,670-672."""
    data = _f32(data)
    N = data.shape[0]
    M = np.asarray(codes).shape[0]
    order = None if rand_seed is None else rand_order(N, rand_seed)
    w = np.float32(winlen)
    win_thr = float(np.float32(np.float32(1) - w) / np.float32(np.float32(1) + w))   # lvq_rout.c:770
    ua = None
    if algo == OLVQ1:
        ua = np.full(M, alpha, np.float32) if unit_alpha is None else _f32(unit_alpha)
    tr = Trainer(codes, data, mask)
    try:
        tr.set_lvq(algo, code_label, data_label, win_thr, epsilon, alpha, ua)
        s, ta = lvq_schedule(0, length, length, alpha, alpha_type, N, order)
        tr.steps(s, ta, None)
        out = tr.codes()
        return (out, tr.unit_alpha()) if algo == OLVQ1 else out
    finally:
        tr.close()
