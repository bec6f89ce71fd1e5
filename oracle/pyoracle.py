"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when it was built in the
container that has /root/reference, for the compiled reference (oracle/_ref/libref_driver.so).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by som_lvq_pak_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_driver.so")
REF_BIN = os.path.join(HERE, "_ref", "bin")

_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int)
_u8 = C.POINTER(C.c_ubyte)
_s = C.POINTER(C.c_short)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def build(ref=True):
    """(Re)build liboracle.so, and the reference under oracle/_ref when /root/reference exists."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _msk(m):
    return None if m is None else np.ascontiguousarray(m, dtype=np.uint8)


class _Lib:
    prefix = ""

    def __init__(self, path):
        self.lib = C.CDLL(path)
        self.path = path

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)

    # ---- batch winner search -------------------------------------------------------
    def search(self, codes, data, k=1, mask=None):
        codes, data, mask = _f32(codes), _f32(data), _msk(mask)
        M, D = codes.shape
        N = data.shape[0]
        idx = np.empty((N, k), np.int32)
        diff = np.empty((N, k), np.float32)
        ret = np.empty(N, np.int32)
        f = self.fn("search")
        f.restype = None if self.prefix == "orc_" else C.c_int
        f(_p(codes, _f), C.c_long(M), C.c_int(D), _p(data, _f), _p(mask, _u8), C.c_long(N),
          C.c_int(k), _p(idx, _i), _p(diff, _f), _p(ret, _i))
        return idx, diff, ret

    def scalar(self, name, *args, restype=C.c_float, argtypes=None):
        f = self.fn(name)
        f.restype = restype
        if argtypes:
            f.argtypes = argtypes
        return f(*args)

    def hexa_dist(self, bx, by, tx, ty):
        return self.scalar("hexa_dist", bx, by, tx, ty, argtypes=[C.c_int] * 4)

    def rect_dist(self, bx, by, tx, ty):
        return self.scalar("rect_dist", bx, by, tx, ty, argtypes=[C.c_int] * 4)

    def linear_alpha(self, it, ln, a):
        return self.scalar("linear_alpha", it, ln, a, argtypes=[C.c_long, C.c_long, C.c_float])

    def inverse_t_alpha(self, it, ln, a):
        return self.scalar("inverse_t_alpha", it, ln, a, argtypes=[C.c_long, C.c_long, C.c_float])

    def vector_dist(self, a, b, ma=None, mb=None):
        a, b, ma, mb = _f32(a), _f32(b), _msk(ma), _msk(mb)
        f = self.fn("vector_dist")
        f.restype = C.c_float
        return f(_p(a, _f), _p(ma, _u8), _p(b, _f), _p(mb, _u8), C.c_int(a.shape[0]))

    def hitlist_vote(self, labels):
        lab = np.ascontiguousarray(labels, dtype=np.int64)
        f = self.fn("hitlist_vote")
        f.restype = C.c_long
        return f(lab.ctypes.data_as(C.POINTER(C.c_long)), C.c_int(lab.shape[0]))

    def qerror(self, codes, data, xdim, ydim, topol, neigh, qetype=0, radius=1.0, mask=None):
        codes, data, mask = _f32(codes), _f32(data), _msk(mask)
        M, D = codes.shape
        f = self.fn("qerror")
        f.restype = C.c_float
        return f(_p(codes, _f), C.c_long(M), C.c_int(D), C.c_int(xdim), C.c_int(ydim),
                 C.c_int(topol), C.c_int(neigh), _p(data, _f), _p(mask, _u8),
                 C.c_long(data.shape[0]), C.c_int(qetype), C.c_float(radius))


    def class_dists(self, codes, labels, median=True, mask=None, per_entry=False):
        codes, mask = _f32(codes), _msk(mask)
        M, D = codes.shape
        lab = np.ascontiguousarray(labels, np.int32)
        cls, noe, dists = np.empty(M, np.int32), np.empty(M, np.int32), np.empty(M, np.float32)
        near, found = np.full(M, np.nan, np.float32), np.full(M, -1, np.int32)
        f = self.fn("class_dists")
        f.restype = C.c_long
        n = f(_p(codes, _f), _p(mask, _u8), _p(lab, _i), C.c_long(M), C.c_int(D), C.c_int(int(median)),
              _p(cls, _i), _p(noe, _i), _p(dists, _f), _p(near, _f), _p(found, _i))
        if n < 0:
            raise RuntimeError("class_dists failed")
        if per_entry:
            return cls[:n], noe[:n], dists[:n], near, found
        return cls[:n], noe[:n], dists[:n]


class Oracle(_Lib):
    """Our C restatement (oracle/oracle.c)."""
    prefix = "orc_"

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        super().__init__(ORACLE_SO)

    def shuffle_order(self, n, seed):
        order = np.empty(n, np.int32)
        self.lib.orc_shuffle_order(C.c_long(n), C.c_int(seed), _p(order, _i))
        return order

    def som_train(self, codes, data, xdim, ydim, topol, neigh, length, alpha, radius,
                  alpha_type=1, order=None, mask=None, weight=None, fixed_xy=None):
        codes = _f32(codes).copy()
        data, mask = _f32(data), _msk(mask)
        M, D = codes.shape
        N = data.shape[0]
        order = None if order is None else np.ascontiguousarray(order, np.int32)
        weight = None if weight is None else np.ascontiguousarray(weight, np.int16)
        fixed_xy = None if fixed_xy is None else np.ascontiguousarray(fixed_xy, np.int16)
        f = self.lib.orc_som_train
        f.restype = C.c_int
        rc = f(_p(codes, _f), C.c_long(M), C.c_int(D), C.c_int(xdim), C.c_int(ydim),
               C.c_int(topol), C.c_int(neigh), _p(data, _f), _p(mask, _u8), _p(weight, _s),
               _p(fixed_xy, _s), C.c_long(N), _p(order, _i), C.c_long(length),
               C.c_float(alpha), C.c_float(radius), C.c_int(alpha_type))
        if rc:
            raise RuntimeError("orc_som_train failed")
        return codes

    def som_train_prefix(self, codes, data, xdim, ydim, topol, neigh, length, nsteps, alpha, radius,
                         alpha_type=1, order=None):
        codes = _f32(codes).copy()
        data = _f32(data)
        M, D = codes.shape
        order = None if order is None else np.ascontiguousarray(order, np.int32)
        f = self.lib.orc_som_train_prefix
        f.restype = C.c_int
        rc = f(_p(codes, _f), C.c_long(M), C.c_int(D), C.c_int(xdim), C.c_int(ydim), C.c_int(topol),
               C.c_int(neigh), _p(data, _f), None, None, None, C.c_long(data.shape[0]), _p(order, _i),
               C.c_long(length), C.c_long(nsteps), C.c_float(alpha), C.c_float(radius),
               C.c_int(alpha_type))
        if rc:
            raise RuntimeError("orc_som_train_prefix failed")
        return codes

    def lvq_train(self, algo, codes, code_label, data, data_label, length, alpha,
                  alpha_type=1, winlen=0.3, epsilon=0.1, order=None, mask=None,
                  unit_alpha=None):
        codes = _f32(codes).copy()
        data, mask = _f32(data), _msk(mask)
        M, D = codes.shape
        N = data.shape[0]
        cl = np.ascontiguousarray(code_label, np.int32)
        dl = np.ascontiguousarray(data_label, np.int32)
        order = None if order is None else np.ascontiguousarray(order, np.int32)
        ua = None
        if algo == 4:
            ua = (np.full(M, alpha, np.float32) if unit_alpha is None
                  else _f32(unit_alpha).copy())
        f = self.lib.orc_lvq_train
        f.restype = C.c_int
        rc = f(C.c_int(algo), _p(codes, _f), _p(cl, _i), C.c_long(M), C.c_int(D), _p(data, _f),
               _p(mask, _u8), _p(dl, _i), C.c_long(N), _p(order, _i), C.c_long(length),
               C.c_float(alpha), C.c_int(alpha_type), C.c_float(winlen), C.c_float(epsilon),
               _p(ua, _f))
        if rc:
            raise RuntimeError("orc_lvq_train failed")
        return codes, ua


    def remove_identicals(self, codes, mask=None):
        codes, mask = _f32(codes), _msk(mask)
        M, D = codes.shape
        keep = np.empty(M, np.int32)
        f = self.lib.orc_remove_identicals
        f.restype = C.c_long
        f(_p(codes, _f), _p(mask, _u8), C.c_long(M), C.c_int(D), _p(keep, _i))
        return np.nonzero(keep)[0]

    def sammon(self, codes, length, x, y, mask=None, errors=False):
        codes, mask = _f32(codes), _msk(mask)
        M, D = codes.shape
        x, y = _f32(x).copy(), _f32(y).copy()
        err = np.empty(length, np.float32) if errors else None
        f = self.lib.orc_sammon
        f.restype = C.c_int
        if f(_p(codes, _f), _p(mask, _u8), C.c_long(M), C.c_int(D), C.c_long(length), _p(x, _f), _p(y, _f),
             _p(err, _f)):
            raise RuntimeError("orc_sammon failed")
        return (x, y, err) if errors else (x, y)


class Reference(_Lib):
    """The unmodified reference behind oracle/ref_driver.c (only if `make ref` was run)."""
    prefix = "ref_"

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        super().__init__(REF_SO)

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def shuffle_order(self, n, seed):
        order = np.empty(n, np.int32)
        self.lib.ref_shuffle_order(C.c_long(n), C.c_int(seed), _p(order, _i))
        return order

    def search_time_only(self, codes, data, k=1):
        codes, data = _f32(codes), _f32(data)
        M, D = codes.shape
        cs = C.c_long(0)
        f = self.lib.ref_search_time_only
        f.restype = C.c_double
        t = f(_p(codes, _f), C.c_long(M), C.c_int(D), _p(data, _f), C.c_long(data.shape[0]),
              C.c_int(k), C.byref(cs))
        return t, cs.value

    def som_train(self, codes, data, xdim, ydim, topol, neigh, length, alpha, radius,
                  alpha_type=1, rand_seed=-1, mask=None, weight=None, fixed_xy=None):
        codes = _f32(codes).copy()
        data, mask = _f32(data), _msk(mask)
        M, D = codes.shape
        weight = None if weight is None else np.ascontiguousarray(weight, np.int16)
        fixed_xy = None if fixed_xy is None else np.ascontiguousarray(fixed_xy, np.int16)
        f = self.lib.ref_som_train
        f.restype = C.c_int
        rc = f(_p(codes, _f), C.c_long(M), C.c_int(D), C.c_int(xdim), C.c_int(ydim),
               C.c_int(topol), C.c_int(neigh), _p(data, _f), _p(mask, _u8), _p(weight, _s),
               _p(fixed_xy, _s), C.c_long(data.shape[0]), C.c_long(length), C.c_float(alpha),
               C.c_float(radius), C.c_int(alpha_type), C.c_int(rand_seed))
        if rc:
            raise RuntimeError("ref_som_train failed")
        return codes

    def lvq_train(self, algo, codes, code_label, data, data_label, length, alpha,
                  alpha_type=1, winlen=0.3, epsilon=0.1, rand_seed=-1, mask=None,
                  lra_in=None, lra_out=None):
        codes = _f32(codes).copy()
        data, mask = _f32(data), _msk(mask)
        M, D = codes.shape
        cl = np.ascontiguousarray(code_label, np.int32)
        dl = np.ascontiguousarray(data_label, np.int32)
        f = self.lib.ref_lvq_train
        f.restype = C.c_int
        rc = f(C.c_int(algo), _p(codes, _f), _p(cl, _i), C.c_long(M), C.c_int(D), _p(data, _f),
               _p(mask, _u8), _p(dl, _i), C.c_long(data.shape[0]), C.c_long(length),
               C.c_float(alpha), C.c_int(alpha_type), C.c_float(winlen), C.c_float(epsilon),
               C.c_int(rand_seed),
               None if lra_in is None else lra_in.encode(),
               None if lra_out is None else lra_out.encode())
        if rc:
            raise RuntimeError("ref_lvq_train failed")
        return codes

    def sammon(self, codes, length, seed, mask=None):
        """init_random(seed); remove_identicals; sammon_iterate -- positions of the surviving entries"""
        codes, mask = _f32(codes), _msk(mask)
        M, D = codes.shape
        x, y = np.empty(M, np.float32), np.empty(M, np.float32)
        f = self.lib.ref_sammon
        f.restype = C.c_long
        n = f(_p(codes, _f), _p(mask, _u8), C.c_long(M), C.c_int(D), C.c_long(length), C.c_int(seed),
              _p(x, _f), _p(y, _f))
        if n < 0:
            raise RuntimeError("ref_sammon failed")
        return x[:n], y[:n]
