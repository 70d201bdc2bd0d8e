"""ctypes binding of tests/emul/libneemul.so: the product's batched
branch-and-bound (csrc/nodeengine.cuh + csrc/bnbpool.cuh) compiled for the host
with one thread per CTA.  TEST INFRASTRUCTURE: lets `-m "not gpu"` tests and the
gloo world-size-2 tests drive the same tree / node / migration code that runs
as k_bnb_nodes on the device.  Never imported by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DIR = os.path.join(HERE, "emul")
LIB = os.path.join(DIR, "libneemul.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", DIR])
        L = C.CDLL(LIB)
        L.ne_emul_create.restype = C.c_void_p
        L.ne_emul_create.argtypes = ([C.c_int, C.c_int, C.c_int, C.c_double] + [C.c_void_p] * 11 +
                                     [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_long,
                                      C.c_int, C.c_int])
        L.ne_emul_destroy.argtypes = [C.c_void_p]
        L.ne_emul_round.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
        L.ne_emul_open.argtypes = [C.c_void_p]
        L.ne_emul_incumbent.argtypes = [C.c_void_p] * 5
        L.ne_emul_set_cutoff.argtypes = [C.c_void_p, C.c_double]
        L.ne_emul_clear.argtypes = [C.c_void_p]
        L.ne_emul_record_bytes.argtypes = [C.c_void_p]
        L.ne_emul_record_bytes.restype = C.c_long
        L.ne_emul_export.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ne_emul_import.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ne_emul_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.ne_emul_error.restype = C.c_char_p
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Pool:
    """d: problem in the glpb_create layout; stat: [m+n] basis to start the root from."""

    def __init__(self, d, stat, br_tech=4, bt_tech=3, pp_tech=2, tol_int=1e-5, tol_obj=1e-7, mip_gap=0.0,
                 node_lim=-1, batch=8, cap=4096, rii=None, sjj=None):
        self.L = lib()
        self.m, self.n, self.dir = int(d["m"]), int(d["n"]), int(d["dir"])
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        arrs = [i32(d["type"]), f64(d["lb"]), f64(d["ub"]), i32(stat), f64(d["coef"]), i32(d["kind"]),
                f64(rii) if rii is not None else None, f64(sjj) if sjj is not None else None,
                i32(d["A_ptr"]), i32(d["A_ind"]), f64(d["A_val"])]
        self.h = self.L.ne_emul_create(self.m, self.n, self.dir, float(d["c0"]), *[_p(a) for a in arrs],
                                       br_tech, bt_tech, pp_tech, tol_int, tol_obj, mip_gap, node_lim, batch, cap)
        if not self.h:
            raise RuntimeError(self.L.ne_emul_error().decode())

    def close(self):
        if self.h:
            self.L.ne_emul_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def round(self, max_tasks=-1):
        done = C.c_long()
        rc = self.L.ne_emul_round(self.h, max_tasks, C.byref(done))
        return rc, done.value

    def run(self):
        """rounds until the pool is empty or an error/limit code comes back"""
        while True:
            rc, _ = self.round()
            if rc != 1:
                return rc

    def open_count(self):
        return self.L.ne_emul_open(self.h)

    def incumbent(self):
        hs, hc, obj = C.c_int(), C.c_int(), C.c_double()
        x = np.zeros(self.m + self.n)
        self.L.ne_emul_incumbent(self.h, C.byref(hs), C.byref(hc), C.byref(obj), _p(x))
        return dict(have_sol=bool(hs.value), have_cut=bool(hc.value), obj=obj.value, x=x)

    def set_cutoff(self, obj):
        self.L.ne_emul_set_cutoff(self.h, float(obj))

    def clear(self):
        self.L.ne_emul_clear(self.h)

    def record_bytes(self):
        return self.L.ne_emul_record_bytes(self.h)

    def export_nodes(self, max_count):
        buf = np.zeros(max(1, max_count) * self.record_bytes(), np.uint8)
        cnt = C.c_int()
        rc = self.L.ne_emul_export(self.h, max_count, _p(buf), C.byref(cnt))
        assert rc == 0
        return buf[:cnt.value * self.record_bytes()].copy(), cnt.value

    def import_nodes(self, buf, count):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        assert self.L.ne_emul_import(self.h, _p(buf), count) == 0

    # ---- pointer-based migration, the interface of glpk_js_b200.bnb.BatchWorker ----
    def export_to(self, ptr, n):
        cnt = C.c_int()
        assert self.L.ne_emul_export(self.h, int(n), C.c_void_p(ptr), C.byref(cnt)) == 0
        return cnt.value

    def import_from(self, ptr, n):
        assert self.L.ne_emul_import(self.h, C.c_void_p(ptr), int(n)) == 0

    def stats(self):
        out = (C.c_long * 5)()
        self.L.ne_emul_stats(self.h, out)
        return dict(solved=out[0], tasks=out[1], rounds=out[2], iters=out[3], refacs=out[4])


class EmulWorker:
    """glpk_js_b200.bnb.BatchWorker over the host emulation (CPU tests of the
    sharding protocol, gloo world-size 2)"""

    def __init__(self, d, stat, **kw):
        self.d, self.stat, self.kw, self.pool = d, stat, kw, None

    def begin(self, batch=0, slab_nodes=0, **iocp):
        kw = dict(self.kw)
        kw.update(iocp)
        self.pool = Pool(self.d, self.stat, batch=batch or 8, cap=slab_nodes or 16384, **kw)
        return 0

    def clear(self):
        self.pool.clear()

    def round(self, max_tasks=-1):
        return self.pool.round(max_tasks)

    def incumbent(self):
        inc = self.pool.incumbent()
        big = 1.7976931348623157e308
        return inc["have_sol"], (inc["obj"] if inc["have_cut"] else (big if self.pool.dir == 1 else -big))

    def set_cutoff(self, obj):
        self.pool.set_cutoff(obj)

    def open_count(self):
        return self.pool.open_count()

    def record_bytes(self):
        return self.pool.record_bytes()

    def export_to(self, ptr, n):
        return self.pool.export_to(ptr, n)

    def import_from(self, ptr, n):
        self.pool.import_from(ptr, n)

    def solved(self):
        return self.pool.stats()["solved"]

    def end(self, ret):
        self.final = self.pool.incumbent()
        return ret
