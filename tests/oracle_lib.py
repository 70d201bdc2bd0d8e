"""ctypes binding of the CPU oracle (oracle/libglpo.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libglpo.so")

GLP_MIN, GLP_MAX = 1, 2
GLP_CV, GLP_IV = 1, 2
GLP_FR, GLP_LO, GLP_UP, GLP_DB, GLP_FX = 1, 2, 3, 4, 5
GLP_BS, GLP_NL, GLP_NU, GLP_NF, GLP_NS = 1, 2, 3, 4, 5
GLP_UNDEF, GLP_FEAS, GLP_INFEAS, GLP_NOFEAS, GLP_OPT, GLP_UNBND = 1, 2, 3, 4, 5, 6
GLP_PRIMAL, GLP_DUALP, GLP_DUAL = 1, 2, 3
GLP_PT_STD, GLP_PT_PSE = 0x11, 0x22
GLP_RT_STD, GLP_RT_HAR = 0x11, 0x22
GLP_MSG_OFF = 0
DBL_MAX = 1.7976931348623157e308
INT_MAX = 2147483647

EV_P_CHUZC, EV_P_CHUZR, EV_P_TROW, EV_P_GAMMA, EV_P_ITER = 1, 2, 3, 4, 5
EV_D_CHUZR, EV_D_CHUZC, EV_D_TROW, EV_D_GAMMA, EV_D_ITER = 11, 12, 13, 14, 15


class SMCP(C.Structure):
    _fields_ = [("msg_lev", C.c_int), ("meth", C.c_int), ("pricing", C.c_int),
                ("r_test", C.c_int), ("tol_bnd", C.c_double), ("tol_dj", C.c_double),
                ("tol_piv", C.c_double), ("obj_ll", C.c_double), ("obj_ul", C.c_double),
                ("it_lim", C.c_int), ("tm_lim", C.c_int), ("out_frq", C.c_int),
                ("out_dly", C.c_int), ("presolve", C.c_int)]


class IOCP(C.Structure):
    _fields_ = [("msg_lev", C.c_int), ("br_tech", C.c_int), ("bt_tech", C.c_int),
                ("tol_int", C.c_double), ("tol_obj", C.c_double), ("tm_lim", C.c_int),
                ("out_frq", C.c_int), ("out_dly", C.c_int), ("pp_tech", C.c_int),
                ("mip_gap", C.c_double), ("presolve", C.c_int), ("node_lim", C.c_long)]


HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_void_p)

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "-j8"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.glpo_create.restype = C.c_void_p
        for name in ("glpo_delete", "glpo_std_basis"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.glpo_read_lp.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
        L.glpo_write_lp.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.glpo_simplex.argtypes = [C.c_void_p, C.POINTER(SMCP)]
        L.glpo_intopt.argtypes = [C.c_void_p, C.POINTER(IOCP)]
        L.glpo_factorize.argtypes = [C.c_void_p]
        L.glpo_get_status.argtypes = [C.c_void_p]
        L.glpo_set_hook.argtypes = [C.c_void_p, HOOK, C.c_void_p]
        L.glpo_csa_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
        L.glpo_csa_scalars.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.glpo_set_col_bnds.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
        L.glpo_set_row_bnds.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
        L.glpo_set_bfcp.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_double]
        L.glpo_chuzc_primal.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        L.glpo_chuzr_dual.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_double, C.c_void_p]
        L.glpo_chuzr_primal.argtypes = ([C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p,
                                        C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_double, C.c_void_p, C.c_void_p, C.c_void_p])
        L.glpo_chuzc_dual.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p,
                                      C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.glpo_sort_list.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                     C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


_CSA_DTYPES = {
    "type": np.int8, "orig_type": np.int8, "stat": np.int8, "refsp": np.int8,
    "A_ptr": np.int32, "A_ind": np.int32, "AT_ptr": np.int32, "AT_ind": np.int32,
    "head": np.int32, "bind": np.int32, "N_ptr": np.int32, "N_len": np.int32,
    "N_ind": np.int32, "tcol_ind": np.int32, "trow_ind": np.int32,
}


def csa_get(csa, name):
    """Copy a live CSA array (1-based, slot 0 included) out of a hook."""
    L = lib()
    nbytes = L.glpo_csa_get(csa, name.encode(), None, 0)
    assert nbytes >= 0, name
    dt = np.dtype(_CSA_DTYPES.get(name, np.float64))
    out = np.empty(nbytes // dt.itemsize, dtype=dt)
    L.glpo_csa_get(csa, name.encode(), _p(out), nbytes)
    return out


def csa_scalars(csa):
    iv = np.zeros(16, dtype=np.int32)
    dv = np.zeros(16, dtype=np.float64)
    lib().glpo_csa_scalars(csa, _p(iv), _p(dv))
    keys_i = ["m", "n", "phase", "p", "q", "p_stat", "tcol_nnz", "tcol_num", "trow_nnz",
              "trow_num", "it_cnt", "refct", "nnz"]
    keys_d = ["teta", "delta", "new_dq", "zeta", "tcol_max", "trow_max", "tol"]
    d = {k: int(iv[i]) for i, k in enumerate(keys_i)}
    d.update({k: float(dv[i]) for i, k in enumerate(keys_d)})
    return d


def generate(which, **kw):
    """The synthetic problems of SURVEY 8d built by the ORACLE's own generator
    (oracle/gen.cpp) -- same dict layout as glpk_js_b200.native.generate
    (type/lb/ub are [m+n], rows first), without loading the product library."""
    L = lib()
    L.glpo_gen_packing.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int]
    L.glpo_gen_covering.argtypes = [C.c_int] * 5
    L.glpo_gen_mkp.argtypes = [C.c_int] * 3
    L.glpo_gen_fetch.argtypes = [C.c_void_p] * 12
    if which == "packing":
        m, n = kw.get("m", 2048), kw.get("n", 4096)
        nnz = L.glpo_gen_packing(m, n, kw.get("density", 0.20), kw.get("seed", 20240501))
    elif which == "covering":
        m, n = kw.get("m", 16384), kw.get("n", 32768)
        nnz = L.glpo_gen_covering(m, n, kw.get("kmin", 8), kw.get("kspan", 17), kw.get("seed", 20240601))
    elif which == "mkp":
        m, n = kw.get("m", 30), kw.get("n", 500)
        nnz = L.glpo_gen_mkp(m, n, kw.get("seed", 20240701))
    else:
        raise ValueError(which)
    i32, f64 = (lambda k: np.zeros(k, np.int32)), (lambda k: np.zeros(k, np.float64))
    r_type, r_lb, r_ub = i32(m), f64(m), f64(m)
    c_type, c_lb, c_ub, c_coef, c_kind = i32(n), f64(n), f64(n), f64(n), i32(n)
    A_ptr, A_ind, A_val = i32(n + 1), i32(nnz), f64(nnz)
    dr = C.c_int()
    L.glpo_gen_fetch(C.byref(dr), *[_p(a) for a in (r_type, r_lb, r_ub, c_type, c_lb, c_ub, c_coef, c_kind,
                                                    A_ptr, A_ind, A_val)])
    return dict(m=m, n=n, nnz=nnz, dir=dr.value, c0=0.0, type=np.concatenate([r_type, c_type]),
                lb=np.concatenate([r_lb, c_lb]), ub=np.concatenate([r_ub, c_ub]), coef=c_coef, kind=c_kind,
                A_ptr=A_ptr, A_ind=A_ind, A_val=A_val)


def csa_lu_stats(csa):
    """nnz of the oracle's F, V, H factors right now (inside a hook)"""
    out = (C.c_long * 5)()
    L = lib()
    L.glpo_csa_lu_stats.argtypes = [C.c_void_p, C.c_void_p]
    L.glpo_csa_lu_stats(csa, out)
    return dict(nnz_f=out[0], nnz_v=out[1], nnz_h=out[2], hh_nfs=out[3], n=out[4])


class Problem:
    """A problem held by the oracle."""

    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(self.L.glpo_create())
        self._hook = None

    def __del__(self):
        try:
            self.L.glpo_delete(self.h)
        except Exception:
            pass

    @classmethod
    def from_lp(cls, text):
        self = cls()
        err = C.create_string_buffer(512)
        ret = self.L.glpo_read_lp(self.h, text.encode(), err, 512)
        if ret != 0:
            raise ValueError(err.value.decode())
        return self

    @classmethod
    def from_arrays(cls, d):
        """d: dict with m, n, dir, c0, r_type, r_lb, r_ub, c_type, c_lb, c_ub,
        c_coef, c_kind, A_ptr, A_ind, A_val (0-based CSC, ascending rows)."""
        self = cls()
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        keep = [i32(d["r_type"]), f64(d["r_lb"]), f64(d["r_ub"]), i32(d["c_type"]),
                f64(d["c_lb"]), f64(d["c_ub"]), f64(d["c_coef"]), i32(d["c_kind"]),
                i32(d["A_ptr"]), i32(d["A_ind"]), f64(d["A_val"])]
        self.L.glpo_load.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double] + [C.c_void_p] * 11
        self.L.glpo_load(self.h, int(d["m"]), int(d["n"]), int(d["dir"]), float(d["c0"]),
                         *[_p(a) for a in keep])
        return self

    def dims(self):
        m, n, nnz, dr = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        c0 = C.c_double()
        self.L.glpo_dims(self.h, C.byref(m), C.byref(n), C.byref(nnz), C.byref(dr), C.byref(c0))
        return m.value, n.value, nnz.value, dr.value, c0.value

    def export(self):
        m, n, nnz, dr, c0 = self.dims()
        d = dict(m=m, n=n, dir=dr, c0=c0,
                 r_type=np.zeros(m, np.int32), r_lb=np.zeros(m), r_ub=np.zeros(m),
                 c_type=np.zeros(n, np.int32), c_lb=np.zeros(n), c_ub=np.zeros(n),
                 c_coef=np.zeros(n), c_kind=np.zeros(n, np.int32),
                 A_ptr=np.zeros(n + 1, np.int32), A_ind=np.zeros(nnz, np.int32),
                 A_val=np.zeros(nnz), rii=np.zeros(m), sjj=np.zeros(n))
        self.L.glpo_export.argtypes = [C.c_void_p] * 14
        self.L.glpo_export(self.h, *[_p(d[k]) for k in
                                     ("r_type", "r_lb", "r_ub", "c_type", "c_lb", "c_ub", "c_coef",
                                      "c_kind", "A_ptr", "A_ind", "A_val", "rii", "sjj")])
        return d

    def write_lp(self):
        n = self.L.glpo_write_lp(self.h, None, 0)
        buf = C.create_string_buffer(n + 1)
        self.L.glpo_write_lp(self.h, buf, n + 1)
        return buf.value.decode()

    def std_basis(self):
        self.L.glpo_std_basis(self.h)

    def set_stat(self, stat):
        s = np.ascontiguousarray(stat, dtype=np.int32)
        self.L.glpo_set_stat.argtypes = [C.c_void_p, C.c_void_p]
        self.L.glpo_set_stat(self.h, _p(s))

    def set_col_bnds(self, j, type_, lb, ub):
        self.L.glpo_set_col_bnds(self.h, j, type_, lb, ub)

    def set_bfcp(self, nfs_max=100, piv_tol=0.10, piv_lim=4, upd_tol=1e-6):
        self.L.glpo_set_bfcp(self.h, nfs_max, piv_tol, piv_lim, upd_tol)

    def smcp(self, **kw):
        p = SMCP()
        self.L.glpo_init_smcp(C.byref(p))
        p.msg_lev = GLP_MSG_OFF
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def set_hook(self, fn):
        """fn(event, csa_ptr) or None."""
        if fn is None:
            self._hook = None
            self.L.glpo_set_hook(self.h, C.cast(None, HOOK), None)
            return
        self._hook = HOOK(lambda user, ev, csa: fn(ev, csa))
        self.L.glpo_set_hook(self.h, self._hook, None)

    def simplex(self, parm=None, **kw):
        if parm is None:
            parm = self.smcp(**kw)
        return self.L.glpo_simplex(self.h, C.byref(parm))

    def factorize(self):
        return self.L.glpo_factorize(self.h)

    def intopt(self, **kw):
        p = IOCP()
        self.L.glpo_init_iocp(C.byref(p))
        p.msg_lev = GLP_MSG_OFF
        for k, v in kw.items():
            setattr(p, k, v)
        return self.L.glpo_intopt(self.h, C.byref(p))

    def solution(self):
        m, n, _, _, _ = self.dims()
        stat = np.zeros(m + n, np.int32)
        prim = np.zeros(m + n)
        dual = np.zeros(m + n)
        head = np.zeros(m, np.int32)
        pbs, dbs, it, some = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        obj = C.c_double()
        self.L.glpo_get_solution.argtypes = [C.c_void_p] * 10
        self.L.glpo_get_solution(self.h, _p(stat), _p(prim), _p(dual), _p(head), C.byref(pbs),
                                 C.byref(dbs), C.byref(obj), C.byref(it), C.byref(some))
        return dict(stat=stat, prim=prim, dual=dual, head=head, pbs=pbs.value, dbs=dbs.value,
                    obj=obj.value, it_cnt=it.value, some=some.value,
                    status=self.L.glpo_get_status(self.h), m=m, n=n)

    def mip(self):
        m, n, _, _, _ = self.dims()
        st = C.c_int()
        obj = C.c_double()
        nodes = C.c_long()
        x = np.zeros(m + n)
        self.L.glpo_get_mip.argtypes = [C.c_void_p] * 5
        self.L.glpo_get_mip(self.h, C.byref(st), C.byref(obj), _p(x), C.byref(nodes))
        return dict(mip_stat=st.value, mip_obj=obj.value, mipx=x, nodes=nodes.value)

    def bfd_stats(self):
        out = (C.c_long * 4)()
        self.L.glpo_get_bfd_stats.argtypes = [C.c_void_p, C.c_void_p]
        self.L.glpo_get_bfd_stats(self.h, out)
        return dict(factorize=out[0], update=out[1], ftran=out[2], btran=out[3])
