"""ctypes binding of libglpb200.so (include/glpb200.h).

The library is built in-tree (``glpk.js_b200/libglpb200.so``) by ``build()``
with nvcc for sm_100a.  Nothing here computes on the CPU: a missing library or
a missing device is an error, never a silent fallback.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GLPB_LIB") or os.path.join(HERE, "libglpb200.so")     # GLPB_LIB: another build of the same library (experiments)
CSRC = os.path.join(HERE, "csrc")

# ---- GLP_* constants (lib/glpk.js) ----
GLP_MIN, GLP_MAX = 1, 2
GLP_CV, GLP_IV, GLP_BV = 1, 2, 3
GLP_FR, GLP_LO, GLP_UP, GLP_DB, GLP_FX = 1, 2, 3, 4, 5
GLP_BS, GLP_NL, GLP_NU, GLP_NF, GLP_NS = 1, 2, 3, 4, 5
GLP_UNDEF, GLP_FEAS, GLP_INFEAS, GLP_NOFEAS, GLP_OPT, GLP_UNBND = 1, 2, 3, 4, 5, 6
GLP_MSG_OFF, GLP_MSG_ERR, GLP_MSG_ON, GLP_MSG_ALL, GLP_MSG_DBG = 0, 1, 2, 3, 4
GLP_PRIMAL, GLP_DUALP, GLP_DUAL = 1, 2, 3
GLP_PT_STD, GLP_PT_PSE = 0x11, 0x22
GLP_RT_STD, GLP_RT_HAR = 0x11, 0x22
GLP_BR_FFV, GLP_BR_LFV, GLP_BR_MFV, GLP_BR_DTH, GLP_BR_PCH = 1, 2, 3, 4, 5
GLP_BT_DFS, GLP_BT_BFS, GLP_BT_BLB, GLP_BT_BPH = 1, 2, 3, 4
GLP_PP_NONE, GLP_PP_ROOT, GLP_PP_ALL = 0, 1, 2
GLP_ON, GLP_OFF = 1, 0
GLP_SOL, GLP_IPT, GLP_MIP = 1, 2, 3
GLP_SF_GM, GLP_SF_EQ, GLP_SF_2N, GLP_SF_SKIP, GLP_SF_AUTO = 0x01, 0x10, 0x20, 0x40, 0x80
(GLP_EBADB, GLP_ESING, GLP_ECOND, GLP_EBOUND, GLP_EFAIL, GLP_EOBJLL, GLP_EOBJUL, GLP_EITLIM,
 GLP_ETMLIM, GLP_ENOPFS, GLP_ENODFS, GLP_EROOT, GLP_ESTOP, GLP_EMIPGAP) = range(1, 15)
GLPB_EINVAL, GLPB_ENODEV, GLPB_ENOMEM, GLPB_ESTATE = -1, -2, -3, -4
DBL_MAX = 1.7976931348623157e308
INT_MAX = 2147483647


class glpb_smcp(C.Structure):
    _fields_ = [("msg_lev", C.c_int), ("meth", C.c_int), ("pricing", C.c_int), ("r_test", C.c_int),
                ("tol_bnd", C.c_double), ("tol_dj", C.c_double), ("tol_piv", C.c_double),
                ("obj_ll", C.c_double), ("obj_ul", C.c_double), ("it_lim", C.c_int),
                ("tm_lim", C.c_int), ("out_frq", C.c_int), ("out_dly", C.c_int), ("presolve", C.c_int)]


class glpb_iocp(C.Structure):
    _fields_ = [("msg_lev", C.c_int), ("br_tech", C.c_int), ("bt_tech", C.c_int),
                ("tol_int", C.c_double), ("tol_obj", C.c_double), ("tm_lim", C.c_int),
                ("out_frq", C.c_int), ("out_dly", C.c_int), ("pp_tech", C.c_int),
                ("mip_gap", C.c_double), ("presolve", C.c_int), ("node_lim", C.c_long)]


class glpb_bfcp(C.Structure):
    _fields_ = [("nfs_max", C.c_int), ("piv_tol", C.c_double), ("upd_tol", C.c_double)]


class glpb_problem_data(C.Structure):
    _fields_ = [("m", C.c_int), ("n", C.c_int), ("nnz", C.c_int), ("dir", C.c_int), ("c0", C.c_double),
                ("type", C.POINTER(C.c_int)), ("lb", C.POINTER(C.c_double)), ("ub", C.POINTER(C.c_double)),
                ("coef", C.POINTER(C.c_double)), ("kind", C.POINTER(C.c_int)),
                ("A_ptr", C.POINTER(C.c_int)), ("A_ind", C.POINTER(C.c_int)), ("A_val", C.POINTER(C.c_double))]


# every symbol include/glpb200.h declares
SYMBOLS = [
    "glpb_init_smcp", "glpb_init_iocp", "glpb_device_count", "glpb_last_error", "glpb_version",
    "glpb_create", "glpb_destroy", "glpb_set_bounds", "glpb_set_basis", "glpb_std_basis",
    "glpb_set_bfcp", "glpb_set_it_cnt", "glpb_factorize", "glpb_simplex", "glpb_intopt",
    "glpb_get_solution", "glpb_get_status", "glpb_get_mip", "glpb_get_counters", "glpb_ftran",
    "glpb_set_profile", "glpb_profile_report",
    "glpb_mip_begin", "glpb_mip_run", "glpb_mip_get_incumbent", "glpb_mip_set_cutoff", "glpb_mip_open_count",
    "glpb_mip_record_bytes", "glpb_mip_export_nodes", "glpb_mip_import_nodes", "glpb_mip_end",
    "glpb_btran", "glpb_k_chuzc_primal", "glpb_k_chuzr_dual", "glpb_k_ratio_primal",
    "glpb_k_ratio_dual", "glpb_k_trow", "glpb_k_sort_list", "glpb_bench_kernel", "glpb_gen_packing",
    "glpb_gen_covering", "glpb_gen_mkp", "glpb_free_problem", "glpb_rng_fill",
    "glpb_scale_prob", "glpb_adv_basis", "glpb_read_lp", "glpb_free_names", "glpb_write_lp",
    "glpb_set_pivot_log", "glpb_get_pivot_log", "glpb_debug_get", "glpb_bnb_begin", "glpb_bnb_round", "glpb_bnb_open_count", "glpb_bnb_get_incumbent", "glpb_bnb_set_cutoff", "glpb_bnb_clear",
    "glpb_bnb_record_bytes", "glpb_bnb_export_nodes", "glpb_bnb_import_nodes", "glpb_bnb_stats", "glpb_bnb_end",
    "glpb_npp_create", "glpb_npp_destroy", "glpb_npp_load_prob", "glpb_npp_simplex", "glpb_npp_integer",
    "glpb_npp_get_counts", "glpb_npp_get_size", "glpb_npp_build_prob", "glpb_npp_postprocess",
]

_lib = None


def build(verbose=False):
    """Compile libglpb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode != 0:
        raise RuntimeError("building libglpb200.so failed")
    return LIB_PATH


def load():
    """Load the CUDA library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "glpk.js_b200: %s is missing -- run __graft_entry__.build(); there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    L.glpb_last_error.restype = C.c_char_p
    L.glpb_version.restype = C.c_char_p
    L.glpb_create.restype = vp
    L.glpb_create.argtypes = [ci, ci, ci, ci, cd] + [vp] * 10 + [ci]
    L.glpb_destroy.argtypes = [vp]
    L.glpb_set_bounds.argtypes = [vp, ci, vp, vp, vp, vp]
    L.glpb_set_basis.argtypes = [vp, vp]
    L.glpb_std_basis.argtypes = [vp]
    L.glpb_set_bfcp.argtypes = [vp, vp]
    L.glpb_set_it_cnt.argtypes = [vp, ci]
    L.glpb_factorize.argtypes = [vp]
    L.glpb_simplex.argtypes = [vp, vp]
    L.glpb_intopt.argtypes = [vp, vp]
    L.glpb_get_solution.argtypes = [vp] * 10
    L.glpb_get_status.argtypes = [vp]
    L.glpb_get_mip.argtypes = [vp] * 5
    L.glpb_get_counters.argtypes = [vp, vp, ci]
    L.glpb_ftran.argtypes = [vp, vp]
    L.glpb_set_profile.argtypes = [vp, ci]
    L.glpb_profile_report.argtypes = [vp]
    L.glpb_profile_report.restype = C.c_char_p
    L.glpb_btran.argtypes = [vp, vp]
    L.glpb_mip_begin.argtypes = [vp, vp]
    L.glpb_mip_run.argtypes = [vp, C.c_long, vp]
    L.glpb_mip_get_incumbent.argtypes = [vp, vp, vp]
    L.glpb_mip_set_cutoff.argtypes = [vp, cd]
    L.glpb_mip_open_count.argtypes = [vp]
    L.glpb_mip_record_bytes.argtypes = [vp]
    L.glpb_mip_record_bytes.restype = C.c_long
    L.glpb_mip_export_nodes.argtypes = [vp, ci, vp, C.c_long, vp]
    L.glpb_mip_import_nodes.argtypes = [vp, vp, ci]
    L.glpb_mip_end.argtypes = [vp, ci]
    L.glpb_k_chuzc_primal.argtypes = [ci, vp, vp, vp, cd, vp]
    L.glpb_k_chuzr_dual.argtypes = [ci, ci] + [vp] * 6 + [cd, vp, vp]
    L.glpb_k_ratio_primal.argtypes = [ci, ci] + [vp] * 5 + [ci, vp, cd, ci, vp, vp, ci, cd, vp, vp, vp]
    L.glpb_k_ratio_dual.argtypes = [ci, vp, vp, cd, vp, vp, ci, cd, vp, vp]
    L.glpb_k_trow.argtypes = [ci, ci] + [vp] * 7
    L.glpb_k_sort_list.argtypes = [ci, vp, cd, vp, vp]
    L.glpb_bench_kernel.argtypes = [C.c_char_p, ci, ci, ci, vp, vp]
    L.glpb_gen_packing.argtypes = [ci, ci, cd, ci, vp]
    L.glpb_gen_covering.argtypes = [ci, ci, ci, ci, ci, vp]
    L.glpb_gen_mkp.argtypes = [ci, ci, ci, vp]
    L.glpb_free_problem.argtypes = [vp]
    L.glpb_rng_fill.argtypes = [ci, ci, vp]
    L.glpb_scale_prob.argtypes = [ci, ci, vp, vp, vp, ci, vp, vp, vp]
    L.glpb_adv_basis.argtypes = [ci, ci] + [vp] * 9
    L.glpb_read_lp.argtypes = [C.c_char_p, C.c_long, vp, vp, vp]
    L.glpb_free_names.argtypes = [vp]
    L.glpb_write_lp.argtypes = [ci, ci, ci, cd] + [vp] * 9 + [C.c_char_p, C.c_char_p, vp, vp, vp]
    L.glpb_set_pivot_log.argtypes = [vp, ci]
    L.glpb_get_pivot_log.argtypes = [vp, vp, ci, vp]
    L.glpb_debug_get.argtypes = [vp, C.c_char_p, vp, ci]
    L.glpb_bnb_begin.argtypes = [vp, vp, ci, ci]
    L.glpb_bnb_round.argtypes = [vp, C.c_long, vp]
    L.glpb_bnb_open_count.argtypes = [vp]
    L.glpb_bnb_get_incumbent.argtypes = [vp, vp, vp]
    L.glpb_bnb_set_cutoff.argtypes = [vp, cd]
    L.glpb_bnb_clear.argtypes = [vp]
    L.glpb_bnb_record_bytes.argtypes = [vp]
    L.glpb_bnb_record_bytes.restype = C.c_long
    L.glpb_bnb_export_nodes.argtypes = [vp, ci, vp, vp]
    L.glpb_bnb_import_nodes.argtypes = [vp, vp, ci]
    L.glpb_bnb_stats.argtypes = [vp, vp, ci]
    L.glpb_bnb_end.argtypes = [vp, ci]
    L.glpb_npp_create.argtypes = []
    L.glpb_npp_create.restype = vp
    L.glpb_npp_destroy.argtypes = [vp]
    L.glpb_npp_destroy.restype = None
    L.glpb_npp_load_prob.argtypes = [vp, ci, ci, ci, cd] + [vp] * 8 + [ci]
    L.glpb_npp_simplex.argtypes = [vp]
    L.glpb_npp_integer.argtypes = [vp, ci]
    L.glpb_npp_get_counts.argtypes = [vp, vp, ci]
    L.glpb_npp_get_size.argtypes = [vp, vp, vp, vp]
    L.glpb_npp_build_prob.argtypes = [vp] * 12
    L.glpb_npp_postprocess.argtypes = [vp] * 9
    _lib = L
    return L


def last_error():
    return load().glpb_last_error().decode()


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i8(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int8)


def generate(which, **kw):
    """Synthetic problems of SURVEY 8d: 'packing' (C2), 'covering' (C3), 'mkp' (C5).
    Returns a dict of numpy arrays in the layout glpb_create takes."""
    L = load()
    pd = glpb_problem_data()
    if which == "packing":
        rc = L.glpb_gen_packing(kw.get("m", 2048), kw.get("n", 4096), kw.get("density", 0.20),
                                kw.get("seed", 20240501), C.byref(pd))
    elif which == "covering":
        rc = L.glpb_gen_covering(kw.get("m", 16384), kw.get("n", 32768), kw.get("kmin", 8),
                                 kw.get("kspan", 17), kw.get("seed", 20240601), C.byref(pd))
    elif which == "mkp":
        rc = L.glpb_gen_mkp(kw.get("m", 30), kw.get("n", 500), kw.get("seed", 20240701), C.byref(pd))
    else:
        raise ValueError(which)
    if rc != 0:
        raise ValueError("generator %s: invalid parameters" % which)
    m, n, nnz = pd.m, pd.n, pd.nnz
    arr = lambda p, k, dt: np.ctypeslib.as_array(p, shape=(max(k, 1),))[:k].astype(dt).copy()
    d = dict(m=m, n=n, nnz=nnz, dir=pd.dir, c0=pd.c0,
             type=arr(pd.type, m + n, np.int32), lb=arr(pd.lb, m + n, np.float64),
             ub=arr(pd.ub, m + n, np.float64), coef=arr(pd.coef, n, np.float64),
             kind=arr(pd.kind, n, np.int32), A_ptr=arr(pd.A_ptr, n + 1, np.int32),
             A_ind=arr(pd.A_ind, nnz, np.int32), A_val=arr(pd.A_val, nnz, np.float64))
    L.glpb_free_problem(C.byref(pd))
    return d


def _problem_arrays(pd):
    m, n, nnz = pd.m, pd.n, pd.nnz
    arr = lambda p, k, dt: np.ctypeslib.as_array(p, shape=(max(k, 1),))[:k].astype(dt).copy()
    return dict(m=m, n=n, nnz=nnz, dir=pd.dir, c0=pd.c0,
                type=arr(pd.type, m + n, np.int32), lb=arr(pd.lb, m + n, np.float64),
                ub=arr(pd.ub, m + n, np.float64), coef=arr(pd.coef, n, np.float64),
                kind=arr(pd.kind, n, np.int32), A_ptr=arr(pd.A_ptr, n + 1, np.int32),
                A_ind=arr(pd.A_ind, nnz, np.int32), A_val=arr(pd.A_val, nnz, np.float64))


def read_lp(text):
    """glpb_read_lp: CPLEX LP text -> (arrays in the layout glpb_create takes, names) where
    names = dict(obj=..., rows=[...], cols=[...]).  Raises ValueError with the reader's
    message (line number included) on a syntax error.  Host only."""
    L = load()
    raw = text.encode() if isinstance(text, str) else bytes(text)
    pd = glpb_problem_data()
    names, nlen = C.c_void_p(), C.c_long()
    rc = L.glpb_read_lp(raw, len(raw), C.byref(pd), C.byref(names), C.byref(nlen))
    if rc != 0:
        raise ValueError(last_error() if rc == 1 else "glpb_read_lp failed (%d)" % rc)
    d = _problem_arrays(pd)
    blob = C.string_at(names, nlen.value).decode()
    L.glpb_free_names(names)
    L.glpb_free_problem(C.byref(pd))
    parts = blob.split("\0")[:-1]
    return d, dict(obj=parts[0], rows=parts[1:1 + d["m"]], cols=parts[1 + d["m"]:1 + d["m"] + d["n"]])


def write_lp(d, col_len, R_ptr, R_ind, R_val, prob_name=None, names=None):
    """glpb_write_lp: (lines, count) of the CPLEX LP text of problem `d` (layout of Problem, unscaled); rows
    in list order; names = (obj, [rows], [cols]) with None for a missing name.  Host only."""
    L = load()
    m, n = int(d["m"]), int(d["n"])
    keep = [_i32(d["type"]), _f64(d["lb"]), _f64(d["ub"]), _f64(d["coef"]), _i32(d.get("kind")), _i32(col_len),
            _i32(R_ptr), _i32(R_ind), _f64(R_val)]
    blob = None
    if names is not None:
        obj, rows, cols = names
        blob = b"".join((x or "").encode() + b"\0" for x in [obj] + list(rows) + list(cols))
    text, tlen, count = C.c_void_p(), C.c_long(), C.c_int()
    rc = L.glpb_write_lp(m, n, int(d["dir"]), float(d["c0"]), *[_p(a) for a in keep],
                         None if prob_name is None else prob_name.encode(), blob,
                         C.byref(text), C.byref(tlen), C.byref(count))
    if rc != 0:
        raise ValueError("glpb_write_lp: invalid arguments (rc=%d)" % rc)
    raw = C.string_at(text, tlen.value).decode()
    L.glpb_free_names(text)
    return raw.split("\n")[:-1], count.value


def scale_prob(m, n, A_ptr, A_ind, A_val, flags):
    """glpb_scale_prob: scale factors (rii[m], sjj[n]) and the per-stage report
    {stage: (min|aij|, max|aij|, ratio)} of lib/glpscl.js scale_prob.  Host only."""
    L = load()
    A_ptr, A_ind, A_val = _i32(A_ptr), _i32(A_ind), _f64(A_val)
    rii, sjj, rep = np.ones(m), np.ones(n), np.zeros(13)
    rc = L.glpb_scale_prob(m, n, _p(A_ptr), _p(A_ind), _p(A_val), int(flags), _p(rii), _p(sjj), _p(rep))
    if rc != 0:
        raise ValueError("glpb_scale_prob: invalid arguments (rc=%d)" % rc)
    mask = int(rep[0])
    report = {name: tuple(rep[1 + 3 * s:4 + 3 * s]) for s, name in enumerate(("A", "GM", "EQ", "2N"))
              if mask & (1 << s)}
    report["skipped"] = bool(mask & 16)
    return rii, sjj, report


def adv_basis(m, n, A_ptr, A_ind, R_ptr, R_ind, type_, lb, ub):
    """glpb_adv_basis: statuses [m+n] of the triangular crash basis and the size
    of the triangular part (lib/glpini01.js).  Host only."""
    L = load()
    A_ptr, A_ind, R_ptr, R_ind = _i32(A_ptr), _i32(A_ind), _i32(R_ptr), _i32(R_ind)
    type_, lb, ub = _i32(type_), _f64(lb), _f64(ub)
    stat = np.zeros(m + n, np.int32)
    size = C.c_int(0)
    rc = L.glpb_adv_basis(m, n, _p(A_ptr), _p(A_ind), _p(R_ptr), _p(R_ind), _p(type_), _p(lb), _p(ub),
                          _p(stat), C.byref(size))
    if rc != 0:
        raise ValueError("glpb_adv_basis: invalid arguments (rc=%d)" % rc)
    return stat, size.value


class Presolver:
    """One presolver workspace (glpb_npp_*, csrc/presolve.cpp): host only.  `d` as for Problem -- type/lb/ub
    [m+n], coef, kind, dir, c0 and the matrix by columns -- but unscaled and with every column's elements in
    the reference's list order."""

    KINDS = ("free_row", "fixed_col", "make_equality", "make_fixed", "empty_col", "eq_singlet", "ineq_singlet",
             "implied_slack", "implied_free", "forcing_row", "inactive_bound", "lbnd_col", "binarize")

    def __init__(self, d, sol):
        L = load()
        self.L, self.sol = L, int(sol)
        self.m, self.n = int(d["m"]), int(d["n"])
        self.h = C.c_void_p(L.glpb_npp_create())
        if not self.h:
            raise MemoryError("glpb_npp_create")
        a = [_i32(d["type"]), _f64(d["lb"]), _f64(d["ub"]), _f64(d["coef"]), _i32(d.get("kind")),
             _i32(d["A_ptr"]), _i32(d["A_ind"]), _f64(d["A_val"])]
        rc = L.glpb_npp_load_prob(self.h, self.m, self.n, int(d["dir"]), float(d["c0"]), *[_p(x) for x in a],
                                  self.sol)
        if rc != 0:
            self.close()
            raise ValueError("glpb_npp_load_prob: invalid arguments (rc=%d)" % rc)

    def close(self):
        if getattr(self, "h", None):
            self.L.glpb_npp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def simplex(self):
        """npp_simplex: 0 / GLP_ENOPFS / GLP_ENODFS"""
        rc = self.L.glpb_npp_simplex(self.h)
        if rc < 0:
            raise ValueError("glpb_npp_simplex (rc=%d)" % rc)
        return rc

    def integer(self, binarize=False):
        """npp_integer: 0 / GLP_ENOPFS / GLP_ENODFS"""
        rc = self.L.glpb_npp_integer(self.h, 1 if binarize else 0)
        if rc < 0:
            raise ValueError("glpb_npp_integer (rc=%d)" % rc)
        return rc

    def counts(self):
        out = np.zeros(21, np.int32)
        self.L.glpb_npp_get_counts(self.h, _p(out), 21)
        return dict(zip(("packing", "covering", "reduced", "bin_vars", "bin_bins", "bin_rows", "bin_fails",
                         "stack") + self.KINDS, (int(x) for x in out)))

    def build(self):
        """npp_build_prob: the reduced problem as a dict in the layout of Problem (+ row_ref / col_ref)."""
        m, n, nz = C.c_int(0), C.c_int(0), C.c_int(0)
        self.L.glpb_npp_get_size(self.h, C.byref(m), C.byref(n), C.byref(nz))
        m, n, nz = m.value, n.value, nz.value
        c0 = C.c_double(0.0)
        type_, lb, ub = np.zeros(m + n, np.int32), np.zeros(m + n), np.zeros(m + n)
        coef, kind = np.zeros(n), np.ones(n, np.int32)
        A_ptr, A_ind, A_val = np.zeros(n + 1, np.int32), np.zeros(nz, np.int32), np.zeros(nz)
        row_ref, col_ref = np.zeros(m, np.int32), np.zeros(n, np.int32)
        rc = self.L.glpb_npp_build_prob(self.h, C.byref(c0), _p(type_), _p(lb), _p(ub), _p(coef), _p(kind),
                                        _p(A_ptr), _p(A_ind), _p(A_val), _p(row_ref), _p(col_ref))
        if rc != 0:
            raise ValueError("glpb_npp_build_prob (rc=%d)" % rc)
        self.rm, self.rn = m, n
        return dict(m=m, n=n, c0=c0.value, type=type_, lb=lb, ub=ub, coef=coef, kind=kind, A_ptr=A_ptr,
                    A_ind=A_ind, A_val=A_val, row_ref=row_ref, col_ref=col_ref)

    def postprocess(self, c_value, r_stat=None, r_dual=None, c_stat=None):
        """npp_postprocess: (r_stat, r_dual, c_stat, c_value) of the ORIGINAL problem (MIP: the first three
        are None)."""
        basic = self.sol == 1
        cv = _f64(c_value)
        o_cv = np.zeros(self.n)
        if basic:
            rs, rd, cs = _i32(r_stat), _f64(r_dual), _i32(c_stat)
            o_rs, o_rd, o_cs = np.zeros(self.m, np.int32), np.zeros(self.m), np.zeros(self.n, np.int32)
            rc = self.L.glpb_npp_postprocess(self.h, _p(rs), _p(rd), _p(cs), _p(cv), _p(o_rs), _p(o_rd),
                                             _p(o_cs), _p(o_cv))
        else:
            o_rs = o_rd = o_cs = None
            rc = self.L.glpb_npp_postprocess(self.h, None, None, None, _p(cv), None, None, None, _p(o_cv))
        if rc != 0:
            raise RuntimeError("glpb_npp_postprocess: solution cannot be recovered (rc=%d)" % rc)
        return o_rs, o_rd, o_cs, o_cv


class Problem:
    """A device-resident problem (one glpb_prob handle)."""

    def __init__(self, d, device=0, rii=None, sjj=None):
        L = load()
        self.L = L
        self.m, self.n, self.nnz = int(d["m"]), int(d["n"]), int(len(d["A_val"]))
        keep = [_i32(d["type"]), _f64(d["lb"]), _f64(d["ub"]), _f64(d["coef"]), _i32(d.get("kind")),
                _f64(rii), _f64(sjj), _i32(d["A_ptr"]), _i32(d["A_ind"]), _f64(d["A_val"])]
        h = L.glpb_create(self.m, self.n, self.nnz, int(d["dir"]), float(d["c0"]),
                          *[_p(a) for a in keep], device)
        if not h:
            raise RuntimeError("glpb_create failed: " + last_error())
        self.h = C.c_void_p(h)

    def close(self):
        if getattr(self, "h", None):
            self.L.glpb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def smcp(self, **kw):
        p = glpb_smcp()
        self.L.glpb_init_smcp(C.byref(p))
        p.msg_lev = GLP_MSG_OFF
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def iocp(self, **kw):
        p = glpb_iocp()
        self.L.glpb_init_iocp(C.byref(p))
        p.msg_lev = GLP_MSG_OFF
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def std_basis(self):
        return self.L.glpb_std_basis(self.h)

    def set_basis(self, stat):
        s = _i32(stat)
        return self.L.glpb_set_basis(self.h, _p(s))

    def set_bounds(self, k, type_, lb, ub):
        k, type_, lb, ub = _i32(k), _i32(type_), _f64(lb), _f64(ub)
        return self.L.glpb_set_bounds(self.h, len(k), _p(k), _p(type_), _p(lb), _p(ub))

    def set_bfcp(self, nfs_max=100, piv_tol=0.10, upd_tol=1e-6):
        b = glpb_bfcp(nfs_max, piv_tol, upd_tol)
        return self.L.glpb_set_bfcp(self.h, C.byref(b))

    def factorize(self):
        return self.L.glpb_factorize(self.h)

    def simplex(self, parm=None, **kw):
        if parm is None:
            parm = self.smcp(**kw)
        rc = self.L.glpb_simplex(self.h, C.byref(parm))
        if rc < 0:
            raise RuntimeError("glpb_simplex failed (%d): %s" % (rc, last_error()))
        return rc

    def intopt(self, parm=None, **kw):
        if parm is None:
            parm = self.iocp(**kw)
        rc = self.L.glpb_intopt(self.h, C.byref(parm))
        if rc < 0:
            raise RuntimeError("glpb_intopt failed (%d): %s" % (rc, last_error()))
        return rc

    def solution(self):
        m, n = self.m, self.n
        stat = np.zeros(m + n, np.int32)
        prim = np.zeros(m + n)
        dual = np.zeros(m + n)
        head = np.zeros(m, np.int32)
        pbs, dbs, it, some = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        obj = C.c_double()
        self.L.glpb_get_solution(self.h, _p(stat), _p(prim), _p(dual), _p(head), C.byref(pbs),
                                 C.byref(dbs), C.byref(obj), C.byref(it), C.byref(some))
        return dict(stat=stat, prim=prim, dual=dual, head=head, pbs=pbs.value, dbs=dbs.value,
                    obj=obj.value, it_cnt=it.value, some=some.value,
                    status=self.L.glpb_get_status(self.h), m=m, n=n)

    def mip(self):
        st = C.c_int()
        obj = C.c_double()
        nodes = C.c_long()
        x = np.zeros(self.m + self.n)
        self.L.glpb_get_mip(self.h, C.byref(st), C.byref(obj), _p(x), C.byref(nodes))
        return dict(mip_stat=st.value, mip_obj=obj.value, mipx=x, nodes=nodes.value)

    # ---- resumable branch-and-bound (multi-GPU node sharding) ----
    def mip_begin(self, parm=None, **kw):
        if parm is None:
            parm = self.iocp(**kw)
        return self.L.glpb_mip_begin(self.h, C.byref(parm))

    def mip_run(self, max_nodes=-1):
        solved = C.c_long()
        rc = self.L.glpb_mip_run(self.h, max_nodes, C.byref(solved))
        if rc < 0:
            raise RuntimeError("glpb_mip_run failed (%d): %s" % (rc, last_error()))
        return rc, solved.value

    def mip_incumbent(self):
        has = C.c_int()
        obj = C.c_double()
        self.L.glpb_mip_get_incumbent(self.h, C.byref(has), C.byref(obj))
        return bool(has.value), obj.value

    def mip_set_cutoff(self, obj):
        return self.L.glpb_mip_set_cutoff(self.h, float(obj))

    def mip_open_count(self):
        return self.L.glpb_mip_open_count(self.h)

    def mip_export(self, max_count):
        rb = self.L.glpb_mip_record_bytes(self.h)
        room = max_count if max_count > 0 else self.mip_open_count()
        buf = np.zeros(max(1, room) * rb, np.uint8)
        cnt = C.c_int()
        rc = self.L.glpb_mip_export_nodes(self.h, max_count, _p(buf), buf.nbytes, C.byref(cnt))
        if rc != 0:
            raise RuntimeError("glpb_mip_export_nodes failed (%d)" % rc)
        return buf[:cnt.value * rb], cnt.value

    def mip_import(self, buf, count):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        rc = self.L.glpb_mip_import_nodes(self.h, _p(buf), int(count))
        if rc != 0:
            raise RuntimeError("glpb_mip_import_nodes failed (%d)" % rc)

    def mip_end(self, ret=0):
        return self.L.glpb_mip_end(self.h, int(ret))

    # ---- batched, device-resident branch-and-bound (one CTA per node) ----
    def bnb_begin(self, parm=None, batch=0, slab_nodes=0, **kw):
        if parm is None:
            parm = self.iocp(**kw)
        rc = self.L.glpb_bnb_begin(self.h, C.byref(parm), int(batch), int(slab_nodes))
        if rc < 0:
            raise RuntimeError("glpb_bnb_begin failed (%d): %s" % (rc, last_error()))
        return rc

    def bnb_round(self, max_tasks=-1):
        done = C.c_long()
        rc = self.L.glpb_bnb_round(self.h, max_tasks, C.byref(done))
        if rc < 0:
            raise RuntimeError("glpb_bnb_round failed (%d): %s" % (rc, last_error()))
        return rc, done.value

    def bnb_open_count(self):
        return self.L.glpb_bnb_open_count(self.h)

    def bnb_incumbent(self):
        has, obj = C.c_int(), C.c_double()
        self.L.glpb_bnb_get_incumbent(self.h, C.byref(has), C.byref(obj))
        return bool(has.value), obj.value

    def bnb_set_cutoff(self, obj):
        return self.L.glpb_bnb_set_cutoff(self.h, float(obj))

    def bnb_clear(self):
        return self.L.glpb_bnb_clear(self.h)

    def bnb_record_bytes(self):
        return int(self.L.glpb_bnb_record_bytes(self.h))

    def bnb_export(self, max_count, dev_ptr):
        """pack up to max_count open nodes into DEVICE memory at dev_ptr; returns the count"""
        cnt = C.c_int()
        rc = self.L.glpb_bnb_export_nodes(self.h, int(max_count), C.c_void_p(dev_ptr), C.byref(cnt))
        if rc != 0:
            raise RuntimeError("glpb_bnb_export_nodes failed (%d): %s" % (rc, last_error()))
        return cnt.value

    def bnb_import(self, dev_ptr, count):
        rc = self.L.glpb_bnb_import_nodes(self.h, C.c_void_p(dev_ptr), int(count))
        if rc != 0:
            raise RuntimeError("glpb_bnb_import_nodes failed (%d): %s" % (rc, last_error()))

    def bnb_stats(self):
        out = (C.c_long * 14)()
        self.L.glpb_bnb_stats(self.h, out, 14)
        keys = ["solved", "tasks", "rounds", "iters", "refacs", "open", "smem_bytes", "a_in_smem",
                "cyc_preprocess", "cyc_invert", "cyc_fresh", "cyc_iterate", "cyc_branch", "cyc_state_io"]
        return {k: int(out[i]) for i, k in enumerate(keys)}

    def bnb_end(self, ret=0):
        return self.L.glpb_bnb_end(self.h, int(ret))

    def debug_get(self, name, count):
        out = np.zeros(max(1, count))
        rc = self.L.glpb_debug_get(self.h, name.encode(), _p(out), int(count))
        if rc != 0:
            raise RuntimeError("glpb_debug_get(%s) failed (%d)" % (name, rc))
        return out[:count]

    def set_pivot_log(self, cap):
        rc = self.L.glpb_set_pivot_log(self.h, int(cap))
        if rc != 0:
            raise RuntimeError("glpb_set_pivot_log failed (%d): %s" % (rc, last_error()))

    def pivot_log(self, cap):
        """[(q, p)] of the first iterations as the reference numbers them (p = -1: bound flip)"""
        qp = np.zeros(2 * max(1, cap), np.int32)
        cnt = C.c_int()
        rc = self.L.glpb_get_pivot_log(self.h, _p(qp), int(cap), C.byref(cnt))
        if rc != 0:
            raise RuntimeError("glpb_get_pivot_log failed (%d): %s" % (rc, last_error()))
        return [(int(qp[2 * i]), int(qp[2 * i + 1])) for i in range(cnt.value)]

    def counters(self):
        out = (C.c_long * 9)()
        self.L.glpb_get_counters(self.h, out, 9)
        keys = ["iterations", "refactorizations", "launches", "syncs", "updates", "k", "solve_us", "graph_launches", "ties"]
        return {k: int(out[i]) for i, k in enumerate(keys)}

    def set_profile(self, on):
        self.L.glpb_set_profile(self.h, int(on))

    def profile(self):
        """{kernel: dict(count, ms, bytes)} from CUDA events on the solve stream"""
        out = {}
        for line in self.L.glpb_profile_report(self.h).decode().splitlines():
            name, cnt, ms, nb = line.split()
            out[name] = dict(count=int(cnt), ms=float(ms), bytes=float(nb))
        return out

    def ftran(self, x):
        x = _f64(x).copy()
        rc = self.L.glpb_ftran(self.h, _p(x))
        if rc != 0:
            raise RuntimeError("glpb_ftran failed (%d): %s" % (rc, last_error()))
        return x

    def btran(self, x):
        x = _f64(x).copy()
        rc = self.L.glpb_btran(self.h, _p(x))
        if rc != 0:
            raise RuntimeError("glpb_btran failed (%d): %s" % (rc, last_error()))
        return x


# ---- kernel-level entry points (CSA layout: 1-based numpy arrays) ----

def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, last_error()))


def k_chuzc_primal(n, stat, cbar, gamma, tol_dj):
    q = C.c_int()
    stat, cbar, gamma = _i8(stat), _f64(cbar), _f64(gamma)
    _check(load().glpb_k_chuzc_primal(n, _p(stat), _p(cbar), _p(gamma), tol_dj, C.byref(q)), "k_chuzc_primal")
    return q.value


def k_chuzr_dual(m, n, type_, lb, ub, head, bbar, gamma, tol_bnd):
    p = C.c_int()
    delta = C.c_double()
    a = [_i8(type_), _f64(lb), _f64(ub), _i32(head), _f64(bbar), _f64(gamma)]
    _check(load().glpb_k_chuzr_dual(m, n, *[_p(x) for x in a], tol_bnd, C.byref(p), C.byref(delta)), "k_chuzr_dual")
    return p.value, delta.value


def k_ratio_primal(m, n, type_, lb, ub, coef, head, phase, bbar, cbar_q, q, tcol_ind, tcol_vec,
                   tcol_num, rtol):
    p, p_stat = C.c_int(), C.c_int()
    teta = C.c_double()
    a = [_i8(type_), _f64(lb), _f64(ub), _f64(coef), _i32(head)]
    bbar, tcol_ind, tcol_vec = _f64(bbar), _i32(tcol_ind), _f64(tcol_vec)
    _check(load().glpb_k_ratio_primal(m, n, *[_p(x) for x in a], phase, _p(bbar), cbar_q, q, _p(tcol_ind),
                                      _p(tcol_vec), tcol_num, rtol, C.byref(p), C.byref(p_stat),
                                      C.byref(teta)), "k_ratio_primal")
    return p.value, p_stat.value, teta.value


def k_ratio_dual(n, stat, cbar, delta, trow_ind, trow_vec, trow_num, rtol):
    q = C.c_int()
    new_dq = C.c_double()
    stat, cbar, trow_ind, trow_vec = _i8(stat), _f64(cbar), _i32(trow_ind), _f64(trow_vec)
    _check(load().glpb_k_ratio_dual(n, _p(stat), _p(cbar), delta, _p(trow_ind), _p(trow_vec), trow_num,
                                    rtol, C.byref(q), C.byref(new_dq)), "k_ratio_dual")
    return q.value, new_dq.value


def k_trow(m, n, A_ptr, A_ind, A_val, head, stat, rho):
    a = [_i32(A_ptr), _i32(A_ind), _f64(A_val), _i32(head), _i8(stat), _f64(rho)]
    out = np.zeros(1 + n)
    _check(load().glpb_k_trow(m, n, *[_p(x) for x in a], _p(out)), "k_trow")
    return out


def k_sort_list(n, vec, eps):
    """sort_tcol / sort_trow: indices (1-based) of the entries of vec[1..n] with |v| >= eps in the order the
    reference's swap loop leaves them"""
    vec = _f64(vec)
    lst = np.zeros(1 + n, np.int32)
    num = C.c_int()
    _check(load().glpb_k_sort_list(n, _p(vec), eps, _p(lst), C.byref(num)), "k_sort_list")
    return lst[1:1 + num.value].copy()


def bench_kernel(name, m, n, reps=50):
    usec, nbytes = C.c_double(), C.c_double()
    _check(load().glpb_bench_kernel(name.encode(), m, n, reps, C.byref(usec), C.byref(nbytes)), "bench_kernel")
    return usec.value, nbytes.value


def rng_fill(seed, count):
    out = np.zeros(count, np.int32)
    load().glpb_rng_fill(seed, count, _p(out))
    return out
