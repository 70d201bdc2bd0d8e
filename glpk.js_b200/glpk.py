"""Host-side mirror of the glpk.js API for the simplex path.

Same names, argument meaning, 1-based indexing and error behaviour as the
reference's JavaScript (``lib/glpapi01.js``, ``glpapi05.js``, ``glpapi06.js``,
``glpapi09.js``, ``glpcpx.js``): invalid arguments raise (the reference's
``xerror`` throws), solver conditions are integer return codes.  The problem
object lives on the host exactly like the reference's ``glp_prob``; the solve
calls marshal it once into a device-resident handle (``native.Problem``) and
write the solution back into the same fields the reference's getters read.

Scaling (``glp_scale_prob``, lib/glpscl.js), the triangular crash basis
(``glp_adv_basis``, lib/glpini01.js) and the LP-format reader (``glp_read_lp``,
lib/glpcpx.js) run in the native library's host code (``glpb_scale_prob`` /
``glpb_adv_basis`` / ``glpb_read_lp``); they prepare inputs of the path.  The
basis-factorisation interface of lib/glpapi12.js (``glp_factorize``,
``glp_ftran/btran``, ``glp_warm_up``, simplex-table rows/columns, textbook ratio
tests) sits on the device's two basis solves.  Of the functions the reference
exports from glpapi01-09/12, glpcpx and glpscl only ``glp_bf_updated`` and the
sensitivity analysis (``glp_analyze_bound/coef``) are not mirrored.

``presolve: GLP_ON`` runs the reference's flow (lib/glpapi06.js:41-146,
lib/glpapi09.js:116-256) with the LP / MIP presolver of the native library
(``glpb_npp_*``, csrc/presolve.cpp: the transformations of lib/glpnpp01-05.js):
reduced problem, ``glp_scale_prob``, ``glp_adv_basis``, solve on the device,
recovery of the solution, stored back the way ``npp_unload_sol`` does.

What is NOT here on purpose ("out of scope"): MathProg, cut
generators, interior point.
"""
import math

import numpy as np

from . import native
from .native import *  # noqa: F401,F403  (GLP_* constants)

GLP_PROB_MAGIC = 0xD7D9D6C2


class GlpkError(Exception):
    """What xerror() throws in the reference (lib/glpapi.js:26)."""


def xerror(msg):
    raise GlpkError(msg)


_print_func = None


def glp_set_print_func(value):
    """lib/glpapi.js:33 -- the default hook prints nothing (:28-30)."""
    global _print_func
    _print_func = value


def glp_get_print_func():
    return _print_func


def xprintf(data):
    if _print_func is not None:
        _print_func(data)


def _num(x):
    """A number the way JavaScript's string concatenation shows it: Number::toString(10) of ECMAScript --
    the shortest round-trip digits (Python's repr yields the same ones) laid out by the position n of the
    decimal point: digits padded with zeros for k <= n <= 21, a point inside for 0 < n <= 21, '0.000ddd'
    for -6 < n <= 0, exponent form otherwise."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (math.inf, -math.inf):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "0"
    if x < 0:
        return "-" + _num(-x)
    r = repr(x)
    if "e" in r:
        mant, ex = r.split("e")
        digits, n = mant.replace(".", ""), int(ex) + 1
    else:
        ip, fp = r.split(".")
        if ip != "0":
            digits, n = ip + fp, len(ip)
        else:
            digits = fp.lstrip("0")
            n = -(len(fp) - len(digits))
    digits = digits.rstrip("0") or "0"
    k = len(digits)
    if k <= n <= 21:
        return digits + "0" * (n - k)
    if 0 < n <= 21:
        return digits[:n] + "." + digits[n:]
    if -6 < n <= 0:
        return "0." + "0" * (-n) + digits
    e = n - 1
    tail = "e" + ("+" if e > 0 else "-") + str(abs(e))
    return (digits if k == 1 else digits[0] + "." + digits[1:]) + tail


class SMCP:
    """lib/glpapi06.js:359-375, including the ``options[k] || default`` quirk:
    falsy values (0) given to the constructor fall back to the default."""

    _defaults = dict(msg_lev=GLP_MSG_ALL, meth=GLP_PRIMAL, pricing=GLP_PT_PSE, r_test=GLP_RT_HAR,
                     tol_bnd=1e-7, tol_dj=1e-7, tol_piv=1e-10, obj_ll=-DBL_MAX, obj_ul=+DBL_MAX,
                     it_lim=INT_MAX, tm_lim=INT_MAX, out_frq=500, out_dly=0, presolve=GLP_OFF)

    def __init__(self, options=None):
        options = options or {}
        for k, v in self._defaults.items():
            setattr(self, k, options.get(k) or v)


class IOCP:
    """lib/glpapi09.js:392-414 (fields of the B&B path only)."""

    _defaults = dict(msg_lev=GLP_MSG_ALL, br_tech=GLP_BR_DTH, bt_tech=GLP_BT_BLB, tol_int=1e-5,
                     tol_obj=1e-7, tm_lim=INT_MAX, out_frq=5000, out_dly=10000, cb_func=None,
                     cb_info=None, cb_size=0, pp_tech=GLP_PP_ALL, mip_gap=0.0, mir_cuts=GLP_OFF,
                     gmi_cuts=GLP_OFF, cov_cuts=GLP_OFF, clq_cuts=GLP_OFF, presolve=GLP_OFF,
                     binarize=GLP_OFF, fp_heur=GLP_OFF)

    def __init__(self, options=None):
        options = options or {}
        for k, v in self._defaults.items():
            setattr(self, k, options.get(k) or v)


class _Row:
    __slots__ = ("i", "name", "type", "lb", "ub", "rii", "stat", "bind", "prim", "dual", "mipx", "elems")

    def __init__(self, i):
        self.i, self.name = i, None
        self.type, self.lb, self.ub = GLP_FR, 0.0, 0.0
        self.rii = 1.0
        self.stat, self.bind = GLP_BS, 0
        self.prim = self.dual = self.mipx = 0.0
        self.elems = []  # [(j, val)] in list order


class _Col:
    __slots__ = ("j", "name", "kind", "type", "lb", "ub", "coef", "sjj", "stat", "bind", "prim",
                 "dual", "mipx", "elems")

    def __init__(self, j):
        self.j, self.name, self.kind = j, None, GLP_CV
        self.type, self.lb, self.ub, self.coef = GLP_FX, 0.0, 0.0, 0.0
        self.sjj = 1.0
        self.stat, self.bind = GLP_NS, 0
        self.prim = self.dual = self.mipx = 0.0
        self.elems = []  # [(i, val)] in list order


class glp_prob:
    def __init__(self):
        self.magic = GLP_PROB_MAGIC
        self.name = self.obj = None
        self.dir, self.c0 = GLP_MIN, 0.0
        self.m = self.n = self.nnz = 0
        self.row, self.col = [None], [None]
        self.valid = 0
        self.head = [0]
        self.pbs_stat = self.dbs_stat = GLP_UNDEF
        self.obj_val = 0.0
        self.it_cnt = self.some = 0
        self.mip_stat, self.mip_obj = GLP_UNDEF, 0.0
        self.bfcp = None
        self.r_tree = self.c_tree = None   # name indices (glp_create_index)
        self._dev = None       # native.Problem
        self._dirty = True     # matrix / costs changed since the handle was built


def glp_create_prob():
    return glp_prob()


def _check(P, who):
    if P is None or getattr(P, "magic", None) != GLP_PROB_MAGIC:
        xerror("%s: P = %r; invalid problem object" % (who, P))


def glp_set_prob_name(P, name):
    P.name = name


def glp_set_obj_name(P, name):
    P.obj = name


def glp_set_obj_dir(P, dir):
    if dir not in (GLP_MIN, GLP_MAX):
        xerror("glp_set_obj_dir: dir = %r; invalid direction flag" % (dir,))
    P.dir = dir
    P._dirty = True


def glp_add_rows(P, nrs):
    if nrs < 1:
        xerror("glp_add_rows: nrs = %d; invalid number of rows" % nrs)
    for i in range(P.m + 1, P.m + nrs + 1):
        P.row.append(_Row(i))
    P.head.extend([0] * nrs)
    P.m += nrs
    P.valid = 0
    P._dirty = True
    return P.m - nrs + 1


def glp_add_cols(P, ncs):
    if ncs < 1:
        xerror("glp_add_cols: ncs = %d; invalid number of columns" % ncs)
    for j in range(P.n + 1, P.n + ncs + 1):
        P.col.append(_Col(j))
    P.n += ncs
    P._dirty = True
    return P.n - ncs + 1


def glp_set_row_name(P, i, name):
    """lib/glpapi01.js:60-84 (the name index, when it exists, follows the change)"""
    if not (1 <= i <= P.m):
        xerror("glp_set_row_name: i = %d; row number out of range" % i)
    row = P.row[i]
    r_tree = getattr(P, "r_tree", None)
    if row.name is not None and r_tree is not None:
        r_tree.pop(row.name, None)
    row.name = name
    if name is not None and r_tree is not None:
        r_tree[name] = row


def glp_set_col_name(P, j, name):
    """lib/glpapi01.js:86-110"""
    if not (1 <= j <= P.n):
        xerror("glp_set_col_name: j = %d; column number out of range" % j)
    col = P.col[j]
    c_tree = getattr(P, "c_tree", None)
    if col.name is not None and c_tree is not None:
        c_tree.pop(col.name, None)
    col.name = name
    if name is not None and c_tree is not None:
        c_tree[name] = col


def _set_bnds(x, who, idx, type, lb, ub):
    # lib/glpapi01.js:217-281
    x.type = type
    if type == GLP_FR:
        x.lb = x.ub = 0.0
        if x.stat != GLP_BS:
            x.stat = GLP_NF
    elif type == GLP_LO:
        x.lb, x.ub = lb, 0.0
        if x.stat != GLP_BS:
            x.stat = GLP_NL
    elif type == GLP_UP:
        x.lb, x.ub = 0.0, ub
        if x.stat != GLP_BS:
            x.stat = GLP_NU
    elif type == GLP_DB:
        x.lb, x.ub = lb, ub
        if x.stat not in (GLP_BS, GLP_NL, GLP_NU):
            x.stat = GLP_NL if abs(lb) <= abs(ub) else GLP_NU
    elif type == GLP_FX:
        x.lb = x.ub = lb
        if x.stat != GLP_BS:
            x.stat = GLP_NS
    else:
        xerror("%s: %d; type = %r; invalid type" % (who, idx, type))


def glp_set_row_bnds(P, i, type, lb, ub):
    if not (1 <= i <= P.m):
        xerror("glp_set_row_bnds: i = %d; row number out of range" % i)
    _set_bnds(P.row[i], "glp_set_row_bnds: i =", i, type, lb, ub)


def glp_set_col_bnds(P, j, type, lb, ub):
    if not (1 <= j <= P.n):
        xerror("glp_set_col_bnds: j = %d; column number out of range" % j)
    _set_bnds(P.col[j], "glp_set_col_bnds: j =", j, type, lb, ub)


def glp_set_obj_coef(P, j, coef):
    if not (0 <= j <= P.n):
        xerror("glp_set_obj_coef: j = %d; column number out of range" % j)
    if j == 0:
        P.c0 = coef
    else:
        P.col[j].coef = coef
    P._dirty = True


def glp_set_mat_row(P, i, length, ind, val):
    """ind/val are 1-based sequences (slot 0 ignored); new elements are
    prepended to the row and column lists (lib/glpapi01.js:295-378)."""
    if not (1 <= i <= P.m):
        xerror("glp_set_mat_row: i = %d; row number out of range" % i)
    if not (0 <= length <= P.n):
        xerror("glp_set_mat_row: i = %d; len = %d; invalid row length" % (i, length))
    row = P.row[i]
    for (j, _) in row.elems:
        col = P.col[j]
        col.elems = [e for e in col.elems if e[0] != i]
        P.nnz -= 1
        if col.stat == GLP_BS:
            P.valid = 0
    row.elems = []
    for k in range(1, length + 1):
        j = ind[k]
        if not (1 <= j <= P.n):
            xerror("glp_set_mat_row: i = %d; ind[%d] = %d; column index out of range" % (i, k, j))
        col = P.col[j]
        if col.elems and col.elems[0][0] == i:
            xerror("glp_set_mat_row: i = %d; ind[%d] = %d; duplicate column indices not allowed" % (i, k, j))
        if val[k] == 0.0:
            continue
        row.elems.insert(0, (j, float(val[k])))
        col.elems.insert(0, (i, float(val[k])))
        P.nnz += 1
        if col.stat == GLP_BS:
            P.valid = 0
    P._dirty = True


def glp_set_mat_col(P, j, length, ind, val):
    """lib/glpapi01.js:380-462"""
    if not (1 <= j <= P.n):
        xerror("glp_set_mat_col: j = %d; column number out of range" % j)
    if not (0 <= length <= P.m):
        xerror("glp_set_mat_col: j = %d; len = %d; invalid column length" % (j, length))
    col = P.col[j]
    for (i, _) in col.elems:
        row = P.row[i]
        row.elems = [e for e in row.elems if e[0] != j]
        P.nnz -= 1
    col.elems = []
    for k in range(1, length + 1):
        i = ind[k]
        if not (1 <= i <= P.m):
            xerror("glp_set_mat_col: j = %d; ind[%d] = %d; row index out of range" % (j, k, i))
        row = P.row[i]
        if row.elems and row.elems[0][0] == j:
            xerror("glp_set_mat_col: j = %d; ind[%d] = %d; duplicate row indices not allowed" % (j, k, i))
        if val[k] == 0.0:
            continue
        col.elems.insert(0, (i, float(val[k])))
        row.elems.insert(0, (j, float(val[k])))
        P.nnz += 1
    if col.stat == GLP_BS:
        P.valid = 0
    P._dirty = True


def glp_load_matrix(P, ne, ia, ja, ar):
    """lib/glpapi01.js:464-560: replace the whole matrix by ne triplets (1-based).  List order as the
    reference leaves it: every ROW list holds its elements in reverse input order (prepended as they
    come); the COLUMN lists are built afterwards by walking rows 1..m and prepending, so every column
    list runs over DESCENDING row indices; zero elements are removed last."""
    for r in P.row[1:]:
        r.elems = []
    for c in P.col[1:]:
        c.elems = []
    P.nnz = 0
    if ne < 0:
        xerror("glp_load_matrix: ne = %d; invalid number of constraint coefficients" % ne)
    for k in range(1, ne + 1):
        i, j = ia[k], ja[k]
        if not (1 <= i <= P.m):
            xerror("glp_load_matrix: ia[%d] = %d; row index out of range" % (k, i))
        if not (1 <= j <= P.n):
            xerror("glp_load_matrix: ja[%d] = %d; column index out of range" % (k, j))
        P.row[i].elems.insert(0, (j, float(ar[k])))
    for i in range(1, P.m + 1):
        for (j, v) in P.row[i].elems:
            col = P.col[j]
            if col.elems and col.elems[0][0] == i:
                k = next(k for k in range(1, ne + 1) if ia[k] == i and ja[k] == j)
                xerror("glp_load_mat: ia[%d] = %d; ja[%d] = %d; duplicate indices not allowed" % (k, i, k, j))
            col.elems.insert(0, (i, v))
    for r in P.row[1:]:
        r.elems = [e for e in r.elems if e[1] != 0.0]
    for c in P.col[1:]:
        c.elems = [e for e in c.elems if e[1] != 0.0]
    P.nnz = sum(len(c.elems) for c in P.col[1:])
    P.valid = 0
    P._dirty = True


def glp_sort_matrix(P):
    """lib/glpapi01.js:620-649"""
    _check(P, "glp_sort_matrix")
    for r in P.row[1:]:
        r.elems.sort(key=lambda e: e[0])
    for c in P.col[1:]:
        c.elems.sort(key=lambda e: e[0])
    P._dirty = True


def glp_set_col_kind(P, j, kind):
    """lib/glpapi09.js:1-38"""
    if not (1 <= j <= P.n):
        xerror("glp_set_col_kind: j = %d; column number out of range" % j)
    col = P.col[j]
    if kind == GLP_CV:
        col.kind = GLP_CV
    elif kind == GLP_IV:
        col.kind = GLP_IV
    elif kind == GLP_BV:
        col.kind = GLP_IV
        if not (col.type == GLP_DB and col.lb == 0.0 and col.ub == 1.0):
            glp_set_col_bnds(P, j, GLP_DB, 0.0, 1.0)
    else:
        xerror("glp_set_col_kind: j = %d; kind = %r; invalid column kind" % (j, kind))
    P._dirty = True


def glp_get_col_kind(P, j):
    col = P.col[j]
    if col.kind == GLP_IV and col.type == GLP_DB and col.lb == 0.0 and col.ub == 1.0:
        return GLP_BV
    return col.kind


def _norm_stat(x, stat):
    if stat != GLP_BS:
        stat = {GLP_FR: GLP_NF, GLP_LO: GLP_NL, GLP_UP: GLP_NU,
                GLP_DB: (GLP_NU if stat == GLP_NU else GLP_NL), GLP_FX: GLP_NS}[x.type]
    return stat


def glp_set_row_stat(P, i, stat):
    """lib/glpapi05.js:1-23"""
    if not (1 <= i <= P.m):
        xerror("glp_set_row_stat: i = %d; row number out of range" % i)
    if stat not in (GLP_BS, GLP_NL, GLP_NU, GLP_NF, GLP_NS):
        xerror("glp_set_row_stat: i = %d; stat = %r; invalid status" % (i, stat))
    row = P.row[i]
    stat = _norm_stat(row, stat)
    if (row.stat == GLP_BS) != (stat == GLP_BS):
        P.valid = 0
    row.stat = stat


def glp_set_col_stat(P, j, stat):
    """lib/glpapi05.js:25-47"""
    if not (1 <= j <= P.n):
        xerror("glp_set_col_stat: j = %d; column number out of range" % j)
    if stat not in (GLP_BS, GLP_NL, GLP_NU, GLP_NF, GLP_NS):
        xerror("glp_set_col_stat: j = %d; stat = %r; invalid status" % (j, stat))
    col = P.col[j]
    stat = _norm_stat(col, stat)
    if (col.stat == GLP_BS) != (stat == GLP_BS):
        P.valid = 0
    col.stat = stat


def glp_std_basis(P):
    """lib/glpapi05.js:49-63"""
    for i in range(1, P.m + 1):
        glp_set_row_stat(P, i, GLP_BS)
    for j in range(1, P.n + 1):
        col = P.col[j]
        if col.type == GLP_DB and abs(col.lb) > abs(col.ub):
            glp_set_col_stat(P, j, GLP_NU)
        else:
            glp_set_col_stat(P, j, GLP_NL)


# ---- problem-object housekeeping (lib/glpapi01.js:559-618,651-860, lib/glpapi03.js, lib/glpapi.js:76) ----
def glp_version():
    return "4.49"


def glp_check_dup(m, n, ne, ia, ja):
    """lib/glpapi01.js:559-618: 0 = no duplicates, -k = ia[k]/ja[k] out of range,
    +k = element k duplicates an earlier one (1-based triplets)"""
    if m < 0:
        xerror("glp_check_dup: m = %d; invalid parameter" % m)
    if n < 0:
        xerror("glp_check_dup: n = %d; invalid parameter" % n)
    if ne < 0:
        xerror("glp_check_dup: ne = %d; invalid parameter" % ne)
    if ne > 0 and ia is None:
        xerror("glp_check_dup: ia = %r; invalid parameter" % (ia,))
    if ne > 0 and ja is None:
        xerror("glp_check_dup: ja = %r; invalid parameter" % (ja,))
    for k in range(1, ne + 1):
        if not (1 <= ia[k] <= m and 1 <= ja[k] <= n):
            return -k
    if m == 0 or n == 0:
        return 0
    # the reference walks the rows in order, each row's elements from the last to the first,
    # and reports the SECOND occurrence (in input order) of the first pair it meets twice
    by_row = {}
    for k in range(1, ne + 1):
        by_row.setdefault(ia[k], []).append(k)
    for i in sorted(by_row):
        seen = set()
        for k in reversed(by_row[i]):
            j = ja[k]
            if j in seen:
                hits = [t for t in by_row[i] if ja[t] == j]
                return hits[1]
            seen.add(j)
    return 0


def glp_create_index(P):
    """lib/glpapi03.js:1-25"""
    if getattr(P, "r_tree", None) is None:
        P.r_tree = {r.name: r for r in P.row[1:] if r.name is not None}
    if getattr(P, "c_tree", None) is None:
        P.c_tree = {c.name: c for c in P.col[1:] if c.name is not None}


def glp_find_row(P, name):
    if getattr(P, "r_tree", None) is None:
        xerror("glp_find_row: row name index does not exist")
    row = P.r_tree.get(name)
    return row.i if row is not None else 0


def glp_find_col(P, name):
    if getattr(P, "c_tree", None) is None:
        xerror("glp_find_col: column name index does not exist")
    col = P.c_tree.get(name)
    return col.j if col is not None else 0


def glp_delete_index(P):
    """lib/glpapi03.js:45-48 -- as written there only the ROW index goes
    (``lp.r_tree = null`` twice); the column index survives."""
    P.r_tree = None


def glp_del_rows(P, nrs, num):
    """lib/glpapi01.js:651-703 (num[1..nrs]); remaining rows are renumbered"""
    if not (1 <= nrs <= P.m):
        xerror("glp_del_rows: nrs = %d; invalid number of rows" % nrs)
    for k in range(1, nrs + 1):
        i = num[k]
        if not (1 <= i <= P.m):
            xerror("glp_del_rows: num[%d] = %d; row number out of range" % (k, i))
        row = P.row[i]
        if row.i == 0:
            xerror("glp_del_rows: num[%d] = %d; duplicate row numbers not allowed" % (k, i))
        glp_set_row_name(P, i, None)
        glp_set_mat_row(P, i, 0, None, None)
        row.i = 0
    remap, kept = {}, [None]
    for i in range(1, P.m + 1):
        row = P.row[i]
        if row.i != 0:
            kept.append(row)
            remap[i] = row.i = len(kept) - 1
    P.row, P.m = kept, len(kept) - 1
    for col in P.col[1:]:
        col.elems = [(remap[i], v) for (i, v) in col.elems]
    P.head = P.head[:P.m + 1]
    P.valid = 0
    P._dirty = True


def glp_del_cols(P, ncs, num):
    """lib/glpapi01.js:705-761"""
    if not (1 <= ncs <= P.n):
        xerror("glp_del_cols: ncs = %d; invalid number of columns" % ncs)
    for k in range(1, ncs + 1):
        j = num[k]
        if not (1 <= j <= P.n):
            xerror("glp_del_cols: num[%d] = %d; column number out of range" % (k, j))
        col = P.col[j]
        if col.j == 0:
            xerror("glp_del_cols: num[%d] = %d; duplicate column numbers not allowed" % (k, j))
        glp_set_col_name(P, j, None)
        glp_set_mat_col(P, j, 0, None, None)
        col.j = 0
        if col.stat == GLP_BS:
            P.valid = 0
    remap, kept = {}, [None]
    for j in range(1, P.n + 1):
        col = P.col[j]
        if col.j != 0:
            kept.append(col)
            remap[j] = col.j = len(kept) - 1
    P.col, P.n = kept, len(kept) - 1
    for row in P.row[1:]:
        row.elems = [(remap[j], v) for (j, v) in row.elems]
    if P.valid:
        for j in range(1, P.n + 1):
            k = P.col[j].bind
            if k != 0:
                P.head[k] = P.m + j
    P._dirty = True


def glp_erase_prob(P):
    """lib/glpapi01.js:836-842: back to the state glp_create_prob leaves"""
    _check(P, "glp_erase_prob")
    _drop_device(P)
    fresh = glp_prob()
    P.__dict__.clear()
    P.__dict__.update(fresh.__dict__)


def glp_copy_prob(dest, prob, names):
    """lib/glpapi01.js:763-834.  Columns are copied through glp_set_mat_col, so the
    copy's lists are in the order that call leaves (each list reversed)."""
    _check(dest, "glp_copy_prob")
    _check(prob, "glp_copy_prob")
    if dest is prob:
        xerror("glp_copy_prob: copying problem object to itself not allowed")
    if names not in (GLP_ON, GLP_OFF):
        xerror("glp_copy_prob: names = %r; invalid parameter" % (names,))
    glp_erase_prob(dest)
    if names and prob.name is not None:
        glp_set_prob_name(dest, prob.name)
    if names and prob.obj is not None:
        glp_set_obj_name(dest, prob.obj)
    dest.dir, dest.c0 = prob.dir, prob.c0
    if prob.m > 0:
        glp_add_rows(dest, prob.m)
    if prob.n > 0:
        glp_add_cols(dest, prob.n)
    dest.bfcp = None if prob.bfcp is None else dict(prob.bfcp)
    dest.pbs_stat, dest.dbs_stat, dest.obj_val, dest.some = prob.pbs_stat, prob.dbs_stat, prob.obj_val, prob.some
    dest.mip_stat, dest.mip_obj = prob.mip_stat, prob.mip_obj
    for i in range(1, prob.m + 1):
        to, src = dest.row[i], prob.row[i]
        if names and src.name is not None:
            glp_set_row_name(dest, i, src.name)
        to.type, to.lb, to.ub, to.rii = src.type, src.lb, src.ub, src.rii
        to.stat, to.prim, to.dual, to.mipx = src.stat, src.prim, src.dual, src.mipx
    for j in range(1, prob.n + 1):
        to, src = dest.col[j], prob.col[j]
        if names and src.name is not None:
            glp_set_col_name(dest, j, src.name)
        to.kind, to.type, to.lb, to.ub, to.coef = src.kind, src.type, src.lb, src.ub, src.coef
        ln = len(src.elems)
        glp_set_mat_col(dest, j, ln, [0] + [i for (i, _) in src.elems], [0.0] + [v for (_, v) in src.elems])
        to.sjj = src.sjj
        to.stat, to.prim, to.dual, to.mipx = src.stat, src.prim, src.dual, src.mipx


# ---- scale factors, scaling, crash basis (lib/glpapi04.js, glpscl.js, glpini01.js) ----
def glp_set_rii(P, i, rii):
    """lib/glpapi04.js:1-16"""
    if not (1 <= i <= P.m):
        xerror("glp_set_rii: i = %d; row number out of range" % i)
    if rii <= 0.0:
        xerror("glp_set_rii: i = %d; rii = %s; invalid scale factor" % (i, _num(rii)))
    row = P.row[i]
    if row.rii != rii:
        if P.valid and any(P.col[j].stat == GLP_BS for (j, _) in row.elems):
            P.valid = 0
        P._dirty = True   # the handle holds rii*a*sjj
    row.rii = float(rii)


def glp_set_sjj(P, j, sjj):
    """lib/glpapi04.js:18-28"""
    if not (1 <= j <= P.n):
        xerror("glp_set_sjj: j = %d; column number out of range" % j)
    if sjj <= 0.0:
        xerror("glp_set_sjj: j = %d; sjj = %s; invalid scale factor" % (j, _num(sjj)))
    col = P.col[j]
    if col.sjj != sjj:
        if P.valid and col.stat == GLP_BS:
            P.valid = 0
        P._dirty = True
    col.sjj = float(sjj)


def glp_get_rii(P, i):
    if not (1 <= i <= P.m):
        xerror("glp_get_rii: i = %d; row number out of range" % i)
    return P.row[i].rii


def glp_get_sjj(P, j):
    if not (1 <= j <= P.n):
        xerror("glp_get_sjj: j = %d; column number out of range" % j)
    return P.col[j].sjj


def glp_unscale_prob(P):
    """lib/glpapi04.js:44-50"""
    for i in range(1, P.m + 1):
        glp_set_rii(P, i, 1.0)
    for j in range(1, P.n + 1):
        glp_set_sjj(P, j, 1.0)


def _csc(P):
    """Columns of A in list order: (ptr[n+1], 0-based row index, value)."""
    ptr = np.zeros(P.n + 1, np.int32)
    ind, val = [], []
    for j in range(1, P.n + 1):
        for (i, v) in P.col[j].elems:
            ind.append(i - 1)
            val.append(v)
        ptr[j] = len(ind)
    return ptr, np.array(ind, np.int32), np.array(val, np.float64)


def _csr(P):
    """Rows of A in list order: (ptr[m+1], 0-based column index, value)."""
    ptr = np.zeros(P.m + 1, np.int32)
    ind, val = [], []
    for i in range(1, P.m + 1):
        for (j, v) in P.row[i].elems:
            ind.append(j - 1)
            val.append(v)
        ptr[i] = len(ind)
    return ptr, np.array(ind, np.int32), np.array(val, np.float64)


def glp_scale_prob(P, flags=None):
    """lib/glpscl.js:216-225.  The min/max sweeps run in the native library
    (``glpb_scale_prob``); the messages are the reference's.  ``flags`` left out
    behaves like JavaScript's ``undefined``: every ``flags & X`` is 0, so the
    call only cancels the current scaling (that is what test/test.js:76 does)."""
    _check(P, "glp_scale_prob")
    flags = 0 if flags is None else int(flags)
    if flags & ~(GLP_SF_GM | GLP_SF_EQ | GLP_SF_2N | GLP_SF_SKIP | GLP_SF_AUTO):
        xerror("glp_scale_prob: flags = %d; invalid scaling options" % flags)
    ptr, ind, val = _csc(P)
    rii, sjj, rep = native.scale_prob(P.m, P.n, ptr, ind, val, flags)
    xprintf("Scaling...")
    for tag in ("A", "GM", "EQ", "2N"):
        if tag in rep:
            lo, hi, ratio = rep[tag]
            xprintf("%2s: min|aij| = %s  max|aij| = %s  ratio = %s" % (tag, _num(lo), _num(hi), _num(ratio)))
            if tag == "A" and lo >= 0.10 and hi <= 10.0:
                xprintf("Problem data seem to be well scaled")
    for i in range(1, P.m + 1):
        glp_set_rii(P, i, float(rii[i - 1]))
    for j in range(1, P.n + 1):
        glp_set_sjj(P, j, float(sjj[j - 1]))


def glp_adv_basis(P, flags=0):
    """lib/glpini01.js:355-363; the triangularisation runs in the native library
    (``glpb_adv_basis``)."""
    _check(P, "glp_adv_basis")
    if flags != 0:
        xerror("glp_adv_basis: flags = %r; invalid flags" % (flags,))
    if P.m == 0 or P.n == 0:
        glp_std_basis(P)
        return
    xprintf("Constructing initial basis...")
    m, n = P.m, P.n
    cptr, cind, _ = _csc(P)
    rptr, rind, _ = _csr(P)
    type_ = np.array([P.row[i].type for i in range(1, m + 1)] + [P.col[j].type for j in range(1, n + 1)], np.int32)
    lb = np.array([P.row[i].lb for i in range(1, m + 1)] + [P.col[j].lb for j in range(1, n + 1)])
    ub = np.array([P.row[i].ub for i in range(1, m + 1)] + [P.col[j].ub for j in range(1, n + 1)])
    stat, size = native.adv_basis(m, n, cptr, cind, rptr, rind, type_, lb, ub)
    P.tri_size = size
    xprintf("Size of triangular part = %d" % size)     # lib/glpini01.js:287-288 (LPX message level: default 3)
    for i in range(1, m + 1):
        glp_set_row_stat(P, i, int(stat[i - 1]))
    for j in range(1, n + 1):
        glp_set_col_stat(P, j, int(stat[m + j - 1]))


# ---- getters (lib/glpapi02.js, lib/glpapi06.js:398-482, lib/glpapi09.js:441-459) ----
def glp_get_num_rows(P): return P.m
def glp_get_num_cols(P): return P.n
def glp_get_num_nz(P): return P.nnz
def glp_get_obj_dir(P): return P.dir
def glp_get_prob_name(P): return P.name
def glp_get_row_name(P, i): return P.row[i].name
def glp_get_col_name(P, j): return P.col[j].name
def glp_get_num_int(P): return sum(1 for c in P.col[1:] if c.kind == GLP_IV)
def glp_get_num_bin(P): return sum(1 for c in P.col[1:] if c.kind == GLP_IV and c.type == GLP_DB and c.lb == 0.0 and c.ub == 1.0)
def glp_get_prim_stat(P): return P.pbs_stat
def glp_get_dual_stat(P): return P.dbs_stat
def glp_get_obj_val(P): return P.obj_val
def glp_mip_status(P): return P.mip_stat
def glp_mip_obj_val(P): return P.mip_obj


def glp_get_status(P):
    """lib/glpapi06.js:398-427"""
    status = P.pbs_stat
    if status == GLP_FEAS:
        if P.dbs_stat == GLP_FEAS:
            status = GLP_OPT
        elif P.dbs_stat == GLP_NOFEAS:
            status = GLP_UNBND
    return status


def _row(P, i, who):
    if not (1 <= i <= P.m):
        xerror("%s: i = %d; row number out of range" % (who, i))
    return P.row[i]


def _col(P, j, who):
    if not (1 <= j <= P.n):
        xerror("%s: j = %d; column number out of range" % (who, j))
    return P.col[j]


def glp_get_row_stat(P, i): return _row(P, i, "glp_get_row_stat").stat
def glp_get_row_prim(P, i): return _row(P, i, "glp_get_row_prim").prim
def glp_get_row_dual(P, i): return _row(P, i, "glp_get_row_dual").dual
def glp_get_col_stat(P, j): return _col(P, j, "glp_get_col_stat").stat
def glp_get_col_prim(P, j): return _col(P, j, "glp_get_col_prim").prim
def glp_get_col_dual(P, j): return _col(P, j, "glp_get_col_dual").dual
def glp_mip_row_val(P, i): return _row(P, i, "glp_mip_row_val").mipx
def glp_mip_col_val(P, j): return _col(P, j, "glp_mip_col_val").mipx


def glp_get_obj_name(P): return P.obj
def glp_get_row_type(P, i): return _row(P, i, "glp_get_row_type").type
def glp_get_col_type(P, j): return _col(P, j, "glp_get_col_type").type


def glp_get_row_lb(P, i):
    """lib/glpapi02.js:36-52: -DBL_MAX when the row has no lower bound"""
    r = _row(P, i, "glp_get_row_lb")
    return -DBL_MAX if r.type in (GLP_FR, GLP_UP) else r.lb


def glp_get_row_ub(P, i):
    r = _row(P, i, "glp_get_row_ub")
    return +DBL_MAX if r.type in (GLP_FR, GLP_LO) else r.ub


def glp_get_col_lb(P, j):
    c = _col(P, j, "glp_get_col_lb")
    return -DBL_MAX if c.type in (GLP_FR, GLP_UP) else c.lb


def glp_get_col_ub(P, j):
    c = _col(P, j, "glp_get_col_ub")
    return +DBL_MAX if c.type in (GLP_FR, GLP_LO) else c.ub


def glp_get_obj_coef(P, j):
    if not (0 <= j <= P.n):
        xerror("glp_get_obj_coef: j = %d; column number out of range" % j)
    return P.c0 if j == 0 else P.col[j].coef


def glp_get_mat_row(P, i, ind=None, val=None):
    """lib/glpapi02.js:127-140: fills ind[1..len] / val[1..len] in list order"""
    r = _row(P, i, "glp_get_mat_row")
    for t, (j, v) in enumerate(r.elems, 1):
        if ind is not None:
            ind[t] = j
        if val is not None:
            val[t] = v
    return len(r.elems)


def glp_get_mat_col(P, j, ind=None, val=None):
    c = _col(P, j, "glp_get_mat_col")
    for t, (i, v) in enumerate(c.elems, 1):
        if ind is not None:
            ind[t] = i
        if val is not None:
            val[t] = v
    return len(c.elems)


def glp_get_unbnd_ray(P):
    k = P.some
    return 0 if k > P.m + P.n else k


# ---- CPLEX LP format (lib/glpcpx.js:10-753), own recursive-descent reader ----
_NAME_EXTRA = set("!\"#$%&()/,.;?@_`'{}|~")
_KEYWORDS = {
    "minimize": "MIN", "minimum": "MIN", "min": "MIN", "maximize": "MAX", "maximum": "MAX", "max": "MAX",
    "st": "ST", "s.t.": "ST", "st.": "ST", "bounds": "BOUNDS", "bound": "BOUNDS",
    "general": "GEN", "generals": "GEN", "gen": "GEN", "integer": "INT", "integers": "INT", "int": "INT",
    "binary": "BIN", "binaries": "BIN", "bin": "BIN", "end": "END",
}


def _tokenize(text):
    """[(kind, image, value, colon_follows, line, end_of_line)].  As the reference's scanner
    (lib/glpcpx.js:79-258): keywords only for a letter in column 0; a name is followed by a
    colon if ':' is the next non-blank character of its line; end_of_line = nothing but blanks
    follows (a comment counts as something)."""
    toks = []
    lineno = 0
    for raw in text.split("\n"):
        lineno += 1
        full = "".join(" " if c in "\t\r\v\f" else c for c in raw)
        n = full.find("\\")
        n = len(full) if n < 0 else n
        line = full

        def rest(j):
            while j < len(full) and full[j] == " ":
                j += 1
            return j
        i = 0
        while i < n:
            ch = line[i]
            if ch == " ":
                i += 1
                continue
            if ch.isalpha() or (ch in _NAME_EXTRA and ch != "."):
                j = i
                while j < n and (line[j].isalnum() or line[j] in _NAME_EXTRA):
                    j += 1
                image = line[i:j]
                kind = "NAME"
                if i == 0 and ch.isalpha():
                    low = image.lower()
                    if low in ("subject", "such"):
                        rst = line[j:n].lstrip(" ")
                        want = "to" if low == "subject" else "that"
                        if rst.lower().startswith(want) and (len(rst) == len(want) or not rst[len(want)].isalnum()):
                            kind = "ST"
                            j = n - len(rst) + len(want)
                    elif low in _KEYWORDS:
                        kind = _KEYWORDS[low]
                r = rest(j)
                toks.append((kind, image, 0.0, r < len(full) and full[r] == ":", lineno, r >= len(full)))
                i = j
            elif ch.isdigit() or ch == ".":
                j = i
                while j < n and line[j].isdigit():
                    j += 1
                if j < n and line[j] == ".":
                    j += 1
                    if j - i == 1 and not (j < n and line[j].isdigit()):
                        xerror("glp_read_lp: line %d: invalid use of decimal point" % lineno)
                    while j < n and line[j].isdigit():
                        j += 1
                if j < n and line[j] in "eE":
                    j += 1
                    if j < n and line[j] in "+-":
                        j += 1
                    if not (j < n and line[j].isdigit()):
                        xerror("glp_read_lp: line %d: numeric constant `%s' incomplete" % (lineno, line[i:j]))
                    while j < n and line[j].isdigit():
                        j += 1
                toks.append(("NUM", line[i:j], float(line[i:j]), False, lineno, rest(j) >= len(full)))
                i = j
            elif ch in "+-:":
                toks.append(({"+": "PLUS", "-": "MINUS", ":": "COLON"}[ch], ch, 0.0, False, lineno, rest(i + 1) >= len(full)))
                i += 1
            elif ch in "<>=":
                j = i + 1
                if j < n and line[j] in "<>=":
                    j += 1
                op = line[i:j]
                kind = "LE" if "<" in op else ("GE" if ">" in op else "EQ")
                toks.append((kind, op, 0.0, False, lineno, rest(j) >= len(full)))
                i = j
            else:
                xerror("glp_read_lp: line %d: character `%s' not recognized" % (lineno, ch))
    toks.append(("EOF", "", 0.0, False, lineno, True))
    return toks


def glp_read_lp_from_string(P, parm, text):
    """lib/glpcpx.js:1000-1010; returns 0 on success, 1 on a syntax error (the
    problem object is then erased, :745-748).  The text is parsed by the native
    reader (``glpb_read_lp``); ``_read_lp`` below is the same grammar in Python,
    kept as its cross-check (tests/test_hostprep.py)."""
    _check(P, "glp_read_lp")
    fresh = glp_prob()
    P.__dict__.update(fresh.__dict__)
    xprintf("Reading problem data")
    try:
        d, names = native.read_lp(text)
    except ValueError as e:
        xprintf(str(e))
        return 1
    m, n = d["m"], d["n"]
    P.dir, P.obj = d["dir"], names["obj"]
    P.m, P.n, P.nnz = m, n, d["nnz"]
    for i in range(1, m + 1):
        r = _Row(i)
        r.name, r.type, r.lb, r.ub = names["rows"][i - 1], int(d["type"][i - 1]), float(d["lb"][i - 1]), float(d["ub"][i - 1])
        r.stat = GLP_BS
        P.row.append(r)
    ptr, ind, val = d["A_ptr"], d["A_ind"], d["A_val"]
    for j in range(1, n + 1):
        c = _Col(j)
        k = m + j - 1
        c.name = names["cols"][j - 1]
        _set_bnds(c, "glp_set_col_bnds: j =", j, int(d["type"][k]), float(d["lb"][k]), float(d["ub"][k]))
        c.coef, c.kind = float(d["coef"][j - 1]), int(d["kind"][j - 1])
        for t in range(ptr[j - 1], ptr[j]):
            i, v = int(ind[t]) + 1, float(val[t])
            c.elems.append((i, v))            # ascending rows: the state after glp_sort_matrix
            P.row[i].elems.append((j, v))     # columns arrive in ascending order too
        P.col.append(c)
    xprintf("%s" % _size_line(P))
    return 0


def glp_read_lp(P, parm, callback):
    """lib/glpcpx.js:10: ``callback()`` returns the next character or a chunk of
    text and a falsy value at end of input (a plain string is accepted too)."""
    if isinstance(callback, str):
        return glp_read_lp_from_string(P, parm, callback)
    parts = []
    while True:
        s = callback()
        if not s or s == -1:      # XEOF = -1 (lib/glpapi.js:21), what test/test.js returns
            break
        parts.append(s)
    return glp_read_lp_from_string(P, parm, "".join(parts))


def _read_lp(P, text):
    fresh = glp_prob()
    P.__dict__.update(fresh.__dict__)
    toks = _tokenize(text)
    pos = [0]
    cols, rows, lbs, ubs = {}, {}, {}, {}

    def tok():
        return toks[pos[0]]

    def adv():
        pos[0] += 1

    def find_col(name):
        j = cols.get(name)
        if j is None:
            j = glp_add_cols(P, 1)
            glp_set_col_name(P, j, name)
            cols[name] = j
        return j

    def linear_form():
        ind, val, used = [0], [0.0], set()
        while True:
            s, coef = 1.0, 1.0
            if tok()[0] in ("PLUS", "MINUS"):
                s = 1.0 if tok()[0] == "PLUS" else -1.0
                adv()
            if tok()[0] == "NUM":
                coef = tok()[2]
                adv()
            if tok()[0] != "NAME":
                xerror("glp_read_lp: missing variable name")
            j = find_col(tok()[1])
            if j in used:
                xerror("glp_read_lp: multiple use of variable `%s' not allowed" % tok()[1])
            used.add(j)
            ind.append(j)
            val.append(s * coef)
            adv()
            if tok()[0] not in ("PLUS", "MINUS"):
                break
        keep = [(j, v) for j, v in zip(ind[1:], val[1:]) if v != 0.0]
        return len(keep), [0] + [j for j, _ in keep], [0.0] + [v for _, v in keep]

    def signed_number(what):
        s = 1.0
        if tok()[0] in ("PLUS", "MINUS"):
            s = 1.0 if tok()[0] == "PLUS" else -1.0
            adv()
        if tok()[0] != "NUM":
            xerror("glp_read_lp: missing " + what)
        v = s * tok()[2]
        adv()
        return v

    def bound_value(lower):
        s, signed = 1.0, False
        if tok()[0] in ("PLUS", "MINUS"):
            s, signed = (1.0 if tok()[0] == "PLUS" else -1.0), True
            adv()
        if tok()[0] == "NUM":
            v = s * tok()[2]
            adv()
            return v
        if signed and tok()[0] == "NAME" and tok()[1].lower() in ("infinity", "inf"):
            if lower and s > 0:
                xerror("glp_read_lp: invalid use of `+inf' as lower bound")
            if not lower and s < 0:
                xerror("glp_read_lp: invalid use of `-inf' as upper bound")
            adv()
            return -DBL_MAX if lower else +DBL_MAX
        xerror("glp_read_lp: missing %s bound" % ("lower" if lower else "upper"))

    if tok()[0] not in ("MIN", "MAX"):
        xerror("glp_read_lp: `minimize' or `maximize' keyword missing")
    glp_set_obj_dir(P, GLP_MIN if tok()[0] == "MIN" else GLP_MAX)
    adv()
    if tok()[0] == "NAME" and tok()[3]:
        glp_set_obj_name(P, tok()[1])
        adv()
        adv()
    else:
        glp_set_obj_name(P, "obj")
    ln, ind, val = linear_form()
    for k in range(1, ln + 1):
        glp_set_obj_coef(P, ind[k], val[k])
    if tok()[0] != "ST":
        xerror("glp_read_lp: constraints section missing")
    adv()
    while True:
        i = glp_add_rows(P, 1)
        if tok()[0] == "NAME" and tok()[3]:
            if tok()[1] in rows:
                xerror("glp_read_lp: constraint `%s' multiply defined" % tok()[1])
            rows[tok()[1]] = i
            glp_set_row_name(P, i, tok()[1])
            adv()
            adv()
        else:
            glp_set_row_name(P, i, "r.%d" % tok()[4])        # "r." + csa.count: the line number (lib/glpcpx.js:395)
        ln, ind, val = linear_form()
        glp_set_mat_row(P, i, ln, ind, val)
        sense = tok()[0]
        if sense not in ("LE", "GE", "EQ"):
            xerror("glp_read_lp: missing constraint sense")
        adv()
        sgn = 1.0
        if tok()[0] in ("PLUS", "MINUS"):
            sgn = 1.0 if tok()[0] == "PLUS" else -1.0
            adv()
        if tok()[0] != "NUM":
            xerror("glp_read_lp: missing right-hand side")
        rhs = sgn * tok()[2]
        if not tok()[5]:                                   # lib/glpcpx.js:431-433
            xerror("glp_read_lp: line %d: invalid symbol(s) beyond right-hand side" % tok()[4])
        adv()
        glp_set_row_bnds(P, i, {"LE": GLP_UP, "GE": GLP_LO, "EQ": GLP_FX}[sense], rhs, rhs)
        if tok()[0] not in ("PLUS", "MINUS", "NUM", "NAME"):
            break
    if tok()[0] == "BOUNDS":
        adv()
        while tok()[0] in ("PLUS", "MINUS", "NUM", "NAME"):
            lb_flag = tok()[0] != "NAME"
            if lb_flag:
                lbv = bound_value(True)
                if tok()[0] != "LE":
                    xerror("glp_read_lp: missing `<', `<=', or `=<' after lower bound")
                adv()
            if tok()[0] != "NAME":
                xerror("glp_read_lp: missing variable name")
            j = find_col(tok()[1])
            if lb_flag:
                lbs[j] = lbv
            adv()
            if tok()[0] == "LE":
                adv()
                ubs[j] = bound_value(False)
            elif tok()[0] == "GE":
                if lb_flag:
                    xerror("glp_read_lp: invalid bound definition")
                adv()
                lbs[j] = bound_value(True)
            elif tok()[0] == "EQ":
                if lb_flag:
                    xerror("glp_read_lp: invalid bound definition")
                adv()
                lbs[j] = ubs[j] = signed_number("fixed value")
            elif tok()[0] == "NAME" and tok()[1].lower() == "free":
                if lb_flag:
                    xerror("glp_read_lp: invalid bound definition")
                lbs[j], ubs[j] = -DBL_MAX, +DBL_MAX
                adv()
            elif not lb_flag:
                xerror("glp_read_lp: invalid bound definition")
    while tok()[0] in ("GEN", "INT", "BIN"):
        binary = tok()[0] == "BIN"
        adv()
        while tok()[0] == "NAME":
            j = find_col(tok()[1])
            glp_set_col_kind(P, j, GLP_IV)
            if binary:
                lbs[j], ubs[j] = 0.0, 1.0
            adv()
    if tok()[0] == "END":
        adv()
    elif tok()[0] != "EOF":
        xerror("glp_read_lp: symbol %s in wrong position" % tok()[1])
    if tok()[0] != "EOF":
        xerror("glp_read_lp: extra symbol(s) detected beyond `end'")
    for j in range(1, P.n + 1):  # lib/glpcpx.js:698-717
        lb, ub = lbs.get(j, 0.0), ubs.get(j, +DBL_MAX)
        if lb == -DBL_MAX and ub == +DBL_MAX:
            t = GLP_FR
        elif ub == +DBL_MAX:
            t = GLP_LO
        elif lb == -DBL_MAX:
            t = GLP_UP
        elif lb != ub:
            t = GLP_DB
        else:
            t = GLP_FX
        glp_set_col_bnds(P, j, t, lb, ub)
    glp_sort_matrix(P)


# ---- marshalling to the device handle ----
def glp_write_lp(P, parm, callback):
    """lib/glpcpx.js:755-998: CPLEX LP text, one ``callback(line)`` per line.  The text is produced by the
    native writer (``glpb_write_lp``, csrc/lpformat.cpp); ``_write_lp_py`` below is the Python restatement
    kept as its cross-check (tests/test_writer_golden.py: both equal the reference's own writer)."""
    _check(P, "glp_write_lp")
    xprintf("Writing problem data")
    if not (P.m > 0 and P.n > 0):
        xprintf("Warning: problem has no rows/columns")
    d, _, _ = _arrays(P)
    rptr, rind, rval = _csr(P)
    col_len = [len(P.col[j].elems) for j in range(1, P.n + 1)]
    names = (P.obj, [P.row[i].name for i in range(1, P.m + 1)], [P.col[j].name for j in range(1, P.n + 1)])
    lines, count = native.write_lp(d, col_len, rptr, rind, rval, P.name, names)
    for line in lines:
        callback(line)
    xprintf("%d lines were written" % count)
    return 0


def _write_lp_py(P, parm, callback):
    """lib/glpcpx.js:755-998 restated in Python (cross-check of the native writer).
    (The reference's adjust_name assigns into an immutable string and so changes
    nothing: a name with a blank or a dash is replaced by r_i / x_j.)"""
    _check(P, "glp_write_lp")

    def valid(name):
        if name is None or name == "" or name[0] == "." or name[0].isdigit():
            return False
        return all((ch.isalnum() and ch.isascii()) or ch in _NAME_EXTRA for ch in name)

    def row_name(i):
        name = P.obj if i == 0 else P.row[i].name
        return name if valid(name) else ("obj" if i == 0 else "r_%d" % i)

    def col_name(j):
        name = P.col[j].name
        return name if valid(name) else "x_%d" % j

    count = [0]

    def out(line):
        callback(line)
        count[0] += 1

    def finish():
        out("End")
        xprintf("%d lines were written" % count[0])
        return 0

    class Line:
        def __init__(self, text):
            self.text = text

        def add(self, term):
            if len(self.text) + len(term) > 72:
                out(self.text)
                self.text = ""
            self.text += term

    def signed_term(v, name):
        if v == +1.0:
            return " + " + name
        if v == -1.0:
            return " - " + name
        if v > 0.0:
            return " + %s %s" % (_num(v), name)
        return " - %s %s" % (_num(-v), name)

    xprintf("Writing problem data")
    out("\\* Problem: %s *\\" % ("Unknown" if P.name is None else P.name))
    out("")
    if not (P.m > 0 and P.n > 0):
        xprintf("Warning: problem has no rows/columns")
        out("\\* WARNING: PROBLEM HAS NO ROWS/COLUMNS *\\")
        out("")
        return finish()
    out("Minimize" if P.dir == GLP_MIN else "Maximize")
    line = Line(" " + row_name(0) + ":")
    terms = 0
    for j in range(1, P.n + 1):
        col = P.col[j]
        if col.coef != 0.0 or not col.elems:
            terms += 1
            line.add((" + 0 " + col_name(j)) if col.coef == 0.0 else signed_term(col.coef, col_name(j)))
    if terms == 0:
        line.text += " 0 " + col_name(1)
    out(line.text)
    if P.c0 != 0.0:
        out("\\* constant term = %s *\\" % _num(P.c0))
    out("")
    out("Subject To")
    for i in range(1, P.m + 1):
        row = P.row[i]
        if row.type == GLP_FR:
            continue
        line = Line(" " + row_name(i) + ":")
        for (j, v) in row.elems:
            line.add(signed_term(v, col_name(j)))
        if row.type == GLP_DB:
            line.add(" - ~r_%d" % i)
        elif not row.elems:
            line.text += " 0 " + col_name(1)
        if row.type == GLP_LO:
            line.add(" >= " + _num(row.lb))
        elif row.type == GLP_UP:
            line.add(" <= " + _num(row.ub))
        else:
            line.add(" = " + _num(row.lb))
        out(line.text)
    out("")
    flag = False
    for i in range(1, P.m + 1):
        row = P.row[i]
        if row.type != GLP_DB:
            continue
        if not flag:
            out("Bounds")
            flag = True
        out(" 0 <= ~r_%d <= %s" % (i, _num(row.ub - row.lb)))
    for j in range(1, P.n + 1):
        col = P.col[j]
        if col.type == GLP_LO and col.lb == 0.0:
            continue
        if not flag:
            out("Bounds")
            flag = True
        name = col_name(j)
        if col.type == GLP_FR:
            out(" %s free" % name)
        elif col.type == GLP_LO:
            out(" %s >= %s" % (name, _num(col.lb)))
        elif col.type == GLP_UP:
            out(" -Inf <= %s <= %s" % (name, _num(col.ub)))
        elif col.type == GLP_DB:
            out(" %s <= %s <= %s" % (_num(col.lb), name, _num(col.ub)))
        else:
            out(" %s = %s" % (name, _num(col.lb)))
    if flag:
        callback("")
    count[0] += 1
    flag = False
    for j in range(1, P.n + 1):
        if P.col[j].kind == GLP_CV:
            continue
        if not flag:
            out("Generals")
            flag = True
        out(" " + col_name(j))
    if flag:
        out("")
    return finish()


def _arrays(P):
    m, n = P.m, P.n
    type_ = np.zeros(m + n, np.int32)
    lb = np.zeros(m + n)
    ub = np.zeros(m + n)
    for i in range(1, m + 1):
        r = P.row[i]
        type_[i - 1], lb[i - 1], ub[i - 1] = r.type, r.lb, r.ub
    coef = np.zeros(n)
    kind = np.zeros(n, np.int32)
    ptr = np.zeros(n + 1, np.int32)
    ind, val = [], []
    for j in range(1, n + 1):
        c = P.col[j]
        type_[m + j - 1], lb[m + j - 1], ub[m + j - 1] = c.type, c.lb, c.ub
        coef[j - 1], kind[j - 1] = c.coef, c.kind
        ptr[j - 1] = len(ind)
        for (i, v) in c.elems:
            ind.append(i - 1)
            val.append(v)
    ptr[n] = len(ind)
    rii = np.array([P.row[i].rii for i in range(1, m + 1)])
    sjj = np.array([P.col[j].sjj for j in range(1, n + 1)])
    return dict(m=m, n=n, dir=P.dir, c0=P.c0, type=type_, lb=lb, ub=ub, coef=coef, kind=kind,
                A_ptr=ptr, A_ind=np.array(ind, np.int32), A_val=np.array(val, np.float64)), rii, sjj


def _device(P, device=0):
    """(Re)build the device handle when the matrix changed; otherwise push only
    bounds and statuses, which is all glp_set_*_bnds / glp_set_*_stat touch."""
    d, rii, sjj = _arrays(P)
    if P._dev is None or P._dirty:
        if P._dev is not None:
            P._dev.close()
        P._dev = native.Problem(d, device=device, rii=rii, sjj=sjj)
        P._dirty = False
        P.valid = 0
        if P.bfcp is not None:
            _push_bfcp(P)
    else:
        k = np.arange(1, P.m + P.n + 1, dtype=np.int32)
        P._dev.set_bounds(k, d["type"], d["lb"], d["ub"])
    stat = np.array([P.row[i].stat for i in range(1, P.m + 1)] +
                    [P.col[j].stat for j in range(1, P.n + 1)], np.int32)
    if not P.valid:
        P._dev.set_basis(stat)
    else:
        # the basis is unchanged: only non-basic statuses may differ
        P._dev.set_basis(stat)
    P._dev.L.glpb_set_it_cnt(P._dev.h, int(P.it_cnt))
    return P._dev


def _pull(P, dev):
    s = dev.solution()
    m, n = P.m, P.n
    P.pbs_stat, P.dbs_stat = s["pbs"], s["dbs"]
    P.obj_val, P.it_cnt, P.some = s["obj"], s["it_cnt"], s["some"]
    bind = {int(k): i + 1 for i, k in enumerate(s["head"])}
    P.head = [0] + [int(k) for k in s["head"]]
    for i in range(1, m + 1):
        r = P.row[i]
        r.stat, r.prim, r.dual = int(s["stat"][i - 1]), float(s["prim"][i - 1]), float(s["dual"][i - 1])
        r.bind = bind.get(i, 0)
    for j in range(1, n + 1):
        c = P.col[j]
        k = m + j
        c.stat, c.prim, c.dual = int(s["stat"][k - 1]), float(s["prim"][k - 1]), float(s["dual"][k - 1])
        c.bind = bind.get(k, 0)


def _check_smcp(parm):
    # lib/glpapi06.js:271-300
    def bad(name):
        xerror("glp_simplex: %s = %r; invalid parameter" % (name, getattr(parm, name)))
    if parm.msg_lev not in (GLP_MSG_OFF, GLP_MSG_ERR, GLP_MSG_ON, GLP_MSG_ALL, GLP_MSG_DBG): bad("msg_lev")
    if parm.meth not in (GLP_PRIMAL, GLP_DUALP, GLP_DUAL): bad("meth")
    if parm.pricing not in (GLP_PT_STD, GLP_PT_PSE): bad("pricing")
    if parm.r_test not in (GLP_RT_STD, GLP_RT_HAR): bad("r_test")
    if not (0.0 < parm.tol_bnd < 1.0): bad("tol_bnd")
    if not (0.0 < parm.tol_dj < 1.0): bad("tol_dj")
    if not (0.0 < parm.tol_piv < 1.0): bad("tol_piv")
    if parm.it_lim < 0: bad("it_lim")
    if parm.tm_lim < 0: bad("tm_lim")
    if parm.out_frq < 1: bad("out_frq")
    if parm.out_dly < 0: bad("out_dly")
    if parm.presolve not in (GLP_ON, GLP_OFF): bad("presolve")


# ---- basis factorisation interface (lib/glpapi12.js:1-244) ----
GLP_BF_FT, GLP_BF_BG, GLP_BF_GR = 1, 2, 3
_BFCP_DEFAULTS = dict(type=GLP_BF_FT, lu_size=0, piv_tol=0.10, piv_lim=4, suhl=GLP_ON, eps_tol=1e-15,
                      max_gro=1e+10, nfs_max=100, upd_tol=1e-6, nrs_max=100, rs_size=0)


def glp_bf_exists(P):
    """lib/glpapi12.js:1-3"""
    return bool(P.m == 0 or P.valid)


def glp_get_bfcp(P, parm):
    """lib/glpapi12.js:108-126: fills the dict ``parm``"""
    parm.update(_BFCP_DEFAULTS if P.bfcp is None else P.bfcp)


def glp_set_bfcp(P, parm):
    """lib/glpapi12.js:134-176.  The device keeps an explicit inverse of the
    structural kernel of B, so of these only nfs_max (refactorisation period),
    piv_tol and upd_tol reach it (glpb_set_bfcp); the rest is validated and kept."""
    if parm is None:
        P.bfcp = None
    else:
        bfcp = dict(_BFCP_DEFAULTS if P.bfcp is None else P.bfcp)
        bfcp.update(parm)

        def bad(name):
            xerror("glp_set_bfcp: %s = %s; invalid parameter" % (name, _num(bfcp[name])))
        if bfcp["type"] not in (GLP_BF_FT, GLP_BF_BG, GLP_BF_GR): bad("type")
        if bfcp["lu_size"] < 0: bad("lu_size")
        if not (0.0 < bfcp["piv_tol"] < 1.0): bad("piv_tol")
        if bfcp["piv_lim"] < 1: bad("piv_lim")
        if bfcp["suhl"] not in (GLP_ON, GLP_OFF): bad("suhl")
        if not (0.0 <= bfcp["eps_tol"] <= 1e-6): bad("eps_tol")
        if bfcp["max_gro"] < 1.0: bad("max_gro")
        if not (1 <= bfcp["nfs_max"] <= 32767): bad("nfs_max")
        if not (0.0 < bfcp["upd_tol"] < 1.0): bad("upd_tol")
        if not (1 <= bfcp["nrs_max"] <= 32767): bad("nrs_max")
        if bfcp["rs_size"] < 0: bad("rs_size")
        if bfcp["rs_size"] == 0:
            bfcp["rs_size"] = 20 * bfcp["nrs_max"]
        P.bfcp = bfcp
    if P._dev is not None:
        _push_bfcp(P)


def _push_bfcp(P):
    b = _BFCP_DEFAULTS if P.bfcp is None else P.bfcp
    P._dev.set_bfcp(nfs_max=int(b["nfs_max"]), piv_tol=float(b["piv_tol"]), upd_tol=float(b["upd_tol"]))


def glp_factorize(P, device=0):
    """lib/glpapi12.js:5-100: basis header from the statuses (basic variables in
    the order k = 1..m+n), then the factorisation -- glpb_factorize on the handle."""
    _check(P, "glp_factorize")
    m, n = P.m, P.n
    P.valid = 0
    head = [0] * (1 + m)
    j = 0
    for k in range(1, m + n + 1):
        x = P.row[k] if k <= m else P.col[k - m]
        x.bind = 0
        if x.stat == GLP_BS:
            j += 1
            if j > m:
                return GLP_EBADB
            head[j] = k
            x.bind = j
    if j < m:
        return GLP_EBADB
    P.head = head
    if m > 0:
        if n == 0 or P.nnz == 0:
            xerror("glp_factorize: problems without constraint coefficients are handled by the host binding only")
        dev = _device(P, device)
        ret = dev.factorize()
        if ret < 0:
            xerror("glp_factorize: device error %d: %s" % (ret, native.last_error()))
        if ret != 0:
            return ret          # GLP_EBADB / GLP_ESING / GLP_ECOND
        P.valid = 1
    return 0


def _need_bf(P, who):
    if not (P.m == 0 or P.valid):
        xerror("%s: basis factorization does not exist" % who)


def glp_get_bhead(P, k):
    _need_bf(P, "glp_get_bhead")
    if not (1 <= k <= P.m):
        xerror("glp_get_bhead: k = %d; index out of range" % k)
    return P.head[k]


def glp_get_row_bind(P, i):
    _need_bf(P, "glp_get_row_bind")
    return _row(P, i, "glp_get_row_bind").bind


def glp_get_col_bind(P, j):
    _need_bf(P, "glp_get_col_bind")
    return _col(P, j, "glp_get_col_bind").bind


def _basis_scale(P):
    """SB of lib/glpapi12.js:178-220: 1/rii for a basic auxiliary variable, sjj for a
    basic structural one, by basis position"""
    m = P.m
    return np.array([(1.0 / P.row[k].rii) if k <= m else P.col[k - m].sjj for k in P.head[1:m + 1]])


def glp_ftran(P, x):
    """lib/glpapi12.js:178-199: solves B x = b in place (x[1..m], slot 0 unused);
    the scaled solve inv(R B SB) runs on the device (glpb_ftran)."""
    _need_bf(P, "glp_ftran")
    m = P.m
    if m == 0:
        return
    rii = np.array([P.row[i].rii for i in range(1, m + 1)])
    b = np.array([float(x[i]) for i in range(1, m + 1)]) * rii
    y = P._dev.ftran(b) * _basis_scale(P)
    for i in range(1, m + 1):
        x[i] = float(y[i - 1])


def glp_btran(P, x):
    """lib/glpapi12.js:201-222: solves B' x = b in place"""
    _need_bf(P, "glp_btran")
    m = P.m
    if m == 0:
        return
    rii = np.array([P.row[i].rii for i in range(1, m + 1)])
    b = np.array([float(x[i]) for i in range(1, m + 1)]) * _basis_scale(P)
    y = P._dev.btran(b) * rii
    for i in range(1, m + 1):
        x[i] = float(y[i - 1])


def glp_warm_up(P, device=0):
    """lib/glpapi12.js:244-398: primal and dual values of the CURRENT basis, without
    iterating: x_B = -inv(B) N x_N, pi = inv(B') c_B, d_N = c_N - N' pi; the two solves
    are glp_ftran / glp_btran on the device's factorisation."""
    _check(P, "glp_warm_up")
    m, n = P.m, P.n
    P.pbs_stat = P.dbs_stat = GLP_UNDEF
    P.obj_val, P.some = 0.0, 0
    for x in P.row[1:] + P.col[1:]:
        x.prim = x.dual = 0.0
    if not glp_bf_exists(P):
        ret = glp_factorize(P, device)
        if ret != 0:
            return ret
    work = [0.0] * (1 + m)
    for i in range(1, m + 1):
        row = P.row[i]
        if row.stat == GLP_BS:
            continue
        row.prim = _bound_value(row)
        work[i] -= row.prim
    for j in range(1, n + 1):
        col = P.col[j]
        if col.stat == GLP_BS:
            continue
        col.prim = _bound_value(col)
        if col.prim != 0.0:
            for (i, v) in col.elems:
                work[i] += v * col.prim
    glp_ftran(P, work)
    P.pbs_stat = GLP_FEAS
    for x in P.row[1:] + P.col[1:]:
        if x.stat != GLP_BS:
            continue
        x.prim = work[x.bind]
        if x.type in (GLP_LO, GLP_DB, GLP_FX) and x.prim < x.lb - (1e-6 + 1e-9 * abs(x.lb)):
            P.pbs_stat = GLP_INFEAS
        if x.type in (GLP_UP, GLP_DB, GLP_FX) and x.prim > x.ub + (1e-6 + 1e-9 * abs(x.ub)):
            P.pbs_stat = GLP_INFEAS
    P.obj_val = P.c0
    for col in P.col[1:]:
        P.obj_val += col.coef * col.prim
    work = [0.0] * (1 + m)
    for col in P.col[1:]:
        if col.stat == GLP_BS:
            work[col.bind] = col.coef
    glp_btran(P, work)
    P.dbs_stat = GLP_FEAS

    def dual_ok(x):
        temp = +x.dual if P.dir == GLP_MIN else -x.dual
        if (x.stat in (GLP_NF, GLP_NL) and temp < -1e-5) or (x.stat in (GLP_NF, GLP_NU) and temp > +1e-5):
            P.dbs_stat = GLP_INFEAS

    for i in range(1, m + 1):
        row = P.row[i]
        if row.stat == GLP_BS:
            row.dual = 0.0
            continue
        row.dual = -work[i]
        dual_ok(row)
    for col in P.col[1:]:
        if col.stat == GLP_BS:
            col.dual = 0.0
            continue
        col.dual = col.coef
        for (i, v) in col.elems:
            col.dual += v * work[i]
        dual_ok(col)
    return 0


def glp_eval_tab_row(P, k, ind, val):
    """lib/glpapi12.js:401-453: row of the simplex table of basic x[k] (ind/val 1-based out)"""
    m, n = P.m, P.n
    _need_bf(P, "glp_eval_tab_row")
    if not (1 <= k <= m + n):
        xerror("glp_eval_tab_row: k = %d; variable number out of range" % k)
    i = glp_get_row_bind(P, k) if k <= m else glp_get_col_bind(P, k - m)
    if i == 0:
        xerror("glp_eval_tab_row: k = %d; variable must be basic" % k)
    rho = [0.0] * (1 + m)
    rho[i] = 1.0
    glp_btran(P, rho)
    ln = 0
    for kk in range(1, m + n + 1):
        if kk <= m:
            if P.row[kk].stat == GLP_BS:
                continue
            alfa = -rho[kk]
        else:
            col = P.col[kk - m]
            if col.stat == GLP_BS:
                continue
            alfa = 0.0
            for (r, v) in col.elems:
                alfa += rho[r] * v
        if alfa != 0.0:
            ln += 1
            ind[ln], val[ln] = kk, alfa
    return ln


def glp_eval_tab_col(P, k, ind, val):
    """lib/glpapi12.js:455-497: column of the simplex table of non-basic x[k]"""
    m, n = P.m, P.n
    _need_bf(P, "glp_eval_tab_col")
    if not (1 <= k <= m + n):
        xerror("glp_eval_tab_col: k = %d; variable number out of range" % k)
    stat = P.row[k].stat if k <= m else P.col[k - m].stat
    if stat == GLP_BS:
        xerror("glp_eval_tab_col: k = %d; variable must be non-basic" % k)
    col = [0.0] * (1 + m)
    if k <= m:
        col[k] = -1.0
    else:
        for (r, v) in P.col[k - m].elems:
            col[r] = v
    glp_ftran(P, col)
    ln = 0
    for t in range(1, m + 1):
        if col[t] != 0.0:
            ln += 1
            ind[ln], val[ln] = P.head[t], col[t]
    return ln


def glp_transform_row(P, length, ind, val):
    """lib/glpapi12.js:499-556: an explicit row sum a_j x_j expressed in the non-basic variables"""
    _need_bf(P, "glp_transform_row")
    m, n = P.m, P.n
    if not (0 <= length <= n):
        xerror("glp_transform_row: len = %d; invalid row length" % length)
    a = [0.0] * (1 + n)
    for t in range(1, length + 1):
        j = ind[t]
        if not (1 <= j <= n):
            xerror("glp_transform_row: ind[%d] = %d; column index out of range" % (t, j))
        if val[t] == 0.0:
            xerror("glp_transform_row: val[%d] = 0; zero coefficient not allowed" % t)
        if a[j] != 0.0:
            xerror("glp_transform_row: ind[%d] = %d; duplicate column indices not allowed" % (t, j))
        a[j] = val[t]
    rho = [0.0] * (1 + m)
    for i in range(1, m + 1):
        k = P.head[i]
        rho[i] = 0.0 if k <= m else a[k - m]
    glp_btran(P, rho)
    ln = 0
    for i in range(1, m + 1):
        if P.row[i].stat != GLP_BS:
            alfa = -rho[i]
            if alfa != 0.0:
                ln += 1
                ind[ln], val[ln] = i, alfa
    for j in range(1, n + 1):
        col = P.col[j]
        if col.stat != GLP_BS:
            alfa = a[j]
            for (r, v) in col.elems:
                alfa += v * rho[r]
            if alfa != 0.0:
                ln += 1
                ind[ln], val[ln] = m + j, alfa
    return ln


def glp_transform_col(P, length, ind, val):
    """lib/glpapi12.js:558-591: an explicit column expressed in the basic variables"""
    _need_bf(P, "glp_transform_col")
    m = P.m
    if not (0 <= length <= m):
        xerror("glp_transform_col: len = %d; invalid column length" % length)
    a = [0.0] * (1 + m)
    for t in range(1, length + 1):
        i = ind[t]
        if not (1 <= i <= m):
            xerror("glp_transform_col: ind[%d] = %d; row index out of range" % (t, i))
        if val[t] == 0.0:
            xerror("glp_transform_col: val[%d] = 0; zero coefficient not allowed" % t)
        if a[i] != 0.0:
            xerror("glp_transform_col: ind[%d] = %d; duplicate row indices not allowed" % (t, i))
        a[i] = val[t]
    glp_ftran(P, a)
    ln = 0
    for i in range(1, m + 1):
        if a[i] != 0.0:
            ln += 1
            ind[ln], val[ln] = P.head[i], a[i]
    return ln


def _var(P, k):
    return P.row[k] if k <= P.m else P.col[k - P.m]


def glp_prim_rtest(P, length, ind, val, dir, eps):
    """lib/glpapi12.js:593-685: textbook primal ratio test over an explicit column;
    returns the position t of the pivot in ind/val or 0"""
    if glp_get_prim_stat(P) != GLP_FEAS:
        xerror("glp_prim_rtest: basic solution is not primal feasible ")
    if dir not in (+1, -1):
        xerror("glp_prim_rtest: dir = %r; invalid parameter" % (dir,))
    if not (0.0 < eps < 1.0):
        xerror("glp_prim_rtest: eps = %s; invalid parameter" % _num(eps))
    piv, teta, big = 0, DBL_MAX, 0.0
    for t in range(1, length + 1):
        k = ind[t]
        if not (1 <= k <= P.m + P.n):
            xerror("glp_prim_rtest: ind[%d] = %d; variable number out of range" % (t, k))
        x = _var(P, k)
        if x.stat != GLP_BS:
            xerror("glp_prim_rtest: ind[%d] = %d; non-basic variable not allowed" % (t, k))
        alfa = +val[t] if dir > 0 else -val[t]
        if x.type == GLP_FR:
            continue
        elif x.type == GLP_LO or (x.type == GLP_DB and alfa < 0.0):
            if alfa > -eps:
                continue
            temp = (x.lb - x.prim) / alfa
        elif x.type == GLP_UP or x.type == GLP_DB:
            if alfa < +eps:
                continue
            temp = (x.ub - x.prim) / alfa
        else:
            if -eps < alfa < +eps:
                continue
            temp = 0.0
        if temp < 0.0:
            temp = 0.0
        if teta > temp or (teta == temp and big < abs(alfa)):
            piv, teta, big = t, temp, abs(alfa)
    return piv


def glp_dual_rtest(P, length, ind, val, dir, eps):
    """lib/glpapi12.js:687-762: textbook dual ratio test over an explicit row"""
    if glp_get_dual_stat(P) != GLP_FEAS:
        xerror("glp_dual_rtest: basic solution is not dual feasible")
    if dir not in (+1, -1):
        xerror("glp_dual_rtest: dir = %r; invalid parameter" % (dir,))
    if not (0.0 < eps < 1.0):
        xerror("glp_dual_rtest: eps = %s; invalid parameter" % _num(eps))
    obj = +1.0 if P.dir == GLP_MIN else -1.0
    piv, teta, big = 0, DBL_MAX, 0.0
    for t in range(1, length + 1):
        k = ind[t]
        if not (1 <= k <= P.m + P.n):
            xerror("glp_dual_rtest: ind[%d] = %d; variable number out of range" % (t, k))
        x = _var(P, k)
        if x.stat == GLP_BS:
            xerror("glp_dual_rtest: ind[%d] = %d; basic variable not allowed" % (t, k))
        alfa = +val[t] if dir > 0 else -val[t]
        if x.stat == GLP_NL:
            if alfa < +eps:
                continue
            temp = (obj * x.dual) / alfa
        elif x.stat == GLP_NU:
            if alfa > -eps:
                continue
            temp = (obj * x.dual) / alfa
        elif x.stat == GLP_NF:
            if -eps < alfa < +eps:
                continue
            temp = 0.0
        else:
            continue
        if temp < 0.0:
            temp = 0.0
        if teta > temp or (teta == temp and big < abs(alfa)):
            piv, teta, big = t, temp, abs(alfa)
    return piv


def _plural(k, word):
    return "%d %s%s" % (k, word, "" if k == 1 else "s")


def _size_line(P):
    return "%s, %s, %s" % (_plural(P.m, "row"), _plural(P.n, "column"), _plural(P.nnz, "non-zero"))


def _check_db_bounds(P, who, msg_lev):
    for i in range(1, P.m + 1):
        r = P.row[i]
        if r.type == GLP_DB and r.lb >= r.ub:
            if msg_lev >= GLP_MSG_ERR:
                xprintf("%s: row %d: lb = %s, ub = %s; incorrect bounds" % (who, i, _num(r.lb), _num(r.ub)))
            return GLP_EBOUND
    for j in range(1, P.n + 1):
        c = P.col[j]
        if c.type == GLP_DB and c.lb >= c.ub:
            if msg_lev >= GLP_MSG_ERR:
                xprintf("%s: column %d: lb = %s, ub = %s; incorrect bounds" % (who, j, _num(c.lb), _num(c.ub)))
            return GLP_EBOUND
    return 0


def _trivial_lp(P, parm):
    """lib/glpapi06.js:148-255: LP with an empty constraint matrix, solved on the
    host (there is nothing to launch)."""
    P.valid = 0
    P.pbs_stat = P.dbs_stat = GLP_FEAS
    P.obj_val = P.c0
    P.some = 0
    p_infeas = d_infeas = 0.0
    for i in range(1, P.m + 1):
        row = P.row[i]
        row.stat = GLP_BS
        row.prim = row.dual = 0.0
        if row.type in (GLP_LO, GLP_DB, GLP_FX):
            if row.lb > +parm.tol_bnd:
                P.pbs_stat = GLP_NOFEAS
                if P.some == 0 and parm.meth != GLP_PRIMAL:
                    P.some = i
            if p_infeas < +row.lb:
                p_infeas = +row.lb
        if row.type in (GLP_UP, GLP_DB, GLP_FX):
            if row.ub < -parm.tol_bnd:
                P.pbs_stat = GLP_NOFEAS
                if P.some == 0 and parm.meth != GLP_PRIMAL:
                    P.some = i
            if p_infeas < -row.ub:
                p_infeas = -row.ub
    zeta = 1.0
    for j in range(1, P.n + 1):
        zeta = max(zeta, abs(P.col[j].coef))
    zeta = (+1.0 if P.dir == GLP_MIN else -1.0) / zeta
    for j in range(1, P.n + 1):
        col = P.col[j]
        if col.type == GLP_FR:
            col.stat, col.prim = GLP_NF, 0.0
        elif col.type == GLP_FX:
            col.stat, col.prim = GLP_NS, col.lb
        else:
            if col.type == GLP_LO:
                at_lower = True
            elif col.type == GLP_UP:
                at_lower = False
            elif zeta * col.coef > 0.0:
                at_lower = True
            elif zeta * col.coef < 0.0:
                at_lower = False
            else:
                at_lower = abs(col.lb) <= abs(col.ub)
            col.stat, col.prim = (GLP_NL, col.lb) if at_lower else (GLP_NU, col.ub)
        col.dual = col.coef
        P.obj_val += col.coef * col.prim
        if col.type in (GLP_FR, GLP_LO):
            if zeta * col.dual < -parm.tol_dj:
                P.dbs_stat = GLP_NOFEAS
                if P.some == 0 and parm.meth == GLP_PRIMAL:
                    P.some = P.m + j
            d_infeas = max(d_infeas, -zeta * col.dual)
        if col.type in (GLP_FR, GLP_UP):
            if zeta * col.dual > +parm.tol_dj:
                P.dbs_stat = GLP_NOFEAS
                if P.some == 0 and parm.meth == GLP_PRIMAL:
                    P.some = P.m + j
            d_infeas = max(d_infeas, +zeta * col.dual)
    if parm.msg_lev >= GLP_MSG_ON and parm.out_dly == 0:
        xprintf("~%d: obj = %s  infeas = %s" % (P.it_cnt, _num(P.obj_val),
                                                _num(p_infeas if parm.meth == GLP_PRIMAL else d_infeas)))
    if parm.msg_lev >= GLP_MSG_ALL and parm.out_dly == 0:
        if P.pbs_stat == GLP_FEAS and P.dbs_stat == GLP_FEAS:
            xprintf("OPTIMAL SOLUTION FOUND")
        elif P.pbs_stat == GLP_NOFEAS:
            xprintf("PROBLEM HAS NO FEASIBLE SOLUTION")
        elif parm.meth == GLP_PRIMAL:
            xprintf("PROBLEM HAS UNBOUNDED SOLUTION")
        else:
            xprintf("PROBLEM HAS NO DUAL FEASIBLE SOLUTION")


def _solve_lp(P, parm, device):
    """lib/glpapi06.js:3-39 -- the drop-in boundary: glp_factorize + spx_primal /
    spx_dual run behind glpb_simplex on the device-resident handle."""
    dev = _device(P, device)
    sp = dev.smcp(msg_lev=parm.msg_lev, meth=parm.meth, pricing=parm.pricing, r_test=parm.r_test,
                  tol_bnd=parm.tol_bnd, tol_dj=parm.tol_dj, tol_piv=parm.tol_piv, obj_ll=parm.obj_ll,
                  obj_ul=parm.obj_ul, it_lim=int(parm.it_lim), tm_lim=int(parm.tm_lim),
                  out_frq=parm.out_frq, out_dly=parm.out_dly, presolve=GLP_OFF)
    ret = dev.simplex(sp)
    _pull(P, dev)
    P.valid = 1 if ret in (0, GLP_EOBJLL, GLP_EOBJUL, GLP_EITLIM, GLP_ETMLIM) else P.valid
    return ret


def _npp_load(P, sol):
    """npp_create_wksp + npp_load_prob(npp, P, GLP_OFF, sol, GLP_OFF) (lib/glpnpp01.js:262-394): the
    problem as it stands -- unscaled, every column in list order -- into a native presolver workspace
    (csrc/presolve.cpp)."""
    d, _, _ = _arrays(P)
    return native.Presolver(d, sol)


def _npp_build(P, npp):
    """npp_build_prob (lib/glpnpp01.js:396-472): the reduced problem as a fresh problem object, built
    through the same API calls in the same order, so rows, columns and both element lists end up in the
    reference's state.  Names are not carried (the callers load with names = GLP_OFF)."""
    r = npp.build()
    Q = glp_prob()
    Q.dir = P.dir
    Q.c0 = float(r["c0"])
    m, n = r["m"], r["n"]
    if m:
        glp_add_rows(Q, m)
    for i in range(1, m + 1):
        glp_set_row_bnds(Q, i, int(r["type"][i - 1]), float(r["lb"][i - 1]), float(r["ub"][i - 1]))
    if n:
        glp_add_cols(Q, n)
    ptr = r["A_ptr"]
    for j in range(1, n + 1):
        k = m + j - 1
        glp_set_col_kind(Q, j, GLP_IV if int(r["kind"][j - 1]) == GLP_IV else GLP_CV)
        glp_set_col_bnds(Q, j, int(r["type"][k]), float(r["lb"][k]), float(r["ub"][k]))
        Q.col[j].coef = float(r["coef"][j - 1])
        a, b = int(ptr[j - 1]), int(ptr[j])
        glp_set_mat_col(Q, j, b - a, [0] + [int(i) + 1 for i in r["A_ind"][a:b]],
                        [0.0] + [float(v) for v in r["A_val"][a:b]])
    Q.bfcp = None if P.bfcp is None else dict(P.bfcp)   # inherited (lib/glpapi06.js:103-105)
    Q._npp_ref = (r["row_ref"], r["col_ref"])
    return Q


def _quiet(msg_lev, fn, *args):
    """lib/glpapi06.js:107-128 / lib/glpapi09.js:201-213 switch env.term_out off around scaling and the
    crash basis unless msg_lev >= GLP_MSG_ALL -- but the reference's xprintf (lib/glpapi.js:30-35) never
    reads env.term_out, so those messages reach the print function at every msg_lev (the reference run
    under minijs prints 'Scaling...' ... 'Size of triangular part = 5' with msg_lev = GLP_MSG_OFF).
    Mirrored as it behaves, not as it was meant."""
    del msg_lev
    return fn(*args)


def _bound_value(x):
    return {GLP_NL: x.lb, GLP_NU: x.ub, GLP_NF: 0.0, GLP_NS: x.lb}[x.stat]


def _postprocess_basic(npp, lp):
    """npp_postprocess, basic solution (lib/glpnpp01.js:474-570)"""
    r_stat = [lp.row[i].stat for i in range(1, lp.m + 1)]
    r_dual = [lp.row[i].dual for i in range(1, lp.m + 1)]
    c_stat = [lp.col[j].stat for j in range(1, lp.n + 1)]
    c_prim = [lp.col[j].prim for j in range(1, lp.n + 1)]
    return npp.postprocess(c_prim, r_stat, r_dual, c_stat)


def _unload_basic(P, p_stat, d_stat, r_stat, r_dual, c_stat, c_value):
    """npp_unload_sol, basic solution (lib/glpnpp01.js:589-682): statuses, row duals and column values
    come from the recovery, the rest is recomputed from the ORIGINAL coefficients."""
    P.valid = 0
    P.pbs_stat, P.dbs_stat = p_stat, d_stat
    P.obj_val = P.c0
    P.some = 0
    for i in range(1, P.m + 1):
        row = P.row[i]
        row.stat = int(r_stat[i - 1])
        row.dual = float(r_dual[i - 1])
        if row.stat == GLP_BS:
            row.dual = 0.0
        else:
            row.prim = _bound_value(row)
    for j in range(1, P.n + 1):
        col = P.col[j]
        col.stat = int(c_stat[j - 1])
        col.prim = float(c_value[j - 1])
        if col.stat == GLP_BS:
            col.dual = 0.0
        else:
            col.prim = _bound_value(col)
        P.obj_val += col.coef * col.prim
    for i in range(1, P.m + 1):
        row = P.row[i]
        if row.stat == GLP_BS:
            temp = 0.0
            for (j, v) in row.elems:
                temp += v * P.col[j].prim
            row.prim = temp
    for j in range(1, P.n + 1):
        col = P.col[j]
        if col.stat != GLP_BS:
            temp = col.coef
            for (i, v) in col.elems:
                temp -= v * P.row[i].dual
            col.dual = temp


def _unload_mip(P, mip_stat, c_value):
    """npp_unload_sol, MIP solution (lib/glpnpp01.js:734-756)"""
    P.mip_stat = mip_stat
    P.mip_obj = P.c0
    for j in range(1, P.n + 1):
        col = P.col[j]
        col.mipx = float(c_value[j - 1])
        P.mip_obj += col.coef * col.mipx
    for i in range(1, P.m + 1):
        temp = 0.0
        for (j, v) in P.row[i].elems:
            temp += v * P.col[j].mipx
        P.row[i].mipx = temp


def _drop_device(P):
    if P._dev is not None:
        P._dev.close()
        P._dev = None


def _preprocess_and_solve_lp(P, parm, device):
    """lib/glpapi06.js:40-146: presolve in the native library (csrc/presolve.cpp), automatic scaling and
    triangular crash basis of the REDUCED problem, solve on the device, recovery of the solution."""
    if parm.msg_lev >= GLP_MSG_ALL:
        xprintf("Preprocessing...")
    npp = _npp_load(P, GLP_SOL)
    try:
        ret = npp.simplex()
        if ret != 0:
            if parm.msg_lev >= GLP_MSG_ALL:
                xprintf("PROBLEM HAS NO PRIMAL FEASIBLE SOLUTION" if ret == GLP_ENOPFS
                        else "PROBLEM HAS NO DUAL FEASIBLE SOLUTION")
            return ret
        lp = _npp_build(P, npp)
        if lp.m == 0 and lp.n == 0:
            lp.pbs_stat = lp.dbs_stat = GLP_FEAS
            lp.obj_val = lp.c0
            if parm.msg_lev >= GLP_MSG_ON and parm.out_dly == 0:
                xprintf("%d: obj = %s  infeas = 0.0" % (P.it_cnt, _num(lp.obj_val)))
            if parm.msg_lev >= GLP_MSG_ALL:
                xprintf("OPTIMAL SOLUTION FOUND BY LP PREPROCESSOR")
        else:
            if parm.msg_lev >= GLP_MSG_ALL:
                xprintf(_size_line(lp))
            try:
                _quiet(parm.msg_lev, glp_scale_prob, lp, GLP_SF_AUTO)
                _quiet(parm.msg_lev, glp_adv_basis, lp, 0)
                lp.it_cnt = P.it_cnt
                ret = _solve_lp(lp, parm, device)
                P.it_cnt = lp.it_cnt
            finally:
                _drop_device(lp)
            if not (ret == 0 and lp.pbs_stat == GLP_FEAS and lp.dbs_stat == GLP_FEAS):
                if parm.msg_lev >= GLP_MSG_ERR:
                    xprintf("glp_simplex: unable to recover undefined or non-optimal solution")
                if ret == 0:
                    if lp.pbs_stat == GLP_NOFEAS:
                        ret = GLP_ENOPFS
                    elif lp.dbs_stat == GLP_NOFEAS:
                        ret = GLP_ENODFS
                return ret
        r_stat, r_dual, c_stat, c_value = _postprocess_basic(npp, lp)
        _unload_basic(P, lp.pbs_stat, lp.dbs_stat, r_stat, r_dual, c_stat, c_value)
        return 0
    finally:
        npp.close()


def glp_simplex(P, parm=None, device=0):
    """lib/glpapi06.js:261-339"""
    _check(P, "glp_simplex")
    if parm is None:
        parm = SMCP()
    _check_smcp(parm)
    P.pbs_stat = P.dbs_stat = GLP_UNDEF
    P.obj_val, P.some = 0.0, 0
    ret = _check_db_bounds(P, "glp_simplex", parm.msg_lev)
    if ret != 0:
        return ret
    if parm.msg_lev >= GLP_MSG_ALL:
        xprintf("GLPK Simplex Optimizer, v4.49")
        xprintf(_size_line(P))
    if P.nnz == 0:
        _trivial_lp(P, parm)
        return 0
    if not parm.presolve:
        return _solve_lp(P, parm, device)
    return _preprocess_and_solve_lp(P, parm, device)


# ---- user callback of the branch-and-cut driver (lib/glpapi13.js:1-7, glpios03.js:533-537) ----
(GLP_IROWGEN, GLP_IBINGO, GLP_IHEUR, GLP_ICUTGEN, GLP_IBRANCH, GLP_ISELECT, GLP_IPREPRO) = range(1, 8)


class _Tree:
    """What a callback can reach through the two ios routines the reference exports
    (``glp_ios_reason``, ``glp_ios_get_prob``)."""
    __slots__ = ("reason", "mip")

    def __init__(self, mip):
        self.reason, self.mip = 0, mip


def glp_ios_reason(tree):
    return tree.reason


def glp_ios_get_prob(tree):
    return tree.mip


def _run_mip_with_callbacks(P, parm, dev, ip):
    """``cb_func != null``: the search runs on the device in slices of ONE node
    (glpb_mip_begin / glpb_mip_run(1) / glpb_mip_end) and the host calls back between
    slices, as SURVEY 8b prescribes.  Emulated reasons: GLP_ISELECT before a node is
    picked and GLP_IBINGO when the incumbent improved (``glp_mip_obj_val`` of
    ``glp_ios_get_prob(tree)`` then gives the new value, which is all the reference's
    test/test.js does).  The other reasons (row/cut generation, heuristics, branching,
    preprocessing requests) belong to features that are off on this path and are not
    raised."""
    rc = dev.mip_begin(ip)
    if rc != 0:
        return rc
    tree = _Tree(P)
    best = None
    while True:
        if dev.mip_open_count() > 0:
            tree.reason = GLP_ISELECT
            parm.cb_func(tree, parm.cb_info)
            tree.reason = 0
        state, _ = dev.mip_run(1)
        has, obj = dev.mip_incumbent()
        if has and (best is None or obj != best):
            best = obj
            P.mip_stat, P.mip_obj = GLP_FEAS, obj
            tree.reason = GLP_IBINGO
            parm.cb_func(tree, parm.cb_info)
            tree.reason = 0
        if state != 1:          # 0 = tree exhausted, GLP_E* = stopped; 1 = slice limit reached
            break
    return dev.mip_end(state)


def _solve_mip(P, parm, device):
    """lib/glpapi09.js:62-114 -- the root LP must be optimal; the tree runs
    behind glpb_intopt on the handle that solved the relaxation."""
    if glp_get_status(P) != GLP_OPT:
        if parm.msg_lev >= GLP_MSG_ERR:
            xprintf("glp_intopt: optimal basis to initial LP relaxation not provided")
        return GLP_EROOT
    if P._dev is None or P._dirty or not P.valid:
        # the optimal basis was found elsewhere (e.g. by a presolve:ON solve, which
        # leaves P.valid = 0): give the handle that basis; the reference's
        # ios_driver re-solves the root from it in the same way (glpios03.js:567)
        quiet = SMCP()
        quiet.msg_lev = GLP_MSG_OFF
        ret = _solve_lp(P, quiet, device)
        if ret != 0 or glp_get_status(P) != GLP_OPT:
            return GLP_EROOT
    dev = P._dev
    ip = dev.iocp(msg_lev=parm.msg_lev, br_tech=parm.br_tech, bt_tech=parm.bt_tech,
                  tol_int=parm.tol_int, tol_obj=parm.tol_obj, tm_lim=int(parm.tm_lim),
                  out_frq=parm.out_frq, out_dly=parm.out_dly, pp_tech=parm.pp_tech,
                  mip_gap=parm.mip_gap, presolve=GLP_OFF, node_lim=getattr(parm, "node_lim", -1))
    ret = dev.intopt(ip) if parm.cb_func is None else _run_mip_with_callbacks(P, parm, dev, ip)
    mp = dev.mip()
    P.mip_stat, P.mip_obj = mp["mip_stat"], mp["mip_obj"]
    for i in range(1, P.m + 1):
        P.row[i].mipx = float(mp["mipx"][i - 1])
    for j in range(1, P.n + 1):
        P.col[j].mipx = float(mp["mipx"][P.m + j - 1])
    _pull(P, dev)
    return ret


def _int_stats_line(P):
    ni, nb = glp_get_num_int(P), glp_get_num_bin(P)
    if nb == 0:
        s = "none of"
    elif ni == 1 and nb == 1:
        s = ""
    elif nb == 1:
        s = "one of"
    elif nb == ni:
        s = "all of"
    else:
        s = "%d of" % nb
    return "%s, %s which %s binary" % (_plural(ni, "integer variable"), s, "is" if nb == 1 else "are")


def _preprocess_and_solve_mip(P, parm, device):
    """lib/glpapi09.js:116-256: MIP presolve in the native library (csrc/presolve.cpp), scaling
    GM|EQ|2N|SKIP and crash basis of the reduced problem, LP relaxation and branch-and-bound on the
    device, recovery of the MIP solution."""
    if parm.msg_lev >= GLP_MSG_ALL:
        xprintf("Preprocessing...")
    npp = _npp_load(P, GLP_MIP)
    try:
        ret = npp.integer(parm.binarize == GLP_ON)
        # what npp_integer / npp_binarize_prob print (glpnpp04.js:92-97, glpnpp05.js:475-514) -- at EVERY msg_lev:
        # the env.term_out switch around the call (glpapi09.js:142-147) is inert, see _quiet
        k = npp.counts()
        if k["bin_vars"] > 0:
            xprintf("%d integer variable(s) were replaced by %d binary ones" % (k["bin_vars"], k["bin_bins"]))
        if k["bin_rows"] > 0:
            xprintf("%d row(s) were added due to binarization" % k["bin_rows"])
        if k["bin_fails"] > 0:
            xprintf("Binarization failed for %d integer variable(s)" % k["bin_fails"])
        if k["packing"] > 0:
            xprintf("%d hidden packing inequaliti(es) were detected" % k["packing"])
        if k["covering"] > 0:
            xprintf("%d hidden covering inequaliti(es) were detected" % k["covering"])
        if k["reduced"] > 0:
            xprintf("%d constraint coefficient(s) were reduced" % k["reduced"])
        if ret != 0:
            if parm.msg_lev >= GLP_MSG_ALL:
                xprintf("PROBLEM HAS NO PRIMAL FEASIBLE SOLUTION" if ret == GLP_ENOPFS
                        else "LP RELAXATION HAS NO DUAL FEASIBLE SOLUTION")
            return ret
        mip = _npp_build(P, npp)
        if mip.m == 0 and mip.n == 0:
            mip.mip_stat = GLP_OPT
            mip.mip_obj = mip.c0
            if parm.msg_lev >= GLP_MSG_ALL:
                xprintf("Objective value = %s" % _num(mip.mip_obj))
                xprintf("INTEGER OPTIMAL SOLUTION FOUND BY MIP PREPROCESSOR")
            ret = 0
        else:
            if parm.msg_lev >= GLP_MSG_ALL:
                xprintf(_size_line(mip))
                xprintf(_int_stats_line(mip))
            try:
                _quiet(parm.msg_lev, glp_scale_prob, mip, GLP_SF_GM | GLP_SF_EQ | GLP_SF_2N | GLP_SF_SKIP)
                _quiet(parm.msg_lev, glp_adv_basis, mip, 0)
                if parm.msg_lev >= GLP_MSG_ALL:
                    xprintf("Solving LP relaxation...")
                smcp = SMCP()
                smcp.msg_lev = parm.msg_lev
                mip.it_cnt = P.it_cnt
                ret = glp_simplex(mip, smcp, device=device)
                P.it_cnt = mip.it_cnt
                if ret != 0:
                    if parm.msg_lev >= GLP_MSG_ERR:
                        xprintf("glp_intopt: cannot solve LP relaxation")
                    return GLP_EFAIL
                ret = glp_get_status(mip)
                if ret == GLP_OPT:
                    ret = 0
                elif ret == GLP_NOFEAS:
                    ret = GLP_ENOPFS
                elif ret == GLP_UNBND:
                    ret = GLP_ENODFS
                if ret != 0:
                    return ret
                mip.it_cnt = P.it_cnt
                ret = _solve_mip(mip, parm, device)
                P.it_cnt = mip.it_cnt
            finally:
                _drop_device(mip)
            if mip.mip_stat not in (GLP_OPT, GLP_FEAS):
                P.mip_stat = mip.mip_stat
                return ret
        _, _, _, c_value = npp.postprocess([mip.col[j].mipx for j in range(1, mip.n + 1)])
        _unload_mip(P, mip.mip_stat, c_value)
        return ret
    finally:
        npp.close()


def _check_iocp(parm):
    # lib/glpapi09.js:265-315
    def bad(name):
        xerror("glp_intopt: %s = %r; invalid parameter" % (name, getattr(parm, name)))
    if parm.msg_lev not in (GLP_MSG_OFF, GLP_MSG_ERR, GLP_MSG_ON, GLP_MSG_ALL, GLP_MSG_DBG): bad("msg_lev")
    if parm.br_tech not in (GLP_BR_FFV, GLP_BR_LFV, GLP_BR_MFV, GLP_BR_DTH, GLP_BR_PCH): bad("br_tech")
    if parm.bt_tech not in (GLP_BT_DFS, GLP_BT_BFS, GLP_BT_BLB, GLP_BT_BPH): bad("bt_tech")
    if not (0.0 < parm.tol_int < 1.0): bad("tol_int")
    if not (0.0 < parm.tol_obj < 1.0): bad("tol_obj")
    if parm.tm_lim < 0: bad("tm_lim")
    if parm.out_frq < 0: bad("out_frq")
    if parm.out_dly < 0: bad("out_dly")
    if not (0 <= parm.cb_size <= 256): bad("cb_size")
    if parm.pp_tech not in (GLP_PP_NONE, GLP_PP_ROOT, GLP_PP_ALL): bad("pp_tech")
    if parm.mip_gap < 0.0: bad("mip_gap")
    for name in ("mir_cuts", "gmi_cuts", "cov_cuts", "clq_cuts", "presolve", "binarize", "fp_heur"):
        if getattr(parm, name) not in (GLP_ON, GLP_OFF): bad(name)


def glp_intopt(P, parm=None, device=0):
    """lib/glpapi09.js:258-390"""
    _check(P, "glp_intopt")
    if parm is None:
        parm = IOCP()
    _check_iocp(parm)
    # options whose machinery is outside this path (SURVEY 8 "out of scope": cut generators lib/glpios05-08,
    # feasibility pump lib/glpios10, pseudocost branching lib/glpios09.js:365-660): valid in the reference,
    # refused here instead of being silently ignored
    for name in ("mir_cuts", "gmi_cuts", "cov_cuts", "clq_cuts", "fp_heur"):
        if getattr(parm, name) == GLP_ON:
            xerror("glp_intopt: %s = GLP_ON; not supported by the B200 path" % name)
    if parm.br_tech == GLP_BR_PCH:
        xerror("glp_intopt: br_tech = GLP_BR_PCH; not supported by the B200 path")
    P.mip_stat, P.mip_obj = GLP_UNDEF, 0.0
    ret = _check_db_bounds(P, "glp_intopt", parm.msg_lev)
    if ret != 0:
        return ret
    for j in range(1, P.n + 1):  # lib/glpapi09.js:337-364
        c = P.col[j]
        if c.kind != GLP_IV:
            continue
        what = None
        if c.type in (GLP_LO, GLP_DB) and c.lb != math.floor(c.lb):
            what = "lower bound %s" % _num(c.lb)
        elif c.type in (GLP_UP, GLP_DB) and c.ub != math.floor(c.ub):
            what = "upper bound %s" % _num(c.ub)
        elif c.type == GLP_FX and c.lb != math.floor(c.lb):
            what = "fixed value %s" % _num(c.lb)
        if what:
            if parm.msg_lev >= GLP_MSG_ERR:
                xprintf("glp_intopt: integer column %d has non-integer %s" % (j, what))
            return GLP_EBOUND
    if parm.msg_lev >= GLP_MSG_ALL:
        xprintf("GLPK Integer Optimizer, v4.49")
        xprintf(_size_line(P))
        xprintf(_int_stats_line(P))
    if not parm.presolve:
        return _solve_mip(P, parm, device)
    return _preprocess_and_solve_mip(P, parm, device)
