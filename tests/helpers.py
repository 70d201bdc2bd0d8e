"""Shared test helpers: golden fixtures, layout conversion, KKT residuals."""
import json
import os

import numpy as np

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        d = json.load(f)
    for k in ("r_type", "c_type", "c_kind", "A_ptr", "A_ind"):
        d[k] = np.array(d[k], dtype=np.int32)
    for k in ("r_lb", "r_ub", "c_lb", "c_ub", "c_coef", "A_val", "rii", "sjj"):
        d[k] = np.array(d[k], dtype=np.float64)
    return d


def golden_text(name):
    with open(os.path.join(GOLDEN, name + ".lp")) as f:
        return f.read()


def to_native(d):
    """oracle layout (rows/cols separate) -> glpb_create layout (type/lb/ub [m+n])"""
    return dict(m=d["m"], n=d["n"], dir=d["dir"], c0=d["c0"],
                type=np.concatenate([d["r_type"], d["c_type"]]).astype(np.int32),
                lb=np.concatenate([d["r_lb"], d["c_lb"]]), ub=np.concatenate([d["r_ub"], d["c_ub"]]),
                coef=d["c_coef"], kind=d["c_kind"], A_ptr=d["A_ptr"], A_ind=d["A_ind"], A_val=d["A_val"])


def to_oracle(d):
    """glpb_create layout -> oracle layout"""
    m = d["m"]
    return dict(m=m, n=d["n"], dir=d["dir"], c0=d["c0"], r_type=d["type"][:m], r_lb=d["lb"][:m],
                r_ub=d["ub"][:m], c_type=d["type"][m:], c_lb=d["lb"][m:], c_ub=d["ub"][m:],
                c_coef=d["coef"], c_kind=d["kind"], A_ptr=d["A_ptr"], A_ind=d["A_ind"], A_val=d["A_val"])


def dense_A(d):
    m, n = d["m"], d["n"]
    A = np.zeros((m, n))
    for j in range(n):
        for t in range(d["A_ptr"][j], d["A_ptr"][j + 1]):
            A[d["A_ind"][t], j] = d["A_val"][t]
    return A


def spmv(d, x):
    """A @ x from CSC without densifying"""
    y = np.zeros(d["m"])
    counts = np.diff(d["A_ptr"])
    cols = np.repeat(np.arange(d["n"]), counts)
    np.add.at(y, d["A_ind"], d["A_val"] * x[cols])
    return y


def spmv_t(d, y):
    counts = np.diff(d["A_ptr"])
    cols = np.repeat(np.arange(d["n"]), counts)
    out = np.zeros(d["n"])
    np.add.at(out, cols, d["A_val"] * y[d["A_ind"]])
    return out


def kkt(d, sol):
    """Residuals in the sense of glp_check_kkt (lib/glpapi10.js:15-58), native layout.
    Returns dict of max relative errors: PE (row activity), PB (bounds),
    DE (reduced-cost equation), DB (dual sign)."""
    m, n = d["m"], d["n"]
    x = sol["prim"][m:]
    r = sol["prim"][:m]
    ax = spmv(d, x)
    pe = np.max(np.abs(r - ax) / (1.0 + np.abs(r))) if m else 0.0
    t, lb, ub = d["type"], d["lb"], d["ub"]
    v = sol["prim"]
    has_lb = np.isin(t, (O.GLP_LO, O.GLP_DB, O.GLP_FX))
    has_ub = np.isin(t, (O.GLP_UP, O.GLP_DB, O.GLP_FX))
    viol = np.zeros(m + n)
    viol = np.where(has_lb & (v < lb), (lb - v) / (1.0 + np.abs(lb)), viol)
    viol = np.where(has_ub & (v > ub), np.maximum(viol, (v - ub) / (1.0 + np.abs(ub))), viol)
    pb = float(viol.max())
    # dual: c_j - A_j' pi = d_j with pi = row duals
    pi = sol["dual"][:m]
    dj = sol["dual"][m:]
    de = np.max(np.abs(d["coef"] - spmv_t(d, pi) - dj) / (1.0 + np.abs(d["coef"])))
    sgn = 1.0 if d["dir"] == O.GLP_MIN else -1.0
    dd = sgn * sol["dual"]
    st = sol["stat"]
    bad = np.zeros(m + n)
    # at optimum: basic -> 0; NL -> >= 0; NU -> <= 0; NF -> 0 (in min form)
    bad = np.where(st == O.GLP_BS, np.abs(dd), bad)
    bad = np.where(st == O.GLP_NL, np.maximum(0.0, -dd), bad)
    bad = np.where(st == O.GLP_NU, np.maximum(0.0, dd), bad)
    bad = np.where(st == O.GLP_NF, np.abs(dd), bad)
    db = float(bad.max())
    return dict(PE=float(pe), PB=pb, DE=float(de), DB=db)


def basis_matrix(d, head):
    """Dense B from the basis header (values k = 1..m+n) of (I | -A), unscaled."""
    m = d["m"]
    A = dense_A(d)
    B = np.zeros((m, m))
    for i, k in enumerate(head):
        if k <= m:
            B[k - 1, i] = 1.0
        else:
            B[:, i] = -A[:, k - m - 1]
    return B


# ---- randomised general LPs / small MIPs (shared by the GPU parity tests and the HiGHS pins) ----
def random_lp(seed):
    rng = np.random.default_rng(seed)
    m, n = int(rng.integers(4, 36)), int(rng.integers(4, 48))
    dens = rng.uniform(0.15, 0.6)
    A = np.where(rng.random((m, n)) < dens, np.round(rng.uniform(-5, 5, (m, n)), 2), 0.0)
    for j in range(n):                      # no empty columns
        if not A[:, j].any():
            A[rng.integers(0, m), j] = float(rng.integers(1, 5))
    x0 = np.round(rng.uniform(-2, 4, n), 1)  # a point that most rows are built around
    ax = A @ x0
    rt = rng.choice([O.GLP_FR, O.GLP_LO, O.GLP_UP, O.GLP_DB, O.GLP_FX], size=m, p=[0.05, 0.3, 0.3, 0.2, 0.15])
    slack = np.round(rng.uniform(0, 3, m), 1)
    shift = np.where(rng.random(m) < 0.1, rng.uniform(-6, 6, m), 0.0)     # now and then infeasible
    rl = np.where(np.isin(rt, (O.GLP_LO, O.GLP_DB, O.GLP_FX)), ax - slack + shift, 0.0)
    ru = np.where(rt == O.GLP_UP, ax + slack + shift, np.where(rt == O.GLP_DB, rl + 2 * slack + 0.5, np.where(rt == O.GLP_FX, rl, 0.0)))
    ct = rng.choice([O.GLP_FR, O.GLP_LO, O.GLP_UP, O.GLP_DB, O.GLP_FX], size=n, p=[0.1, 0.45, 0.1, 0.3, 0.05])
    cl = np.where(np.isin(ct, (O.GLP_LO, O.GLP_DB, O.GLP_FX)), np.round(x0 - rng.uniform(0, 3, n), 1), 0.0)
    cu = np.where(ct == O.GLP_UP, np.round(x0 + rng.uniform(0, 3, n), 1),
                  np.where(ct == O.GLP_DB, cl + np.round(rng.uniform(0.5, 6, n), 1), np.where(ct == O.GLP_FX, cl, 0.0)))
    coef = np.round(rng.uniform(-4, 4, n), 1)
    ptr, ind, val = [0], [], []
    for j in range(n):
        nz = np.nonzero(A[:, j])[0]
        ind.extend(nz.tolist())
        val.extend(A[nz, j].tolist())
        ptr.append(len(ind))
    return dict(m=m, n=n, dir=int(rng.choice([O.GLP_MIN, O.GLP_MAX])), c0=float(np.round(rng.uniform(-3, 3), 1)),
                r_type=rt.astype(np.int32), r_lb=rl, r_ub=ru, c_type=ct.astype(np.int32), c_lb=cl, c_ub=cu,
                c_coef=coef, c_kind=np.full(n, O.GLP_CV, np.int32), A_ptr=np.array(ptr, np.int32),
                A_ind=np.array(ind, np.int32), A_val=np.array(val, np.float64))


def random_mip(seed):
    """small bounded integer programs on top of random_lp: every column boxed, about
    two thirds of them integer, integral bounds"""
    rng = np.random.default_rng(5000 + seed)
    d = random_lp(7000 + seed)
    n = d["n"]
    d["c_type"] = np.full(n, O.GLP_DB, np.int32)
    lo = np.floor(rng.uniform(-3, 2, n))
    d["c_lb"] = lo
    d["c_ub"] = lo + rng.integers(1, 6, n).astype(float)
    d["c_kind"] = np.where(rng.random(n) < 0.65, O.GLP_IV, O.GLP_CV).astype(np.int32)
    # rows rebuilt around an integer point inside the boxes, so that the program is feasible
    x0 = np.floor(rng.uniform(d["c_lb"], d["c_ub"] + 1.0 - 1e-9))
    A = dense_A(dict(m=d["m"], n=n, A_ptr=d["A_ptr"], A_ind=d["A_ind"], A_val=d["A_val"]))
    ax = A @ x0
    slack = np.round(rng.uniform(0.5, 4, d["m"]), 1)
    rt = d["r_type"]
    d["r_lb"] = np.where(np.isin(rt, (O.GLP_LO, O.GLP_DB)), ax - slack, np.where(rt == O.GLP_FX, ax, 0.0))
    d["r_ub"] = np.where(rt == O.GLP_UP, ax + slack, np.where(rt == O.GLP_DB, ax + slack, np.where(rt == O.GLP_FX, ax, 0.0)))
    return d


def transport_lp(seed, ns=6, nd=7):
    """A small transportation LP with unit coefficients and small integer data: totally unimodular, so every
    basis inverse, every ratio and every reduced cost is an integer computed exactly in floating point --
    whatever the order of the arithmetic.  Its ratio tests tie all the time, and the ties are EXACT on every
    implementation, which makes the pivot sequence a test of tie order alone (sort_tcol / sort_trow)."""
    rng = np.random.default_rng(seed)
    supply = rng.integers(3, 9, ns)
    demand = rng.integers(1, 6, nd)
    demand[-1] += max(0, int(supply.sum() - demand.sum()) // 2)
    m, n = ns + nd, ns * nd
    ptr, ind, val = [0], [], []
    for s in range(ns):
        for t in range(nd):
            ind += [s, ns + t]
            val += [1.0, 1.0]
            ptr.append(len(ind))
    rt = np.array([O.GLP_UP] * ns + [O.GLP_LO] * nd, np.int32)
    rl = np.concatenate([np.zeros(ns), np.minimum(demand, supply.sum() // nd).astype(float)])
    ru = np.concatenate([supply.astype(float), np.zeros(nd)])
    ct = np.where(rng.random(n) < 0.3, O.GLP_DB, O.GLP_LO).astype(np.int32)
    cu = np.where(ct == O.GLP_DB, rng.integers(1, 4, n), 0).astype(float)
    return dict(m=m, n=n, dir=O.GLP_MIN, c0=0.0, r_type=rt, r_lb=rl, r_ub=ru, c_type=ct, c_lb=np.zeros(n), c_ub=cu,
                c_coef=rng.integers(1, 5, n).astype(float), c_kind=np.full(n, O.GLP_CV, np.int32),
                A_ptr=np.array(ptr, np.int32), A_ind=np.array(ind, np.int32), A_val=np.array(val, np.float64))
