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
