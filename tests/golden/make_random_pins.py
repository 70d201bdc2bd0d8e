"""Independent pins for the randomised LPs / MIPs of tests/helpers.py
(random_lp, random_mip): status and optimum computed with scipy's HiGHS, stored
in tests/golden/random_pins.json.  The reference ships no asserted results
("parity unpinned" there), so these -- like the fixture optima of
make_golden.py -- are what pins the oracle, and through the oracle the device.

    python tests/golden/make_random_pins.py          (build container; needs scipy)

Each entry carries a checksum of the generated coefficients so that a drift of
the generator (another numpy) is reported as such, not as a parity failure.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import helpers as H  # noqa: E402
import oracle_lib as O  # noqa: E402

INF = float("inf")


def checksum(d):
    return float(np.sum(d["A_val"] * (1.0 + np.arange(len(d["A_val"])) % 7)) + np.sum(d["c_coef"]) + np.sum(d["r_lb"])
                 + np.sum(d["r_ub"]) + np.sum(d["c_lb"]) + np.sum(d["c_ub"]))


def highs(d, integer, presolve=False):
    from scipy.optimize import milp, LinearConstraint, Bounds
    from scipy.sparse import csc_matrix
    m, n = d["m"], d["n"]
    A = csc_matrix((d["A_val"], d["A_ind"], d["A_ptr"]), shape=(m, n))
    sgn = 1.0 if d["dir"] == O.GLP_MIN else -1.0

    def bnd(t, lb, ub):
        lo = np.where(np.isin(t, (O.GLP_LO, O.GLP_DB, O.GLP_FX)), lb, -INF)
        hi = np.where(np.isin(t, (O.GLP_UP, O.GLP_DB)), ub, np.where(t == O.GLP_FX, lb, INF))
        return lo, hi
    rl, ru = bnd(d["r_type"], d["r_lb"], d["r_ub"])
    cl, cu = bnd(d["c_type"], d["c_lb"], d["c_ub"])
    integrality = (d["c_kind"] == O.GLP_IV).astype(int) if integer else np.zeros(n, int)
    res = milp(sgn * d["c_coef"], constraints=LinearConstraint(A, rl, ru), bounds=Bounds(cl, cu),
               integrality=integrality, options=dict(mip_rel_gap=0.0, presolve=presolve, time_limit=120.0))
    # scipy: 0 optimal, 2 infeasible, 3 unbounded, 4 other (HiGHS "unbounded or infeasible")
    kind = {0: "optimal", 2: "infeasible", 3: "unbounded"}.get(res.status, "other")
    obj = float(sgn * res.fun + d["c0"]) if res.status == 0 else None
    return kind, obj


def main():
    out = {"lp": [], "mip": []}
    for seed in range(48):
        d = H.random_lp(1000 + seed)
        kind, obj = highs(d, False)
        out["lp"].append(dict(seed=1000 + seed, m=d["m"], n=d["n"], checksum=checksum(d), highs=kind, obj=obj))
    for seed in range(24):
        d = H.random_mip(seed)
        kind, obj = highs(d, True)
        lkind, lobj = highs(d, False)
        out["mip"].append(dict(seed=seed, m=d["m"], n=d["n"], checksum=checksum(d), highs=kind, obj=obj,
                               lp_highs=lkind, lp_obj=lobj))
    # the knapsack instances of the branch-and-bound tests (generator = glpb_gen_mkp, C5 family)
    import glpk_js_b200 as G
    out["mkp"] = []
    for (m, n, seed) in ((5, 30, 20240701), (5, 30, 20240702), (8, 40, 20240702)):
        d = H.to_oracle(G.native.generate("mkp", m=m, n=n, seed=seed))
        kind, obj = highs(d, True, presolve=True)
        lkind, lobj = highs(d, False)
        out["mkp"].append(dict(m=m, n=n, seed=seed, checksum=checksum(d), highs=kind, obj=obj, lp_obj=lobj))
    with open(os.path.join(HERE, "random_pins.json"), "w") as f:
        json.dump(out, f, indent=1)
    print({k: {s: sum(1 for e in v if e["highs"] == s) for s in ("optimal", "infeasible", "unbounded", "other")}
           for k, v in out.items()})


if __name__ == "__main__":
    main()
