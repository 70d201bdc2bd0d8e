"""Regenerates tests/golden/*.json from the reference's own fixtures.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):   python tests/golden/make_golden.py

For each of the reference's LP files (test/test.lpt, test/gap.lpt,
test/todd.lpt) it stores
  * the problem as plain arrays (0-based CSC, columns ascending in row index =
    the state glp_read_lp leaves after glp_sort_matrix), names included, so the
    tests can rebuild it with glp_add_rows/glp_set_mat_col or glpb_create;
  * independent optima computed with scipy/HiGHS (LP relaxation and MIP) --
    the reference asserts no results of its own ("parity unpinned" there);
  * the hand trace of test.lpt through spx_primal.
The LP text itself is not copied: it is re-emitted by our own writer so that
the LP reader still has files to parse (tests/golden/*.lp).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

REF = "/root/reference/test"
INF = float("inf")


def highs(d, integer):
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
               integrality=integrality, options=dict(mip_rel_gap=0.0))
    assert res.status == 0, res
    return float(sgn * res.fun + d["c0"]), [float(v) for v in res.x]


def main():
    for name in ("test", "gap", "todd"):
        text = open(os.path.join(REF, name + ".lpt")).read()
        P = O.Problem.from_lp(text)
        d = P.export()
        lp_obj, lp_x = highs(d, False)
        out = {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in d.items()}
        out["source"] = "reference test/%s.lpt, parsed by oracle/prob.cpp read_lp" % name
        out["highs_lp_obj"] = lp_obj
        if (d["c_kind"] == O.GLP_IV).any():
            out["highs_mip_obj"] = highs(d, True)[0]
        if name == "test":
            out["highs_lp_x"] = lp_x
            out["hand_trace_primal"] = {"pivots": [[1, 2], [2, 1]], "it_cnt": 2,
                                        "cbar_after_1": [-100.0, -200.0, 100.0],
                                        "gamma_after_1": [0.01, 1.16, 1.25],
                                        "bbar_after_1": [60.0, 60.0, 120.0], "head_final": [5, 4, 3]}
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(out, f)
        with open(os.path.join(HERE, name + ".lp"), "w") as f:
            f.write(P.write_lp())
        print(name, d["m"], d["n"], len(d["A_val"]), "LP", lp_obj, "MIP", out.get("highs_mip_obj"))


if __name__ == "__main__":
    main()
