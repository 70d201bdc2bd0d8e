"""LP / MIP presolver (glpb_npp_*, csrc/presolve.cpp) against the REFERENCE'S OWN presolver.

tests/golden/ref_npp.json was produced by oracle/jsref/make_npp_golden.py: the unmodified
lib/glpnpp01-05.js run on the reference's fixtures and on 90 generated problems.  Everything here is
host code (no device): bit-exact comparison of
  * the return code (0 / GLP_ENOPFS / GLP_ENODFS) and the depth of the recovery stack,
  * the reduced problem npp_build_prob leaves -- row order, column order, bounds, costs, the constant
    term and the element order inside every column (it decides scaling, crash basis and pivots),
  * npp_postprocess: fed the solution the reference found for the reduced problem, the recovered
    statuses / multipliers / values must be the reference's,
  * npp_unload_sol through the facade: the solution stored into the original problem object.
"""
import json
import os

import numpy as np
import pytest

import glpk_js_b200 as G
from glpk_js_b200 import glpk as F, native

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "ref_npp.json")) as f:
    CASES = {k: v for k, v in json.load(f).items() if not k.startswith("_")}


def facade_problem(d):
    """the problem object in the reference's list state: columns exactly in the recorded list order"""
    P = F.glp_create_prob()
    F.glp_set_obj_dir(P, d["dir"])
    P.c0 = d["c0"]
    if d["m"]:
        F.glp_add_rows(P, d["m"])
    if d["n"]:
        F.glp_add_cols(P, d["n"])
    for i in range(d["m"]):
        r = P.row[i + 1]
        r.type, r.lb, r.ub = d["r_type"][i], d["r_lb"][i], d["r_ub"][i]
        r.stat = F.GLP_BS
    for j in range(d["n"]):
        c = P.col[j + 1]
        c.type, c.lb, c.ub, c.coef, c.kind = d["c_type"][j], d["c_lb"][j], d["c_ub"][j], d["c_coef"][j], d["c_kind"][j]
        for e in range(d["A_ptr"][j], d["A_ptr"][j + 1]):
            c.elems.append((d["A_ind"][e] + 1, d["A_val"][e]))
    # row lists as glp_sort_matrix leaves them (ascending columns, lib/glpapi01.js:620-649): the presolver
    # never reads them, the recomputation of row activities in npp_unload_sol does
    for j in range(1, d["n"] + 1):
        for (i, v) in P.col[j].elems:
            P.row[i].elems.append((j, v))
    P.nnz = len(d["A_val"])
    return P


def same_problem(Q, red):
    assert (Q.m, Q.n) == (red["m"], red["n"])
    assert Q.c0 == red["c0"] and Q.dir == red["dir"]
    for i in range(Q.m):
        r = Q.row[i + 1]
        assert (r.type, r.lb, r.ub) == (red["r_type"][i], red["r_lb"][i], red["r_ub"][i]), ("row", i + 1)
    ptr, ind, val = F._csc(Q)
    assert list(ptr) == red["A_ptr"]
    assert list(ind) == red["A_ind"]
    assert list(val) == red["A_val"]
    for j in range(Q.n):
        c = Q.col[j + 1]
        assert (c.type, c.lb, c.ub, c.coef, c.kind) == (
            red["c_type"][j], red["c_lb"][j], red["c_ub"][j], red["c_coef"][j], red["c_kind"][j]), ("col", j + 1)


@pytest.mark.parametrize("name", sorted(CASES))
def test_presolver_matches_reference(name):
    case = CASES[name]
    P = facade_problem(case["problem"])
    npp = F._npp_load(P, case["sol"])
    try:
        ret = npp.simplex() if case["sol"] == F.GLP_SOL else npp.integer(bool(case["binarize"]))
        assert ret == case["ret"]
        assert npp.counts()["stack"] == case["n_tse"]
        if ret != 0:
            return
        Q = F._npp_build(P, npp)
        same_problem(Q, case["reduced"])
        assert list(Q._npp_ref[0]) == case["reduced"]["row_ref"]
        assert list(Q._npp_ref[1]) == case["reduced"]["col_ref"]
        if "post" not in case:
            return
        gin, post, un = case["in"], case["post"], case["unloaded"]
        sgn = 1.0 if case["problem"]["dir"] == F.GLP_MIN else -1.0
        if case["sol"] == F.GLP_SOL:
            r_stat, r_dual, c_stat, c_value = npp.postprocess(gin["c_value"], gin["r_stat"], gin["r_dual"],
                                                              gin["c_stat"])
            assert list(r_stat) == post["r_stat"]
            assert list(c_stat) == post["c_stat"]
            assert list(c_value) == post["c_value"]
            assert list(r_dual) == [sgn * v for v in post["r_pi"]]
            F._unload_basic(P, un["prim_stat"], un["dual_stat"], r_stat, r_dual, c_stat, c_value)
            assert P.obj_val == un["obj"]
            assert [P.row[i].stat for i in range(1, P.m + 1)] == un["row_stat"]
            assert [P.col[j].stat for j in range(1, P.n + 1)] == un["col_stat"]
            assert [P.row[i].prim for i in range(1, P.m + 1)] == un["row_prim"]
            assert [P.row[i].dual for i in range(1, P.m + 1)] == un["row_dual"]
            assert [P.col[j].prim for j in range(1, P.n + 1)] == un["col_prim"]
            assert [P.col[j].dual for j in range(1, P.n + 1)] == un["col_dual"]
        else:
            _, _, _, c_value = npp.postprocess(gin["c_value"])
            assert list(c_value) == post["c_value"]
            F._unload_mip(P, un["mip_stat"], c_value)
            assert P.mip_obj == un["mip_obj"]
            assert [P.col[j].mipx for j in range(1, P.n + 1)] == un["col_val"]
            assert [P.row[i].mipx for i in range(1, P.m + 1)] == un["row_val"]
    finally:
        npp.close()


def test_golden_covers_the_transformations():
    """the generated cases are only worth something if they reach the presolver's branches"""
    rets = [c["ret"] for c in CASES.values()]
    assert rets.count(0) >= 45 and rets.count(F.GLP_ENOPFS) >= 5 and rets.count(F.GLP_ENODFS) >= 5
    assert sum(1 for c in CASES.values() if c["ret"] == 0 and c["reduced"]["m"] == 0) >= 2      # solved by the presolver
    assert sum(1 for c in CASES.values() if c["ret"] == 0 and c["reduced"]["m"] >= 4) >= 20
    grown = [c for c in CASES.values() if c["ret"] == 0 and c["sol"] == F.GLP_MIP and
             (c["reduced"]["n"] > c["problem"]["n"] or max(c["reduced"]["row_ref"] + [0]) > c["problem"]["m"])]
    assert len(grown) >= 3          # binarization / row copies added rows or columns
    total = {}
    for case in CASES.values():
        npp = F._npp_load(facade_problem(case["problem"]), case["sol"])
        npp.simplex() if case["sol"] == F.GLP_SOL else npp.integer(bool(case["binarize"]))
        for k, v in npp.counts().items():
            total[k] = total.get(k, 0) + v
        npp.close()
    for kind in native.Presolver.KINDS + ("packing", "covering", "reduced", "bin_vars", "bin_rows", "bin_fails"):
        assert total[kind] > 0, kind    # every transformation on the path is reached by some case


def test_workspace_argument_checks():
    L = native.load()
    assert L.glpb_npp_simplex(None) == native.GLPB_EINVAL
    assert L.glpb_npp_postprocess(None, *([None] * 8)) == native.GLPB_EINVAL
    L.glpb_npp_destroy(None)
    d = dict(m=1, n=1, dir=1, c0=0.0, type=[F.GLP_UP, F.GLP_LO], lb=[0.0, 0.0], ub=[4.0, 0.0], coef=[-1.0],
             kind=[1], A_ptr=[0, 1], A_ind=[0], A_val=[2.0])
    with pytest.raises(ValueError):
        native.Presolver(dict(d, A_ind=[3]), F.GLP_SOL)         # row index out of range
    with pytest.raises(ValueError):
        native.Presolver(dict(d, type=[9, 2]), F.GLP_SOL)       # invalid type
    npp = native.Presolver(d, F.GLP_SOL)
    assert L.glpb_npp_integer(npp.h, 0) == native.GLPB_EINVAL   # loaded for a basic solution
    assert L.glpb_npp_postprocess(npp.h, *([None] * 8)) == native.GLPB_EINVAL   # not built yet
    assert npp.simplex() == 0
    red = npp.build()
    # max x s.t. 2x <= 4, x >= 0: the singleton row becomes a column bound, the column is then empty and fixed
    assert (red["m"], red["n"]) == (0, 0) and red["c0"] == -2.0
    r_stat, r_dual, c_stat, c_value = npp.postprocess([], [], [], [])
    assert list(c_value) == [2.0] and list(c_stat) == [F.GLP_BS] and list(r_stat) == [F.GLP_NU]
    assert list(r_dual) == [-0.5]
    npp.close()


@pytest.mark.parametrize("opt", ["mir_cuts", "gmi_cuts", "cov_cuts", "clq_cuts", "fp_heur", "br_tech"])
def test_intopt_refuses_options_outside_the_path(opt):
    """valid in the reference, no machinery behind them here: refused loudly, never silently ignored"""
    P = facade_problem(CASES["npp_mip_2"]["problem"])
    parm = F.IOCP({opt: F.GLP_BR_PCH if opt == "br_tech" else F.GLP_ON})
    with pytest.raises(F.GlpkError, match="not supported by the B200 path"):
        F.glp_intopt(P, parm)


HOST_ONLY = sorted(k for k, c in CASES.items() if "terminal" in c)


@pytest.mark.parametrize("name", HOST_ONLY)
def test_whole_call_on_the_host_when_the_presolver_decides(name):
    """Problems the presolver rejects (GLP_ENOPFS / GLP_ENODFS) or solves outright (empty reduced problem) never
    reach the device: glp_simplex / glp_intopt(presolve: GLP_ON) of the facade, whole call, against the
    reference's whole call -- return code, every line of terminal output at msg_lev OFF and ALL, and the
    solution stored in the problem object."""
    case = CASES[name]
    for lev in (F.GLP_MSG_OFF, F.GLP_MSG_ALL):
        ref = case["terminal"][str(lev)]
        P = facade_problem(case["problem"])
        lines = []
        F.glp_set_print_func(lines.append)
        try:
            if case["sol"] == F.GLP_SOL:
                parm = F.SMCP({"presolve": F.GLP_ON})
                parm.msg_lev = lev
                ret = F.glp_simplex(P, parm)
            else:
                parm = F.IOCP({"presolve": F.GLP_ON, "binarize": F.GLP_ON if case["binarize"] else F.GLP_OFF})
                parm.msg_lev = lev
                ret = F.glp_intopt(P, parm)
        finally:
            F.glp_set_print_func(None)
        assert ret == ref["ret"] == case["ret"]
        assert lines == ref["lines"]
        assert P._dev is None                       # nothing was uploaded
        if ret != 0:
            continue
        un = case["unloaded"]
        if case["sol"] == F.GLP_SOL:
            assert (F.glp_get_status(P), P.obj_val) == (un["status"], un["obj"])
            assert [P.col[j].prim for j in range(1, P.n + 1)] == un["col_prim"]
            assert [P.col[j].stat for j in range(1, P.n + 1)] == un["col_stat"]
            assert [P.row[i].dual for i in range(1, P.m + 1)] == un["row_dual"]
        else:
            assert (F.glp_mip_status(P), P.mip_obj) == (un["mip_stat"], un["mip_obj"])
            assert [P.col[j].mipx for j in range(1, P.n + 1)] == un["col_val"]


class _DeviceReached(Exception):
    pass


NEEDS_DEVICE = sorted(k for k, c in CASES.items() if "prep" in c)       # generated problems and the fixtures


@pytest.mark.parametrize("name", NEEDS_DEVICE)
def test_messages_up_to_the_device_solve(name, monkeypatch):
    """At GLP_MSG_OFF the solvers are silent, so everything the reference printed during the whole
    presolve-ON call came from npp_integer, glp_scale_prob and glp_adv_basis ON THE REDUCED PROBLEM (its
    term_out switch is inert): scale ranges after every pass, size of the triangular part.  The facade must
    have printed exactly those lines by the time it hands the reduced problem to the device -- a check of the
    reduced problem, the native scaling and the native crash basis in one."""
    case = CASES[name]

    handed = []

    def stop(lp, *a, **k):
        handed.append(lp)
        raise _DeviceReached()
    monkeypatch.setattr(F, "_solve_lp", stop)
    P = facade_problem(case["problem"])
    lines = []
    F.glp_set_print_func(lines.append)
    try:
        with pytest.raises(_DeviceReached):
            if case["sol"] == F.GLP_SOL:
                parm = F.SMCP({"presolve": F.GLP_ON})
                parm.msg_lev = F.GLP_MSG_OFF
                F.glp_simplex(P, parm)
            else:
                parm = F.IOCP({"presolve": F.GLP_ON, "binarize": F.GLP_ON if case["binarize"] else F.GLP_OFF})
                parm.msg_lev = F.GLP_MSG_OFF
                F.glp_intopt(P, parm)
    finally:
        F.glp_set_print_func(None)
    if "terminal_off" in case:
        assert lines == case["terminal_off"]["lines"]
    # ... and the problem handed over carries the reference's scale factors and crash basis, bit for bit
    lp, prep = handed[0], case["prep"]
    assert [lp.row[i].rii for i in range(1, lp.m + 1)] == prep["rii"]
    assert [lp.col[j].sjj for j in range(1, lp.n + 1)] == prep["sjj"]
    assert [lp.row[i].stat for i in range(1, lp.m + 1)] == prep["row_stat"]
    assert [lp.col[j].stat for j in range(1, lp.n + 1)] == prep["col_stat"]
