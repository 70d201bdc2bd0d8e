"""presolve: GLP_ON end to end on the device against the REFERENCE'S OWN presolve-ON runs.

glp_simplex / glp_intopt of the facade with the native presolver (glpb_npp_*), scaling, crash basis, the
device solve of the REDUCED problem and the recovery -- compared with what the unmodified reference did on
the same problems (tests/golden/ref_runs.json "presolve_1", tests/golden/ref_npp.json): same return code,
the SAME NUMBER OF SIMPLEX ITERATIONS on the reduced LP, identical status vectors, values to 1e-9."""
import json
import os

import pytest

import glpk_js_b200 as G
import helpers as H
import test_presolve as T

glpk = G.glpk
pytestmark = pytest.mark.gpu
with open(os.path.join(H.GOLDEN, "ref_runs.json")) as f:
    REF = json.load(f)
TOL = 1e-9


def close(a, b):
    return abs(a - b) <= TOL * max(1.0, abs(b))


def assert_basic_solution(P, ref):
    assert [P.row[i].stat for i in range(1, P.m + 1)] == ref["row_stat"]
    assert [P.col[j].stat for j in range(1, P.n + 1)] == ref["col_stat"]
    assert close(P.obj_val, ref["obj"])
    for j in range(1, P.n + 1):
        assert close(P.col[j].prim, ref["col_prim"][j - 1]) and close(P.col[j].dual, ref["col_dual"][j - 1]), j
    for i in range(1, P.m + 1):
        assert close(P.row[i].prim, ref["row_prim"][i - 1]) and close(P.row[i].dual, ref["row_dual"][i - 1]), i


@pytest.mark.parametrize("name", ["test", "gap", "todd"])
def test_fixtures_presolve_on_like_the_reference(name):
    """test/test.js with presolve ON: 2 / 49 / 8 iterations in the reference (57 on gap.lpt without the
    presolver), MIP optimum of the reduced problem recovered into the original columns"""
    ref = REF[name]["presolve_1"]
    lp = glpk.glp_create_prob()
    assert glpk.glp_read_lp_from_string(lp, None, H.golden_text(name)) == 0
    p = glpk.SMCP({"presolve": glpk.GLP_ON})
    p.msg_lev = glpk.GLP_MSG_OFF
    assert glpk.glp_simplex(lp, p) == ref["lp"]["ret"] == 0
    assert lp.it_cnt == ref["lp"]["it_cnt"]
    assert (glpk.glp_get_status(lp), lp.pbs_stat, lp.dbs_stat) == (ref["lp"]["status"], ref["lp"]["prim_stat"],
                                                                   ref["lp"]["dual_stat"])
    assert_basic_solution(lp, ref["lp"])
    io = glpk.IOCP({"presolve": glpk.GLP_ON})
    io.msg_lev = glpk.GLP_MSG_OFF
    assert glpk.glp_intopt(lp, io) == ref["mip"]["ret"] == 0
    assert glpk.glp_mip_status(lp) == ref["mip"]["mip_stat"]
    assert close(glpk.glp_mip_obj_val(lp), ref["mip"]["mip_obj"])
    if name == "todd":      # unique optimum
        assert [lp.col[j].mipx for j in range(1, lp.n + 1)] == ref["mip"]["col_val"]


LP_CASES = sorted(k for k, c in T.CASES.items() if k.startswith("npp_lp_"))
MIP_CASES = sorted(k for k, c in T.CASES.items() if k.startswith("npp_mip_"))


@pytest.mark.parametrize("name", LP_CASES)
def test_generated_lps_presolve_on(name):
    case = T.CASES[name]
    P = T.facade_problem(case["problem"])
    p = glpk.SMCP({"presolve": glpk.GLP_ON})
    p.msg_lev = glpk.GLP_MSG_OFF
    ret = glpk.glp_simplex(P, p)
    if case["ret"] != 0:
        assert ret == case["ret"] and glpk.glp_get_status(P) == glpk.GLP_UNDEF
        return
    if "post" not in case:      # the reference could not solve the reduced LP to optimality either
        assert ret in (glpk.GLP_ENOPFS, glpk.GLP_ENODFS) or ret == case["reduced_lp_ret"] != 0
        return
    assert ret == 0
    assert P.it_cnt == (case["reduced_lp"]["it_cnt"] if "reduced_lp" in case else 0)
    assert_basic_solution(P, case["unloaded"])


@pytest.mark.parametrize("name", MIP_CASES)
def test_generated_mips_presolve_on(name):
    case = T.CASES[name]
    P = T.facade_problem(case["problem"])
    io = glpk.IOCP({"presolve": glpk.GLP_ON, "binarize": glpk.GLP_ON if case["binarize"] else glpk.GLP_OFF})
    io.msg_lev = glpk.GLP_MSG_OFF
    ret = glpk.glp_intopt(P, io)
    if case["ret"] != 0:
        assert ret == case["ret"] and glpk.glp_mip_status(P) == glpk.GLP_UNDEF
        return
    un = case["unloaded"]
    assert ret == case["reduced_mip_ret"] == 0
    assert glpk.glp_mip_status(P) == un["mip_stat"]
    assert close(glpk.glp_mip_obj_val(P), un["mip_obj"])
    # the recovered point is feasible for the ORIGINAL problem and integral
    d = case["problem"]
    for j in range(1, P.n + 1):
        x = P.col[j].mipx
        if d["c_kind"][j - 1] == glpk.GLP_IV:
            assert x == round(x)
        if d["c_type"][j - 1] in (glpk.GLP_LO, glpk.GLP_DB, glpk.GLP_FX):
            assert x >= d["c_lb"][j - 1] - 1e-7
        if d["c_type"][j - 1] in (glpk.GLP_UP, glpk.GLP_DB):
            assert x <= d["c_ub"][j - 1] + 1e-7
    for i in range(1, P.m + 1):
        r = P.row[i].mipx
        if d["r_type"][i - 1] in (glpk.GLP_LO, glpk.GLP_DB, glpk.GLP_FX):
            assert r >= d["r_lb"][i - 1] - 1e-6
        if d["r_type"][i - 1] in (glpk.GLP_UP, glpk.GLP_DB):
            assert r <= d["r_ub"][i - 1] + 1e-6
        if d["r_type"][i - 1] == glpk.GLP_FX:
            assert r <= d["r_lb"][i - 1] + 1e-6
