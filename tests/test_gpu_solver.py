"""GPU parity tests, solver level, through the C ABI: identical status, optimum
within 1e-9 relative, primal/dual residuals <= 1e-9, against the oracle and the
golden pins; plus the size-independent properties of the basis solves."""
import numpy as np
import pytest

import glpk_js_b200 as G
import oracle_lib as O
import helpers as H

nat = G.native
glpk = G.glpk
pytestmark = pytest.mark.gpu
METHS = [nat.GLP_PRIMAL, nat.GLP_DUAL, nat.GLP_DUALP]


def solve_both(dn, **kw):
    P = nat.Problem(dn)
    rc = P.simplex(**kw)
    s = P.solution()
    Q = O.Problem.from_arrays(H.to_oracle(dn))
    orc = Q.simplex(**kw)
    o = Q.solution()
    return P, rc, s, orc, o


def assert_parity(dn, rc, s, orc, o, what):
    assert rc == orc, (what, rc, orc)
    assert (s["status"], s["pbs"], s["dbs"]) == (o["status"], o["pbs"], o["dbs"]), what
    if o["status"] == O.GLP_OPT:
        assert abs(s["obj"] - o["obj"]) <= 1e-9 * max(1.0, abs(o["obj"])), (what, s["obj"], o["obj"])
        r = H.kkt(dn, s)
        assert max(r.values()) <= 1e-9, (what, r)


@pytest.mark.parametrize("name", ["test", "gap", "todd"])
@pytest.mark.parametrize("meth", METHS)
def test_fixtures_match_oracle_and_pins(name, meth):
    d = H.load_golden(name)
    dn = H.to_native(d)
    P, rc, s, orc, o = solve_both(dn, meth=meth)
    assert_parity(dn, rc, s, orc, o, (name, meth))
    assert abs(s["obj"] - d["highs_lp_obj"]) <= 1e-9 * max(1.0, abs(d["highs_lp_obj"]))
    P.close()


def test_test_lpt_follows_the_hand_trace():
    """no ties on this problem, so the device path must take the same two pivots"""
    d = H.load_golden("test")
    P = nat.Problem(H.to_native(d))
    assert P.simplex(meth=nat.GLP_PRIMAL) == 0
    s = P.solution()
    assert s["it_cnt"] == 2 and list(s["head"]) == d["hand_trace_primal"]["head_final"]
    np.testing.assert_allclose(s["prim"][3:], d["highs_lp_x"], atol=1e-9)
    np.testing.assert_allclose(s["dual"][:3], [10.0 / 3, 2.0 / 3, 0.0], atol=1e-9)
    P.close()


@pytest.mark.parametrize("meth", METHS)
@pytest.mark.parametrize("pricing", [nat.GLP_PT_PSE, nat.GLP_PT_STD])
@pytest.mark.parametrize("r_test", [nat.GLP_RT_HAR, nat.GLP_RT_STD])
def test_synthetic_small(meth, pricing, r_test):
    for which, kw in (("packing", dict(m=48, n=96, density=0.3, seed=21)),
                      ("covering", dict(m=96, n=192, kmin=3, kspan=4, seed=22))):
        dn = nat.generate(which, **kw)
        P, rc, s, orc, o = solve_both(dn, meth=meth, pricing=pricing, r_test=r_test)
        assert_parity(dn, rc, s, orc, o, (which, meth, pricing, r_test))
        P.close()


@pytest.mark.parametrize("which,kw,meth", [
    ("packing", dict(m=256, n=512, density=0.2, seed=20240501), nat.GLP_PRIMAL),
    ("covering", dict(m=1024, n=2048, kmin=8, kspan=17, seed=20240601), nat.GLP_DUAL),
])
def test_synthetic_medium_with_refactorisations(which, kw, meth):
    dn = nat.generate(which, **kw)
    P, rc, s, orc, o = solve_both(dn, meth=meth)
    assert_parity(dn, rc, s, orc, o, which)
    c = P.counters()
    assert c["refactorizations"] >= 2 and c["launches"] > 0, c
    P.close()


def test_infeasible_unbounded_and_limits():
    # infeasible: x1 + x2 <= 1, x1 + x2 >= 3
    P = glpk.glp_create_prob()
    txt = "Minimize\n obj: x1 + x2\nSubject To\n a: x1 + x2 <= 1\n b: x1 + x2 >= 3\nEnd\n"
    assert glpk.glp_read_lp_from_string(P, None, txt) == 0
    for meth in METHS:
        Q = O.Problem.from_lp(txt)
        orc = Q.simplex(meth=meth)
        parm = glpk.SMCP({"meth": meth})
        parm.msg_lev = glpk.GLP_MSG_OFF
        glpk.glp_std_basis(P)
        assert glpk.glp_simplex(P, parm) == orc
        assert glpk.glp_get_status(P) == Q.solution()["status"] == glpk.GLP_NOFEAS
    # unbounded: max x1 + x2 s.t. x1 - x2 <= 1
    txt = "Maximize\n obj: x1 + x2\nSubject To\n a: x1 - x2 <= 1\nEnd\n"
    assert glpk.glp_read_lp_from_string(P, None, txt) == 0
    Q = O.Problem.from_lp(txt)
    assert Q.simplex(meth=O.GLP_PRIMAL) == 0
    parm = glpk.SMCP()
    parm.msg_lev = glpk.GLP_MSG_OFF
    assert glpk.glp_simplex(P, parm) == 0
    assert glpk.glp_get_status(P) == Q.solution()["status"] == glpk.GLP_UNBND
    assert glpk.glp_get_unbnd_ray(P) == Q.solution()["some"]
    # iteration limit
    dn = nat.generate("packing", m=48, n=96, density=0.3, seed=21)
    R = nat.Problem(dn)
    assert R.simplex(meth=nat.GLP_PRIMAL, it_lim=5) == nat.GLP_EITLIM
    assert R.solution()["it_cnt"] == 5 and R.solution()["pbs"] == nat.GLP_FEAS
    assert R.simplex(meth=nat.GLP_PRIMAL) == 0          # warm start from the stored basis
    Q = O.Problem.from_arrays(H.to_oracle(dn))
    Q.simplex(meth=O.GLP_PRIMAL)
    assert abs(R.solution()["obj"] - Q.solution()["obj"]) <= 1e-9 * abs(Q.solution()["obj"])
    R.close()


def test_basis_solves_roundtrip_and_linearity():
    """ftran/btran against a dense numpy solve with the returned basis header;
    B * ftran(b) = b, B' * btran(c) = c, and linearity -- at a size where the
    structural kernel T is a few hundred wide."""
    dn = nat.generate("packing", m=256, n=512, density=0.2, seed=9)
    P = nat.Problem(dn)
    assert P.simplex(meth=nat.GLP_PRIMAL) == 0
    s = P.solution()
    B = H.basis_matrix(dn, s["head"])
    rng = np.random.default_rng(1)
    b1, b2 = rng.standard_normal(256), rng.standard_normal(256)
    x1, x2 = P.ftran(b1), P.ftran(b2)
    assert np.max(np.abs(B @ x1 - b1)) <= 1e-9 * (1 + np.abs(x1).max())
    np.testing.assert_allclose(P.ftran(2 * b1 - 3 * b2), 2 * x1 - 3 * x2, rtol=0, atol=1e-9 * (1 + np.abs(x1).max()))
    z = P.btran(b1)
    assert np.max(np.abs(B.T @ z - b1)) <= 1e-9 * (1 + np.abs(z).max())
    assert P.counters()["k"] == int((s["stat"][256:] == nat.GLP_BS).sum())
    P.close()


def test_facade_like_reference_test_js():
    """the flow of the reference's test/test.js cplex(): read LP, simplex, print"""
    lp = glpk.glp_create_prob()
    assert glpk.glp_read_lp_from_string(lp, None, H.golden_text("test")) == 0
    smcp = glpk.SMCP({"presolve": glpk.GLP_OFF})
    smcp.msg_lev = glpk.GLP_MSG_OFF
    assert glpk.glp_simplex(lp, smcp) == 0
    assert glpk.glp_get_status(lp) == glpk.GLP_OPT
    assert abs(glpk.glp_get_obj_val(lp) - 733.3333333333333) <= 1e-9 * 733.4
    x = [glpk.glp_get_col_prim(lp, j) for j in range(1, glpk.glp_get_num_cols(lp) + 1)]
    np.testing.assert_allclose(x, [33.333333333333336, 66.66666666666666, 0.0], atol=1e-9)
    assert [glpk.glp_get_col_name(lp, j) for j in (1, 2, 3)] == ["x1", "x2", "x3"]


@pytest.mark.parametrize("name", ["test", "gap", "todd"])
@pytest.mark.parametrize("flags", [glpk.GLP_SF_AUTO, glpk.GLP_SF_GM | glpk.GLP_SF_EQ | glpk.GLP_SF_2N,
                                   glpk.GLP_SF_EQ])
def test_scaled_problem_and_crash_basis_reach_the_same_optimum(name, flags):
    """glp_scale_prob + glp_adv_basis feed the path (rii/sjj into glpb_create,
    statuses into glpb_set_basis); the un-scaled solution must satisfy the KKT
    conditions of the ORIGINAL problem and match the pins."""
    d = H.load_golden(name)
    lp = glpk.glp_create_prob()
    assert glpk.glp_read_lp_from_string(lp, None, H.golden_text(name)) == 0
    glpk.glp_scale_prob(lp, flags)
    glpk.glp_adv_basis(lp, 0)
    parm = glpk.SMCP()
    parm.msg_lev = glpk.GLP_MSG_OFF
    for meth in (glpk.GLP_PRIMAL, glpk.GLP_DUALP):
        parm.meth = meth
        assert glpk.glp_simplex(lp, parm) == 0
        assert glpk.glp_get_status(lp) == glpk.GLP_OPT
        assert abs(glpk.glp_get_obj_val(lp) - d["highs_lp_obj"]) <= 1e-9 * max(1.0, abs(d["highs_lp_obj"]))
        m, n = lp.m, lp.n
        sol = dict(prim=np.array([lp.row[i].prim for i in range(1, m + 1)] + [lp.col[j].prim for j in range(1, n + 1)]),
                   dual=np.array([lp.row[i].dual for i in range(1, m + 1)] + [lp.col[j].dual for j in range(1, n + 1)]),
                   stat=np.array([lp.row[i].stat for i in range(1, m + 1)] + [lp.col[j].stat for j in range(1, n + 1)]))
        arrays, _, _ = glpk._arrays(lp)
        r = H.kkt(arrays, sol)
        assert max(r.values()) <= 1e-9, r


@pytest.mark.parametrize("name", ["test", "gap", "todd"])
def test_presolve_on_flow_of_test_js(name):
    """test/test.js cplex(): glp_simplex with presolve ON, then glp_intopt with
    presolve ON (copy, scaling, crash basis, solve, solution stored back)."""
    d = H.load_golden(name)
    lp = glpk.glp_create_prob()
    assert glpk.glp_read_lp_from_string(lp, None, H.golden_text(name)) == 0
    msgs = []
    glpk.glp_set_print_func(msgs.append)
    try:
        assert glpk.glp_simplex(lp, glpk.SMCP({"presolve": glpk.GLP_ON})) == 0
    finally:
        glpk.glp_set_print_func(None)
    assert msgs[0].startswith("GLPK Simplex Optimizer") and "Preprocessing..." in msgs and "Scaling..." in msgs
    assert "Constructing initial basis..." in msgs
    assert lp.valid == 0 and glpk.glp_get_status(lp) == glpk.GLP_OPT
    assert abs(glpk.glp_get_obj_val(lp) - d["highs_lp_obj"]) <= 1e-9 * max(1.0, abs(d["highs_lp_obj"]))
    assert all(glpk.glp_get_rii(lp, i) == 1.0 for i in range(1, lp.m + 1))   # the original stays unscaled
    m, n = lp.m, lp.n
    sol = dict(prim=np.array([lp.row[i].prim for i in range(1, m + 1)] + [lp.col[j].prim for j in range(1, n + 1)]),
               dual=np.array([lp.row[i].dual for i in range(1, m + 1)] + [lp.col[j].dual for j in range(1, n + 1)]),
               stat=np.array([lp.row[i].stat for i in range(1, m + 1)] + [lp.col[j].stat for j in range(1, n + 1)]))
    arrays, _, _ = glpk._arrays(lp)
    r = H.kkt(arrays, sol)
    assert max(r.values()) <= 1e-9, r
    if "highs_mip_obj" in d and d["highs_mip_obj"] is not None:
        iocp = glpk.IOCP({"presolve": glpk.GLP_ON})
        iocp.msg_lev = glpk.GLP_MSG_OFF
        assert glpk.glp_intopt(lp, iocp) == 0
        assert glpk.glp_mip_status(lp) == glpk.GLP_OPT
        assert glpk.glp_mip_obj_val(lp) == d["highs_mip_obj"]
        # and B&B without the presolver from the optimal basis a presolve:ON solve left (valid = 0)
        iocp2 = glpk.IOCP()
        iocp2.msg_lev = glpk.GLP_MSG_OFF
        assert glpk.glp_intopt(lp, iocp2) == 0
        assert glpk.glp_mip_obj_val(lp) == d["highs_mip_obj"]


def test_facade_factorize_ftran_btran_in_unscaled_space():
    """glp_factorize / glp_ftran / glp_btran (lib/glpapi12.js:5-222) on a SCALED problem:
    the facade applies R and SB around the device's scaled solves, so B x = b and
    B' y = b must hold for the unscaled basis matrix -- for the crash basis and for
    the optimal basis the simplex leaves."""
    lp = glpk.glp_create_prob()
    assert glpk.glp_read_lp_from_string(lp, None, H.golden_text("gap")) == 0
    glpk.glp_scale_prob(lp, glpk.GLP_SF_GM | glpk.GLP_SF_EQ)
    glpk.glp_adv_basis(lp, 0)
    arrays, _, _ = glpk._arrays(lp)
    rng = np.random.default_rng(5)
    m = lp.m

    def check():
        assert glpk.glp_bf_exists(lp)
        head = [glpk.glp_get_bhead(lp, i) for i in range(1, m + 1)]
        for i, k in enumerate(head, 1):
            assert (glpk.glp_get_row_bind(lp, k) if k <= m else glpk.glp_get_col_bind(lp, k - m)) == i
        B = H.basis_matrix(arrays, head)
        b = rng.uniform(-2, 2, m)
        x = [0.0] + b.tolist()
        glpk.glp_ftran(lp, x)
        np.testing.assert_allclose(B @ np.array(x[1:]), b, atol=1e-9)
        y = [0.0] + b.tolist()
        glpk.glp_btran(lp, y)
        np.testing.assert_allclose(B.T @ np.array(y[1:]), b, atol=1e-9)

    assert glpk.glp_factorize(lp) == 0
    check()
    parm = glpk.SMCP()
    parm.msg_lev = glpk.GLP_MSG_OFF
    assert glpk.glp_simplex(lp, parm) == 0 and glpk.glp_get_status(lp) == glpk.GLP_OPT
    check()
    # a singular basis: two identical structural columns cannot both be basic
    sp = glpk.glp_create_prob()
    glpk.glp_add_rows(sp, 2)
    glpk.glp_add_cols(sp, 2)
    for j in (1, 2):
        glpk.glp_set_mat_col(sp, j, 2, [0, 1, 2], [0.0, 1.0, 2.0])
        glpk.glp_set_col_bnds(sp, j, glpk.GLP_LO, 0.0, 0.0)
        glpk.glp_set_col_stat(sp, j, glpk.GLP_BS)
    for i in (1, 2):
        glpk.glp_set_row_bnds(sp, i, glpk.GLP_UP, 0.0, 4.0)
        glpk.glp_set_row_stat(sp, i, glpk.GLP_NU)
    assert glpk.glp_factorize(sp) in (glpk.GLP_ESING, glpk.GLP_ECOND) and not glpk.glp_bf_exists(sp)


def test_time_limit_is_honoured_inside_long_engine_runs():
    """smcp.tm_lim (lib/glpspx02.js:1801-1830): the host tests it between engine launches, which are kept short
    when a limit is set -- the covering LP 8192 x 16384 needs seconds, a 150 ms limit must come back as GLP_ETMLIM
    within a few launches of the limit and leave a basis the next call continues from"""
    import time
    d = nat.generate("covering", m=8192, n=16384, kmin=8, kspan=17, seed=20240601)
    P = nat.Problem(d)
    t0 = time.perf_counter()
    rc = P.simplex(meth=nat.GLP_DUAL, tm_lim=150)
    dt = time.perf_counter() - t0
    it1 = P.solution()["it_cnt"]
    assert rc == nat.GLP_ETMLIM, rc
    assert 0.15 <= dt <= 0.60, dt
    assert it1 > 100
    rc = P.simplex(meth=nat.GLP_DUAL)
    s = P.solution()
    assert rc == 0 and s["status"] == O.GLP_OPT and s["it_cnt"] > it1
    P.close()
