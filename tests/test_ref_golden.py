"""The oracle against THE REFERENCE ITSELF.

tests/golden/ref_runs.json and ref_vectors.npz were produced by executing the
unmodified reference sources (/root/reference/lib/*.js) in the build container
with the ES5 interpreter of oracle/jsref (generator: make_ref_golden.py).  Here
the C++ oracle must reproduce, on the same inputs,

  * return code, status, objective, iteration count and solution of glp_simplex
    for the reference's fixtures and for generated LPs, primal and dual;
  * the MIP optimum AND the number of node LPs of glp_intopt;
  * the complete pivot sequence (q, p, p_stat / delta, step) of every solve;
  * the choice of chuzc / chuzr / both Harris ratio tests when its stateless
    selection routines are fed the reference's live arrays.

This is what pins the oracle; the GPU tests then compare the
device with the oracle and with the same reference vectors."""
import json
import os

import numpy as np
import pytest

import helpers as H
import oracle_lib as O

with open(os.path.join(H.GOLDEN, "ref_runs.json")) as _f:
    REF = json.load(_f)
VEC = np.load(os.path.join(H.GOLDEN, "ref_vectors.npz"))
METH = {"primal": O.GLP_PRIMAL, "dual": O.GLP_DUAL, "dualp": O.GLP_DUALP}


def group(prefix):
    """{field: array} of one captured call"""
    p = prefix + "/"
    return {k[len(p):]: VEC[k] for k in VEC.files if k.startswith(p)}


def groups(kind):
    """prefixes of all captured calls of one kind, e.g. 'p_chuzc'"""
    return sorted({k.rsplit("/", 1)[0] for k in VEC.files if k.split("/")[1].startswith(kind + "_")})


def oracle_problem(name):
    if name in ("test", "gap", "todd"):
        return O.Problem.from_arrays(H.load_golden(name))
    e = REF["generated"][name]
    if name.startswith("random_lp_"):
        d = H.random_lp(int(name.rsplit("_", 1)[1]))
        assert abs(float(np.sum(d["A_val"]) + np.sum(d["c_coef"])) - e["checksum"]) < 1e-9     # same generated instance
        return O.Problem.from_arrays(d)
    if name.startswith("random_mip_"):
        return O.Problem.from_arrays(H.random_mip(int(name.rsplit("_", 1)[1])))
    if name.startswith("mkp_"):
        m, n = name[4:].split("x")
        return O.Problem.from_arrays(H.to_oracle(O.generate("mkp", m=int(m), n=int(n), seed={"5x30": 20240701, "4x16": 11}[name[4:]])))
    g = dict(e["gen"])
    return O.Problem.from_arrays(H.to_oracle(O.generate(g.pop("kind"), **g)))


def oracle_trace(P, meth):
    """the pivot sequence in the shape make_ref_golden.py records it"""
    seq, cur = [], {}

    def hook(ev, csa):
        if ev == O.EV_P_CHUZC:
            s = O.csa_scalars(csa)
            cur.clear()
            cur.update(kind="P", q=s["q"])
            if s["q"] == 0:
                seq.append(dict(cur))
        elif ev == O.EV_P_CHUZR:
            s = O.csa_scalars(csa)
            cur.update(p=s["p"], p_stat=s["p_stat"], teta=s["teta"], phase=s["phase"])
            seq.append(dict(cur))
        elif ev == O.EV_D_CHUZR:
            s = O.csa_scalars(csa)
            cur.clear()
            cur.update(kind="D", p=s["p"], delta=s["delta"], phase=s["phase"])
            if s["p"] == 0:
                seq.append(dict(cur))
        elif ev == O.EV_D_CHUZC:
            s = O.csa_scalars(csa)
            cur.update(q=s["q"], new_dq=s["new_dq"])
            seq.append(dict(cur))
    P.set_hook(hook)
    rc = P.simplex(meth=meth)
    P.set_hook(None)
    return rc, seq


def close(a, b, rtol=1e-11):
    return abs(a - b) <= rtol * max(1.0, abs(a), abs(b))


TRACES = [(n, k[6:]) for n in ("test", "gap", "todd") for k in REF[n] if k.startswith("trace_")] + \
         [(n, k[6:]) for n, e in REF["generated"].items() for k in e if k.startswith("trace_")]


@pytest.mark.parametrize("name,mname", TRACES)
def test_oracle_reproduces_the_references_solve_and_pivot_sequence(name, mname):
    ref = (REF[name] if name in REF else REF["generated"][name])["trace_" + mname]
    P = oracle_problem(name)
    rc, seq = oracle_trace(P, METH[mname])
    s = P.solution()
    assert rc == ref["ret"] and s["status"] == ref["status"]
    assert s["pbs"] == ref["prim_stat"] and s["dbs"] == ref["dual_stat"]
    assert s["it_cnt"] == ref["it_cnt"], (s["it_cnt"], ref["it_cnt"])
    assert close(s["obj"], ref["obj"]), (s["obj"], ref["obj"])
    m = s["m"]
    np.testing.assert_array_equal(s["stat"][:m], ref["row_stat"])
    np.testing.assert_array_equal(s["stat"][m:], ref["col_stat"])
    np.testing.assert_allclose(s["prim"][m:], ref["col_prim"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(s["dual"][:m], ref["row_dual"], rtol=1e-9, atol=1e-9)
    # the pivot sequence: identical indices, steps to rounding
    rp = ref["pivots"]
    assert len(seq) == len(rp), (len(seq), len(rp))
    for i, (a, b) in enumerate(zip(seq, rp)):
        assert a["kind"] == b["kind"]
        for key in ("q", "p", "p_stat", "phase"):
            if key in b:
                assert a.get(key) == b[key], (i, key, a, b)
        for key in ("teta", "delta", "new_dq"):
            if key in b:
                assert close(a[key], b[key], 1e-9), (i, key, a, b)


@pytest.mark.parametrize("name", ["test", "gap", "todd"])
def test_oracle_reproduces_what_test_js_prints(name):
    """test/test.js: glp_simplex then glp_intopt, values printed; presolve OFF here (the oracle
    restates the path without the presolver) -- optimum and NODE COUNT equal the reference's"""
    ref = REF[name]["presolve_0"]
    P = O.Problem.from_arrays(H.load_golden(name))
    assert P.simplex(meth=O.GLP_PRIMAL) == ref["lp"]["ret"] == 0
    s = P.solution()
    assert s["status"] == ref["lp"]["status"] == 5 and s["it_cnt"] == ref["lp"]["it_cnt"]
    assert close(s["obj"], ref["lp"]["obj"])
    assert P.intopt() == ref["mip"]["ret"] == 0
    mp = P.mip()
    assert mp["mip_stat"] == ref["mip"]["mip_stat"] == 5 and mp["mip_obj"] == ref["mip"]["mip_obj"]
    assert mp["nodes"] == ref["mip"]["nodes_solved"]
    np.testing.assert_allclose(mp["mipx"][s["m"]:], ref["mip"]["col_val"], atol=1e-9)
    # with the presolver the reference reaches the same optimum
    assert REF[name]["presolve_1"]["mip"]["mip_obj"] == ref["mip"]["mip_obj"]
    assert close(REF[name]["presolve_1"]["lp"]["obj"], ref["lp"]["obj"])


@pytest.mark.parametrize("name", [n for n in REF["generated"] if n.startswith(("mkp_", "random_mip_"))])
def test_oracle_branch_and_bound_equals_the_references(name):
    ref = REF["generated"][name]
    P = oracle_problem(name)
    assert P.simplex(meth=O.GLP_PRIMAL) == ref["lp"]["ret"]
    s = P.solution()
    assert s["status"] == ref["lp"]["status"] and close(s["obj"], ref["lp"]["obj"]) and s["it_cnt"] == ref["lp"]["it_cnt"]
    if ref["mip"] is None:
        return
    assert P.intopt() == ref["mip"]["ret"]
    mp = P.mip()
    assert mp["mip_stat"] == ref["mip"]["mip_stat"]
    assert close(mp["mip_obj"], ref["mip"]["mip_obj"]) and mp["nodes"] == ref["mip"]["nodes_solved"]


# ---- the stateless selection routines fed the reference's live arrays ----
def test_chuzc_primal_on_reference_arrays():
    gs = groups("p_chuzc")
    assert len(gs) >= 30
    for g in gs:
        v = group(g)
        q = O.lib().glpo_chuzc_primal(int(v["n"]), O._p(v["stat"]), O._p(v["cbar"]), O._p(v["gamma"]), float(v["tol"]))
        assert q == int(v["q"]), g


def test_chuzr_dual_on_reference_arrays():
    import ctypes as C
    gs = groups("d_chuzr")
    assert len(gs) >= 30
    for g in gs:
        v = group(g)
        delta = C.c_double()
        p = O.lib().glpo_chuzr_dual(int(v["m"]), O._p(v["type"]), O._p(v["lb"]), O._p(v["ub"]), O._p(v["head"]),
                                    O._p(v["bbar"]), O._p(v["gamma"]), float(v["tol"]), C.byref(delta))
        assert p == int(v["p"]) and delta.value == float(v["delta"]), g


def test_ratio_tests_on_reference_arrays():
    import ctypes as C
    gp, gd = groups("p_chuzr"), groups("d_chuzc")
    assert len(gp) >= 30 and len(gd) >= 30
    for g in gp:
        v = group(g)
        p, ps, teta = C.c_int(), C.c_int(), C.c_double()
        q = int(v["q"])
        O.lib().glpo_chuzr_primal(int(v["m"]), O._p(v["type"]), O._p(v["lb"]), O._p(v["ub"]), O._p(v["coef"]),
                                  O._p(v["head"]), int(v["phase"]), O._p(v["bbar"]), float(v["cbar"][q]), q,
                                  O._p(v["tcol_ind"]), O._p(v["tcol_vec"]), int(v["tcol_num"]), float(v["rtol"]),
                                  C.byref(p), C.byref(ps), C.byref(teta))
        assert (p.value, ps.value) == (int(v["p"]), int(v["p_stat"])) and teta.value == float(v["teta"]), g
    for g in gd:
        v = group(g)
        q, ndq = C.c_int(), C.c_double()
        O.lib().glpo_chuzc_dual(O._p(v["stat"]), O._p(v["cbar"]), float(v["delta"]), O._p(v["trow_ind"]),
                                O._p(v["trow_vec"]), int(v["trow_num"]), float(v["rtol"]), C.byref(q), C.byref(ndq))
        assert q.value == int(v["q"]) and ndq.value == float(v["new_dq"]), g
