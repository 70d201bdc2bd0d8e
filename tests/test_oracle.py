"""CPU tests: the oracle against the pins, the host logic, the C-ABI surface.
(No CUDA device is needed; nothing here launches a kernel.)"""
import re
import os

import numpy as np
import pytest

import glpk_js_b200 as G
import oracle_lib as O
import helpers as H

nat = G.native
glpk = G.glpk
PINS = {"test": 733.3333333333333, "gap": 254.35771655880353, "todd": 4194303.5}


@pytest.mark.parametrize("name", ["test", "gap", "todd"])
@pytest.mark.parametrize("meth", [O.GLP_PRIMAL, O.GLP_DUAL, O.GLP_DUALP])
def test_oracle_matches_highs_pins(name, meth):
    d = H.load_golden(name)
    assert abs(d["highs_lp_obj"] - PINS[name]) <= 1e-9 * max(1, abs(PINS[name]))
    P = O.Problem.from_arrays(d)
    assert P.simplex(meth=meth) == 0
    s = P.solution()
    assert s["status"] == O.GLP_OPT
    assert abs(s["obj"] - PINS[name]) <= 1e-9 * max(1.0, abs(PINS[name]))
    r = H.kkt(H.to_native(d), s)
    assert max(r.values()) <= 1e-9, r


def test_oracle_hand_trace_test_lpt():
    """pivots (q=1,p=2),(q=2,p=1); cbar/gamma/bbar after iteration 1."""
    d = H.load_golden("test")
    tr = d["hand_trace_primal"]
    P = O.Problem.from_arrays(d)
    seen = []

    def hook(ev, csa):
        if ev == O.EV_P_CHUZR:
            s = O.csa_scalars(csa)
            seen.append((s["q"], s["p"], O.csa_get(csa, "cbar")[1:].copy(),
                         O.csa_get(csa, "gamma")[1:].copy(), O.csa_get(csa, "bbar")[1:].copy(),
                         O.csa_get(csa, "tcol_ind")[1:4].copy()))
    P.set_hook(hook)
    assert P.simplex(meth=O.GLP_PRIMAL) == 0
    assert [[q, p] for q, p, *_ in seen] == tr["pivots"]
    assert list(seen[0][5]) == [3, 1, 2]          # order left by sort_tcol
    np.testing.assert_allclose(seen[1][2], tr["cbar_after_1"], rtol=1e-12)
    np.testing.assert_allclose(seen[1][3], tr["gamma_after_1"], rtol=1e-12)
    np.testing.assert_allclose(seen[1][4], tr["bbar_after_1"], rtol=1e-12)
    s = P.solution()
    assert s["it_cnt"] == tr["it_cnt"] and list(s["head"]) == tr["head_final"]
    np.testing.assert_allclose(s["prim"][3:], d["highs_lp_x"], atol=1e-9)


def test_oracle_invariants_updated_vs_recomputed():
    """The reference's own self-check (err_in_bbar/cbar, lib/glpspx01.js:1257-1289):
    updated bbar/cbar must agree with values recomputed from scratch."""
    d = nat.generate("packing", m=40, n=80, density=0.3, seed=5)
    P = O.Problem.from_arrays(H.to_oracle(d))
    A = H.dense_A(d)
    worst = [0.0]

    def hook(ev, csa):
        if ev != O.EV_P_ITER:
            return
        s = O.csa_scalars(csa)
        m, n = s["m"], s["n"]
        head = O.csa_get(csa, "head")
        stat = O.csa_get(csa, "stat")
        lb, ub = O.csa_get(csa, "lb"), O.csa_get(csa, "ub")
        bbar = O.csa_get(csa, "bbar")[1:]
        full = np.hstack([np.eye(m), -A])
        B = full[:, head[1:m + 1] - 1]
        N = full[:, head[m + 1:] - 1]
        xn = np.array([{2: lb[head[m + j]], 3: ub[head[m + j]], 4: 0.0, 5: lb[head[m + j]]}[int(stat[j])]
                       for j in range(1, n + 1)])
        beta = np.linalg.solve(B, -N @ xn)
        worst[0] = max(worst[0], np.max(np.abs(beta - bbar) / (1 + np.abs(beta))))
    P.set_hook(hook)
    assert P.simplex(meth=O.GLP_PRIMAL) == 0
    assert worst[0] <= 1e-9


def test_lp_reader_roundtrip_and_python_reader_agree():
    for name in ("test", "gap", "todd"):
        d = H.load_golden(name)
        text = H.golden_text(name)
        e = O.Problem.from_lp(text).export()
        for k in ("r_type", "r_lb", "r_ub", "c_type", "c_lb", "c_ub", "c_coef", "c_kind", "A_ptr", "A_ind", "A_val"):
            np.testing.assert_array_equal(e[k], d[k], err_msg=name + ":" + k)
        P = glpk.glp_create_prob()
        assert glpk.glp_read_lp_from_string(P, None, text) == 0
        a, _, _ = glpk._arrays(P)
        n = H.to_native(d)
        for k in ("type", "lb", "ub", "coef", "kind", "A_ptr", "A_ind", "A_val"):
            np.testing.assert_array_equal(a[k], n[k], err_msg=name + ":" + k)
        assert (a["dir"], P.m, P.n, P.nnz) == (d["dir"], d["m"], d["n"], len(d["A_val"]))


def test_lp_reader_syntax_error_and_bounds_forms():
    P = glpk.glp_create_prob()
    assert glpk.glp_read_lp_from_string(P, None, "Maximize\n obj: x\nEnd\n") == 1   # no constraints
    txt = ("Minimize\n z: 2 a - b + 0 c\nSubject To\n c1: a + b >= -1.5\n c2: - a + 3 c = 2\n"
           "Bounds\n -inf <= a <= 4\n b free\n c >= 1\n 0 <= d <= 1\nBinary\n e\nGeneral\n d\nEnd\n")
    assert glpk.glp_read_lp_from_string(P, None, txt) == 0
    assert [P.col[j].type for j in range(1, 6)] == [glpk.GLP_UP, glpk.GLP_FR, glpk.GLP_LO, glpk.GLP_DB, glpk.GLP_DB]
    assert glpk.glp_get_num_int(P) == 2 and glpk.glp_get_num_bin(P) == 2
    assert P.row[1].type == glpk.GLP_LO and P.row[1].lb == -1.5 and P.row[2].type == glpk.GLP_FX
    Q = O.Problem.from_lp(txt)
    e = Q.export()
    a, _, _ = glpk._arrays(P)
    np.testing.assert_array_equal(a["type"], np.concatenate([e["r_type"], e["c_type"]]))
    np.testing.assert_array_equal(a["A_val"], e["A_val"])


def test_rng_and_generators_are_deterministic_and_match_oracle_rng():
    a = nat.rng_fill(20240501, 500)
    b = np.zeros(500, np.int32)
    O.lib().glpo_rng_fill(20240501, 500, b.ctypes.data_as(O.C.c_void_p))
    np.testing.assert_array_equal(a, b)
    assert a.min() >= 0
    d1 = nat.generate("covering", m=64, n=128, kmin=4, kspan=3, seed=7)
    d2 = nat.generate("covering", m=64, n=128, kmin=4, kspan=3, seed=7)
    np.testing.assert_array_equal(d1["A_ind"], d2["A_ind"])
    np.testing.assert_array_equal(d1["A_val"], d2["A_val"])
    rows = np.zeros(64, int)
    np.add.at(rows, d1["A_ind"], 1)
    assert rows.min() >= 1                       # every row covered
    lens = np.diff(d1["A_ptr"])
    assert lens[1:].min() >= 4 and lens[1:].max() <= 6
    k = nat.generate("mkp", m=5, n=20, seed=3)
    assert (k["kind"] == nat.GLP_IV).all() and (k["ub"][5:] == 1).all()


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(os.path.dirname(H.HERE), "include", "glpb200.h")).read()
    declared = set(re.findall(r"\b(glpb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"glpb_prob"}
    lib = nat.load()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(nat.SYMBOLS) == declared


def test_no_cpu_fallback_without_device():
    """Without a CUDA device every compute entry point must fail loudly."""
    lib = nat.load()
    if lib.glpb_device_count() > 0:
        pytest.skip("a GPU is present")
    d = nat.generate("packing", m=4, n=8, density=0.5, seed=1)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        nat.Problem(d)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        nat.k_chuzc_primal(3, np.zeros(4, np.int8), np.zeros(4), np.ones(4), 1e-7)
    P = glpk.glp_create_prob()
    glpk.glp_read_lp_from_string(P, None, H.golden_text("test"))
    with pytest.raises(RuntimeError):
        glpk.glp_simplex(P, glpk.SMCP({"msg_lev": glpk.GLP_MSG_ERR}))


def test_facade_parameter_checks_and_quirks():
    s = glpk.SMCP({"msg_lev": 0, "it_lim": 5})
    assert s.msg_lev == glpk.GLP_MSG_ALL and s.it_lim == 5     # the `|| default` quirk
    P = glpk.glp_create_prob()
    glpk.glp_read_lp_from_string(P, None, H.golden_text("test"))
    bad = glpk.SMCP()
    bad.tol_bnd = 2.0
    with pytest.raises(glpk.GlpkError, match="tol_bnd"):
        glpk.glp_simplex(P, bad)
    with pytest.raises(glpk.GlpkError, match="out of range"):
        glpk.glp_get_col_prim(P, 99)
    glpk.glp_set_col_bnds(P, 1, glpk.GLP_DB, 3.0, 1.0)
    assert glpk.glp_simplex(P, glpk.SMCP()) == glpk.GLP_EBOUND


# ---- the randomised LPs / MIPs of the GPU parity tests, pinned by HiGHS
# ---- (tests/golden/random_pins.json, generator tests/golden/make_random_pins.py)
def _pins():
    import json
    with open(os.path.join(H.GOLDEN, "random_pins.json")) as f:
        return json.load(f)


def _checksum(d):
    return float(np.sum(d["A_val"] * (1.0 + np.arange(len(d["A_val"])) % 7)) + np.sum(d["c_coef"]) + np.sum(d["r_lb"])
                 + np.sum(d["r_ub"]) + np.sum(d["c_lb"]) + np.sum(d["c_ub"]))


def test_oracle_matches_highs_on_random_lps():
    """the same instances test_gpu_random.py feeds the device: status class and
    optimum (1e-9 relative) of the oracle agree with HiGHS for every method"""
    seen = {"optimal": 0, "infeasible": 0, "unbounded": 0}
    for pin in _pins()["lp"]:
        d = H.random_lp(pin["seed"])
        assert (d["m"], d["n"]) == (pin["m"], pin["n"]) and abs(_checksum(d) - pin["checksum"]) < 1e-9, \
            "random_lp drifted from the pinned generator (seed %d)" % pin["seed"]
        for meth in (O.GLP_PRIMAL, O.GLP_DUAL, O.GLP_DUALP):
            Q = O.Problem.from_arrays(d)
            rc = Q.simplex(meth=meth)
            s = Q.solution()
            what = (pin["seed"], meth, rc, s["status"], s["pbs"], s["dbs"])
            if pin["highs"] == "optimal":
                assert rc == 0 and s["status"] == O.GLP_OPT, what
                assert abs(s["obj"] - pin["obj"]) <= 1e-9 * max(1.0, abs(pin["obj"])), (what, s["obj"], pin["obj"])
            elif pin["highs"] == "infeasible":
                # primal simplex proves it in phase 1; the dual method either proves it (dual
                # unbounded) or, without a dual feasible basis, gives up with an undefined status
                assert s["status"] != O.GLP_OPT, what
                if meth == O.GLP_PRIMAL:
                    assert rc == 0 and s["status"] == O.GLP_NOFEAS, what
            else:
                assert s["status"] != O.GLP_OPT, what
                if meth == O.GLP_PRIMAL:
                    assert rc == 0 and s["status"] == O.GLP_UNBND, what
        seen[pin["highs"]] += 1
    assert min(seen.values()) >= 10, seen


def test_oracle_matches_highs_on_random_mips():
    """identical MIP optimum (north_star) on the random integer programs of the
    GPU test; seed 8 is skipped there and here (26 802 nodes)"""
    n_opt = 0
    for pin in _pins()["mip"]:
        if pin["seed"] == 8:
            continue
        d = H.random_mip(pin["seed"])
        assert abs(_checksum(d) - pin["checksum"]) < 1e-9, "random_mip drifted (seed %d)" % pin["seed"]
        Q = O.Problem.from_arrays(d)
        assert Q.simplex(meth=O.GLP_PRIMAL) == 0 and Q.solution()["status"] == O.GLP_OPT
        assert abs(Q.solution()["obj"] - pin["lp_obj"]) <= 1e-9 * max(1.0, abs(pin["lp_obj"])), pin["seed"]
        assert Q.intopt() == 0
        mp = Q.mip()
        assert mp["mip_stat"] == O.GLP_OPT, pin["seed"]
        assert abs(mp["mip_obj"] - pin["obj"]) <= 1e-9 * max(1.0, abs(pin["obj"])), (pin["seed"], mp["mip_obj"], pin["obj"])
        n_opt += 1
    assert n_opt == 23


def test_oracle_matches_highs_on_knapsack_mips():
    """the C5 family (glpb_gen_mkp) at sizes the oracle's branch-and-bound finishes in seconds;
    the 5x30 instance is the one the GPU branching-rule tests use"""
    for pin in _pins()["mkp"]:
        d = H.to_oracle(nat.generate("mkp", m=pin["m"], n=pin["n"], seed=pin["seed"]))
        assert abs(_checksum(d) - pin["checksum"]) < 1e-9
        Q = O.Problem.from_arrays(d)
        assert Q.simplex(meth=O.GLP_PRIMAL) == 0
        assert abs(Q.solution()["obj"] - pin["lp_obj"]) <= 1e-9 * abs(pin["lp_obj"])
        assert Q.intopt() == 0 and Q.mip()["mip_stat"] == O.GLP_OPT
        assert abs(Q.mip()["mip_obj"] - pin["obj"]) <= 1e-9 * abs(pin["obj"]), (pin, Q.mip()["mip_obj"])


def test_napi_shim_is_well_formed_c_and_create_validates_the_column_pointer():
    """bindings/glpb200_napi.c cannot be loaded here (no Node.js): it is compiled against a
    declarations-only stub of node_api.h so that it stays valid C; the C ABI itself rejects a
    column pointer that is not a monotone prefix sum ending at nnz (no device needed for that)."""
    import subprocess
    root = os.path.abspath(os.path.join(H.HERE, ".."))
    out = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-Wno-unused-function", "-I" + os.path.join(root, "tests", "napi_stub"),
                          "-I" + os.path.join(root, "include"), os.path.join(root, "bindings", "glpb200_napi.c")],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    src = open(os.path.join(root, "bindings", "glpb200_napi.c")).read()
    for entry in ("Create", "SetBasis", "SetBounds", "Simplex", "Intopt", "GetSolution", "GetMip", "ScaleProb", "AdvBasis", "ReadLp", "WriteLp",
                  "NppCreate", "NppLoadProb", "NppSimplex", "NppInteger", "NppBuildProb", "NppPostprocess"):
        assert ("static napi_value %s(" % entry) in src
    assert src.count("typed_n(") >= 3 and "typed(env" not in src          # every array goes through the checked accessor
    import ctypes as C
    L = nat.load()
    t = np.array([3, 2, 2], np.int32)
    z = np.zeros(3)
    for bad_ptr in ([0, 2, 1], [1, 1, 2], [0, 1, 3]):
        ptr = np.array(bad_ptr, np.int32)
        h = L.glpb_create(1, 2, 2, 1, 0.0, t.ctypes.data, z.ctypes.data, z.ctypes.data, z[:2].ctypes.data, None, None, None,
                          ptr.ctypes.data, np.zeros(2, np.int32).ctypes.data, np.ones(2).ctypes.data, 0)
        assert not h and "A_ptr" in nat.last_error()


def test_oracle_own_c3_run_agrees_with_the_highs_pin_once_complete():
    """tests/golden/c3_oracle_run.json is written by the oracle's own uninterrupted solve of the headline LP
    (tests/golden/make_c3_pins.py oracle-c3; > 10 h on one core).  While it is a progress log there is nothing
    to check beyond its shape; once complete, the oracle's optimum must be the HiGHS pin's."""
    import json
    with open(os.path.join(H.GOLDEN, "c3_oracle_run.json")) as f:
        run = json.load(f)
    with open(os.path.join(H.GOLDEN, "lp_pins.json")) as f:
        pin = json.load(f)["c3"]["highs"]["obj"]
    its = [e["it"] for e in run["log"]]
    assert its == sorted(its) and all(e["k"] <= e["n"] for e in run["log"])
    if not run.get("partial", True):
        assert run["rc"] == 0 and run["status"] == O.GLP_OPT
        assert abs(run["obj"] - pin) <= 1e-9 * abs(pin)
        assert run["it_cnt"] >= its[-1]
