"""GPU parity tests, kernel level: every selection kernel is fed the oracle's
live CSA arrays (the reference's basis, reduced costs and weights) at every
iteration of real solves and must return the very same index (bit-exact),
through the C ABI (glpb_k_*)."""
import numpy as np
import pytest

import glpk_js_b200 as G
import oracle_lib as O
import helpers as H

nat = G.native
pytestmark = pytest.mark.gpu


def problems():
    yield "gap", H.load_golden("gap")
    yield "todd", H.load_golden("todd")
    yield "packing", H.to_oracle(nat.generate("packing", m=60, n=120, density=0.25, seed=3))
    yield "covering", H.to_oracle(nat.generate("covering", m=120, n=240, kmin=3, kspan=4, seed=4))


def boxed(d, seed):
    """give some columns upper bounds so that DB / bound-flip paths are hit"""
    rng = np.random.default_rng(seed)
    d = dict(d)
    ct, cu = d["c_type"].copy(), d["c_ub"].copy()
    pick = rng.random(d["n"]) < 0.4
    ct[pick & (ct == O.GLP_LO)] = O.GLP_DB
    cu[pick] = np.where(cu[pick] > 0, cu[pick], rng.integers(1, 4, pick.sum()))
    d["c_type"], d["c_ub"] = ct, cu
    return d


def test_primal_pricing_and_ratio_bit_exact_on_reference_state():
    checked = {"chuzc": 0, "chuzr": 0, "flip": 0, "phase1": 0}
    for name, d in problems():
        for variant in (d, boxed(d, 1)):
            P = O.Problem.from_arrays(variant)

            def hook(ev, csa):
                s = O.csa_scalars(csa)
                m, n = s["m"], s["n"]
                if ev == O.EV_P_CHUZC:
                    q = nat.k_chuzc_primal(n, O.csa_get(csa, "stat"), O.csa_get(csa, "cbar"),
                                           O.csa_get(csa, "gamma"), s["tol"])
                    assert q == s["q"], (name, s["it_cnt"], q, s["q"])
                    checked["chuzc"] += 1
                elif ev == O.EV_P_CHUZR:
                    cbar = O.csa_get(csa, "cbar")
                    p, p_stat, teta = nat.k_ratio_primal(
                        m, n, O.csa_get(csa, "type"), O.csa_get(csa, "lb"), O.csa_get(csa, "ub"),
                        O.csa_get(csa, "coef"), O.csa_get(csa, "head"), s["phase"], O.csa_get(csa, "bbar"),
                        float(cbar[s["q"]]), s["q"], O.csa_get(csa, "tcol_ind"), O.csa_get(csa, "tcol_vec"),
                        s["tcol_num"], s["tol"])
                    assert (p, teta) == (s["p"], s["teta"]), (name, s["it_cnt"], p, s["p"], teta, s["teta"])
                    if p != 0:
                        assert p_stat == s["p_stat"]
                    checked["chuzr"] += 1
                    checked["flip"] += p < 0
                    checked["phase1"] += s["phase"] == 1
            P.set_hook(hook)
            assert P.simplex(meth=O.GLP_PRIMAL) == 0
    assert checked["chuzc"] > 200 and checked["chuzr"] > 200 and checked["flip"] > 0 and checked["phase1"] > 0, checked


def test_dual_pricing_and_ratio_bit_exact_on_reference_state():
    checked = {"chuzr": 0, "chuzc": 0}
    for name, d in problems():
        for variant in (d, boxed(d, 2)):
            P = O.Problem.from_arrays(variant)

            def hook(ev, csa):
                s = O.csa_scalars(csa)
                m, n = s["m"], s["n"]
                if ev == O.EV_D_CHUZR:
                    p, delta = nat.k_chuzr_dual(m, n, O.csa_get(csa, "type"), O.csa_get(csa, "lb"),
                                                O.csa_get(csa, "ub"), O.csa_get(csa, "head"),
                                                O.csa_get(csa, "bbar"), O.csa_get(csa, "gamma"), s["tol"])
                    assert (p, delta) == (s["p"], s["delta"]), (name, s["it_cnt"])
                    checked["chuzr"] += 1
                elif ev == O.EV_D_CHUZC:
                    q, new_dq = nat.k_ratio_dual(n, O.csa_get(csa, "stat"), O.csa_get(csa, "cbar"), s["delta"],
                                                 O.csa_get(csa, "trow_ind"), O.csa_get(csa, "trow_vec"),
                                                 s["trow_num"], s["tol"])
                    assert (q, new_dq) == (s["q"], s["new_dq"]), (name, s["it_cnt"], q, s["q"])
                    checked["chuzc"] += 1
            P.set_hook(hook)
            assert P.simplex(meth=O.GLP_DUAL) == 0
    assert checked["chuzr"] > 200 and checked["chuzc"] > 200, checked


def test_pricing_ties_break_to_lowest_index():
    """Exact ties in d^2/gamma and in (ratio, |alfa|): lowest index / first in list."""
    rng = np.random.default_rng(0)
    for n in (1, 7, 300, 5000):
        stat = np.full(1 + n, O.GLP_NL, np.int8)
        cbar = np.zeros(1 + n)
        gamma = np.ones(1 + n)
        vals = rng.choice([-2.0, -1.0, 0.0, 1.0], size=n)
        cbar[1:] = vals
        stat[1:][rng.random(n) < 0.2] = O.GLP_NS
        stat[1:][rng.random(n) < 0.2] = O.GLP_NU
        stat[1:][rng.random(n) < 0.1] = O.GLP_NF
        want = O.lib().glpo_chuzc_primal(n, stat.ctypes.data, cbar.ctypes.data, gamma.ctypes.data, 1e-7)
        assert nat.k_chuzc_primal(n, stat, cbar, gamma, 1e-7) == want
    # empty / ineligible
    assert nat.k_chuzc_primal(3, np.full(4, O.GLP_NS, np.int8), -np.ones(4), np.ones(4), 1e-7) == 0
    # dual ratio test with many exact ties
    for n in (5, 64, 3000):
        stat = np.full(1 + n, O.GLP_NL, np.int8)
        cbar = np.zeros(1 + n)
        cbar[1:] = rng.choice([0.0, 1.0, 2.0], size=n)
        trow = np.zeros(1 + n)
        trow[1:] = rng.choice([0.5, 1.0, 2.0], size=n)
        ind = np.zeros(1 + n, np.int32)
        ind[1:] = rng.permutation(n) + 1
        q = np.zeros(1, np.int32)
        dq = np.zeros(1)
        O.lib().glpo_chuzc_dual(stat.ctypes.data, cbar.ctypes.data, 1.0, ind.ctypes.data, trow.ctypes.data, n,
                                3e-8, q.ctypes.data, dq.ctypes.data)
        assert nat.k_ratio_dual(n, stat, cbar, 1.0, ind, trow, n, 3e-8) == (int(q[0]), float(dq[0]))


def test_pivot_row_spmv_matches_reference_eval_trow():
    """eval_trow1 (column dots): warp-level summation order differs from the
    sequential loop, so the tolerance is 1e-12 relative to |rho|.|A_j|."""
    seen = [0]
    for name, d in problems():
        P = O.Problem.from_arrays(d)

        def hook(ev, csa):
            if ev != O.EV_D_TROW:
                return
            s = O.csa_scalars(csa)
            m, n = s["m"], s["n"]
            got = nat.k_trow(m, n, O.csa_get(csa, "A_ptr"), O.csa_get(csa, "A_ind"), O.csa_get(csa, "A_val"),
                             O.csa_get(csa, "head"), O.csa_get(csa, "stat"), O.csa_get(csa, "work4"))
            ref = O.csa_get(csa, "trow_vec")
            scale = 1.0 + np.abs(O.csa_get(csa, "work4")).max() * np.abs(O.csa_get(csa, "A_val")).max()
            assert np.max(np.abs(got[1:] - ref[1:])) <= 1e-12 * scale * 16
            seen[0] += 1
        P.set_hook(hook)
        assert P.simplex(meth=O.GLP_DUAL) == 0
    assert seen[0] > 100


# ======================================================================
# the same kernels fed THE REFERENCE'S OWN arrays (tests/golden/ref_vectors.npz:
# captured from lib/glpspx01.js / lib/glpspx02.js running in the minijs
# interpreter, oracle/jsref/make_ref_golden.py)
# ======================================================================
import os  # noqa: E402

VEC = np.load(os.path.join(H.GOLDEN, "ref_vectors.npz"))


def _group(prefix):
    p = prefix + "/"
    return {k[len(p):]: VEC[k] for k in VEC.files if k.startswith(p)}


def _groups(kind):
    return sorted({k.rsplit("/", 1)[0] for k in VEC.files if k.split("/")[1].startswith(kind + "_")})


def test_selection_kernels_bit_exact_on_the_references_arrays():
    n_checked = 0
    for g in _groups("p_chuzc"):
        v = _group(g)
        assert nat.k_chuzc_primal(int(v["n"]), v["stat"], v["cbar"], v["gamma"], float(v["tol"])) == int(v["q"]), g
        n_checked += 1
    for g in _groups("d_chuzr"):
        v = _group(g)
        p, delta = nat.k_chuzr_dual(int(v["m"]), int(v["n"]), v["type"], v["lb"], v["ub"], v["head"], v["bbar"], v["gamma"],
                                    float(v["tol"]))
        assert (p, delta) == (int(v["p"]), float(v["delta"])), g
        n_checked += 1
    for g in _groups("p_chuzr"):
        v = _group(g)
        q = int(v["q"])
        p, p_stat, teta = nat.k_ratio_primal(int(v["m"]), int(v["n"]), v["type"], v["lb"], v["ub"], v["coef"], v["head"],
                                             int(v["phase"]), v["bbar"], float(v["cbar"][q]), q, v["tcol_ind"], v["tcol_vec"],
                                             int(v["tcol_num"]), float(v["rtol"]))
        assert (p, teta) == (int(v["p"]), float(v["teta"])), g
        if p != 0:
            assert p_stat == int(v["p_stat"]), g
        n_checked += 1
    for g in _groups("d_chuzc"):
        v = _group(g)
        q, new_dq = nat.k_ratio_dual(int(v["n"]), v["stat"], v["cbar"], float(v["delta"]), v["trow_ind"], v["trow_vec"],
                                     int(v["trow_num"]), float(v["rtol"]))
        assert (q, new_dq) == (int(v["q"]), float(v["new_dq"])), g
        n_checked += 1
    assert n_checked >= 150


def test_pivot_row_kernel_on_the_references_rho():
    n_checked = 0
    for g in _groups("d_trow"):
        v = _group(g)
        got = nat.k_trow(int(v["m"]), int(v["n"]), v["A_ptr"], v["A_ind"], v["A_val"], v["head"], v["stat"], v["rho"])
        scale = 1.0 + np.abs(v["rho"]).max() * np.abs(v["A_val"]).max()
        assert np.max(np.abs(got[1:] - v["trow_vec"][1:])) <= 1e-12 * scale * 16, g
        n_checked += 1
    assert n_checked >= 20


# ---- sort_tcol / sort_trow: the list order that settles exact ties ----
def _literal_sort(vec, eps):
    """the reference's loop, statement for statement (lib/glpspx01.js:764-806): non-zeros in ascending order,
    then the swap loop that examines the last entry of the shrinking list"""
    n = len(vec) - 1
    ind = [0] + [i for i in range(1, n + 1) if vec[i] != 0.0]
    nnz, num = len(ind) - 1, 0
    while num < nnz:
        i = ind[nnz]
        if abs(vec[i]) < eps:
            nnz -= 1
        else:
            num += 1
            ind[nnz] = ind[num]
            ind[num] = i
    return ind[1:num + 1]


def test_sort_list_equals_the_references_swap_order():
    """k_sort_list (closed form, two scans) against (a) the lists the reference itself handed to its ratio tests
    (tests/golden/ref_vectors.npz: tcol_ind / trow_ind after sort_tcol / sort_trow, captured from lib/glpspx0[12].js
    run by minijs) and (b) the literal loop on random vectors with many insignificant and zero entries"""
    Z = np.load(H.GOLDEN + "/ref_vectors.npz")
    groups = sorted({k.rsplit("/", 1)[0] for k in Z.files if "/p_chuzr_" in k or "/d_chuzc_" in k})
    assert len(groups) >= 80
    checked = nontrivial = 0
    for g in groups:
        primal = "/p_chuzr_" in g
        vec = Z[g + ("/tcol_vec" if primal else "/trow_vec")]
        ind = Z[g + ("/tcol_ind" if primal else "/trow_ind")]
        num = int(Z[g + ("/tcol_num" if primal else "/trow_num")])
        n = len(vec) - 1
        big = float(np.max(np.abs(vec[1:]))) if n else 0.0
        eps = (1e-10 if primal else 1e-7) * (1.0 + 0.01 * big)      # smcp.tol_piv / sic tol_bnd (lib/glpspx02.js:1851)
        got = nat.k_sort_list(n, vec, eps)
        assert list(got) == [int(x) for x in ind[1:num + 1]], (g, list(got)[:10], list(ind[1:num + 1])[:10])
        checked += 1
        nontrivial += list(got) != sorted(got)
    assert checked >= 80 and nontrivial >= 50
    rng = np.random.default_rng(7)
    for trial in range(60):
        n = int(rng.integers(1, 5000)) if trial < 50 else int(rng.integers(40000, 70000))
        vec = np.zeros(1 + n)
        kind = rng.random(n)
        vec[1:] = np.where(kind < 0.3, 0.0, np.where(kind < 0.6, 1e-12 * rng.standard_normal(n), rng.standard_normal(n)))
        if trial % 7 == 0:
            vec[1:] = np.where(rng.random(n) < 0.5, 1.0, 0.0)      # everything significant or zero
        if trial % 11 == 0:
            vec[1:] = 1e-13                                        # nothing significant
        eps = 1e-9
        got = nat.k_sort_list(n, vec, eps)
        assert list(got) == _literal_sort(vec, eps), (trial, n)
