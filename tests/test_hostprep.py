"""CPU tests of the host-side preparation steps:
``glpb_scale_prob`` / ``glpb_adv_basis`` (csrc/hostprep.cpp, through the C ABI
and through the facade's ``glp_scale_prob`` / ``glp_adv_basis``) against the
plain-Python restatement of lib/glpscl.js and lib/glpini01.js in
oracle/hostprep.py -- bit-exact scale factors, identical statuses -- plus the
properties the algorithms guarantee.  No device is needed."""
import importlib.util
import json
import math
import os
import random

import numpy as np
import pytest

import glpk_js_b200 as G
import helpers as H

nat = G.native
glpk = G.glpk

_spec = importlib.util.spec_from_file_location(
    "oracle_hostprep", os.path.join(H.HERE, "..", "oracle", "hostprep.py"))
OH = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(OH)

ALL_FLAGS = [0, nat.GLP_SF_GM, nat.GLP_SF_EQ, nat.GLP_SF_2N, nat.GLP_SF_GM | nat.GLP_SF_EQ,
             nat.GLP_SF_GM | nat.GLP_SF_EQ | nat.GLP_SF_2N | nat.GLP_SF_SKIP, nat.GLP_SF_AUTO,
             nat.GLP_SF_EQ | nat.GLP_SF_2N, nat.GLP_SF_GM | nat.GLP_SF_SKIP]


def lists_of(P):
    """rows/cols of a facade problem in the reference's list order, 1-based"""
    rows = [None] + [list(P.row[i].elems) for i in range(1, P.m + 1)]
    cols = [None] + [list(P.col[j].elems) for j in range(1, P.n + 1)]
    return rows, cols


def read_fixture(name):
    P = glpk.glp_create_prob()
    glpk.glp_read_lp_from_string(P, None, H.golden_text(name))
    return P


def random_problem(seed, m, n, density, wide=False):
    """Facade problem with mixed row/column types (fixed, free, empty rows and
    columns included), values over many orders of magnitude when ``wide``."""
    rnd = random.Random(seed)
    P = glpk.glp_create_prob()
    glpk.glp_add_rows(P, m)
    glpk.glp_add_cols(P, n)
    types = [glpk.GLP_FR, glpk.GLP_LO, glpk.GLP_UP, glpk.GLP_DB, glpk.GLP_FX]
    for i in range(1, m + 1):
        t = rnd.choice(types)
        lb = rnd.uniform(-5, 5)
        glpk.glp_set_row_bnds(P, i, t, lb, lb + rnd.uniform(0.5, 4))
    for j in range(1, n + 1):
        t = rnd.choice(types)
        lb = rnd.uniform(-5, 5)
        glpk.glp_set_col_bnds(P, j, t, lb, lb + rnd.uniform(0.5, 4))
        glpk.glp_set_obj_coef(P, j, rnd.uniform(-3, 3))
    for j in range(1, n + 1):
        if rnd.random() < 0.05:
            continue  # empty column
        rows = [i for i in range(1, m + 1) if rnd.random() < density]
        rnd.shuffle(rows)
        vals = [(rnd.choice((-1, 1)) * (10.0 ** rnd.uniform(-4, 4) if wide else rnd.uniform(0.1, 3)))
                for _ in rows]
        glpk.glp_set_mat_col(P, j, len(rows), [0] + rows, [0.0] + vals)
    return P


def product_scale(P, flags):
    ptr, ind, val = glpk._csc(P)
    return nat.scale_prob(P.m, P.n, ptr, ind, val, flags)


def check_scale_parity(P, flags):
    rows, cols = lists_of(P)
    rii, sjj, rep = product_scale(P, flags)
    o_rii, o_sjj, o_rep = OH.scale_prob(P.m, P.n, rows, cols, flags)
    assert rii.tolist() == o_rii[1:], "rii differs (flags=%#x)" % flags     # bit-exact
    assert sjj.tolist() == o_sjj[1:], "sjj differs (flags=%#x)" % flags
    for ent in o_rep:
        if ent[0] == "skipped":
            assert rep["skipped"]
        else:
            assert rep[ent[0]] == tuple(ent[1:])
    assert rep["skipped"] == any(e[0] == "skipped" for e in o_rep)
    return rii, sjj, rep


def test_round2n_and_hand_case():
    """3x3 case worked by hand: A = diag-ish with entries 4, 1/4 ... ; one GM
    sweep on a matrix whose rows/cols each hold one entry makes every scaled
    entry exactly 1."""
    assert [OH.round2n(x) for x in (1.0, 0.75, 0.76, 1.5, 1.51, 3.0, 0.1)] == [1.0, 0.5, 1.0, 1.0, 2.0, 2.0, 0.125]
    P = glpk.glp_create_prob()
    glpk.glp_add_rows(P, 3)
    glpk.glp_add_cols(P, 3)
    for k, v in ((1, 4.0), (2, 0.25), (3, 64.0)):
        glpk.glp_set_mat_col(P, k, 1, [0, k], [0.0, v])
    rii, sjj, rep = check_scale_parity(P, nat.GLP_SF_GM)
    # rows and columns tie (ratio 1 each) -> rows first: rii = 1/|a|, then columns see 1
    assert rii.tolist() == [0.25, 4.0, 1.0 / 64.0] and sjj.tolist() == [1.0, 1.0, 1.0]
    assert rep["A"] == (0.25, 64.0, 256.0) and rep["GM"] == (1.0, 1.0, 1.0)


@pytest.mark.parametrize("name", ["test", "gap", "todd"])
def test_scale_fixtures_bit_exact(name):
    P = read_fixture(name)
    for flags in ALL_FLAGS:
        check_scale_parity(P, flags)


@pytest.mark.parametrize("seed", range(8))
def test_scale_random_bit_exact(seed):
    P = random_problem(100 + seed, m=7 + 3 * seed, n=9 + 4 * seed, density=0.3, wide=True)
    for flags in ALL_FLAGS:
        check_scale_parity(P, flags)


def test_scale_properties():
    P = random_problem(7, m=40, n=60, density=0.2, wide=True)
    ptr, ind, val = glpk._csc(P)
    cols = np.repeat(np.arange(P.n), np.diff(ptr))

    def scaled(rii, sjj):
        return np.abs(val) * rii[ind] * sjj[cols]

    rii, sjj, rep = product_scale(P, nat.GLP_SF_GM | nat.GLP_SF_EQ)
    s = scaled(rii, sjj)
    assert rep["GM"][2] < rep["A"][2]               # geometric mean scaling shrinks the spread
    assert s.max() <= 1.0 + 1e-12                    # equilibration: largest entry 1 ...
    colmax = np.zeros(P.n)
    np.maximum.at(colmax, cols, s)
    rowmax = np.zeros(P.m)
    np.maximum.at(rowmax, ind, s)
    # ... and the lines of the second pass all reach it
    assert (np.allclose(colmax[colmax > 0], 1.0, rtol=1e-12) or
            np.allclose(rowmax[rowmax > 0], 1.0, rtol=1e-12))
    rii2, sjj2, _ = product_scale(P, nat.GLP_SF_GM | nat.GLP_SF_EQ | nat.GLP_SF_2N)
    for x in list(rii2) + list(sjj2):
        mant, _ = math.frexp(x)
        assert mant == 0.5                           # exact powers of two
    assert np.all(rii2 / rii < 4.0 / 3 + 1e-12) and np.all(rii2 / rii >= 2.0 / 3 - 1e-12)
    # well-scaled data + SKIP: nothing changes
    Q = random_problem(8, m=20, n=30, density=0.3, wide=False)
    r, s_, rep = product_scale(Q, nat.GLP_SF_AUTO)
    assert rep["skipped"] and np.all(r == 1.0) and np.all(s_ == 1.0)


def test_scale_rejects_bad_arguments():
    ptr = np.array([0, 1], np.int32)
    with pytest.raises(ValueError):
        nat.scale_prob(1, 1, ptr, np.array([0], np.int32), np.array([1.0]), 0x02)   # unknown flag
    with pytest.raises(ValueError):
        nat.scale_prob(1, 1, ptr, np.array([3], np.int32), np.array([1.0]), 0)      # row out of range
    P = read_fixture("test")
    with pytest.raises(glpk.GlpkError):
        glpk.glp_scale_prob(P, 0x02)
    with pytest.raises(glpk.GlpkError):
        glpk.glp_set_rii(P, 1, 0.0)
    with pytest.raises(glpk.GlpkError):
        glpk.glp_adv_basis(P, 1)


def oracle_adv(P):
    rows, cols = lists_of(P)
    m, n = P.m, P.n
    return OH.adv_basis(
        m, n, rows, cols,
        [0] + [P.row[i].type for i in range(1, m + 1)], [0] + [P.row[i].lb for i in range(1, m + 1)],
        [0] + [P.row[i].ub for i in range(1, m + 1)],
        [0] + [P.col[j].type for j in range(1, n + 1)], [0] + [P.col[j].lb for j in range(1, n + 1)],
        [0] + [P.col[j].ub for j in range(1, n + 1)])


def check_adv_parity(P):
    tagx, size = oracle_adv(P)   # its own asserts check the triangular form
    glpk.glp_adv_basis(P, 0)
    got = [P.row[i].stat for i in range(1, P.m + 1)] + [P.col[j].stat for j in range(1, P.n + 1)]
    assert got == tagx[1:]
    assert P.tri_size == size
    assert sum(1 for s in got if s == glpk.GLP_BS) == P.m
    return got, size


@pytest.mark.parametrize("name", ["test", "gap", "todd"])
def test_adv_basis_fixtures(name):
    P = read_fixture(name)
    got, size = check_adv_parity(P)
    assert size == P.m   # every fixture row has a slack or a singleton: full triangle


@pytest.mark.parametrize("seed", range(12))
def test_adv_basis_random(seed):
    P = random_problem(300 + seed, m=5 + 4 * seed, n=6 + 5 * seed, density=0.25)
    if seed % 3 == 0:
        glpk.glp_sort_matrix(P)   # another list order, another tie-breaking
    got, size = check_adv_parity(P)
    # the basis is non-singular: permuted lower triangular with non-zero diagonal
    m, n = P.m, P.n
    A = np.zeros((m, n))
    for j in range(1, n + 1):
        for (i, v) in P.col[j].elems:
            A[i - 1, j - 1] = v
    B = np.zeros((m, m))
    c = 0
    for k, s in enumerate(got):
        if s == glpk.GLP_BS:
            if k < m:
                B[k, c] = 1.0
            else:
                B[:, c] = -A[:, k - m]
            c += 1
    assert np.linalg.matrix_rank(B) == m
    # fixed variables never enter the triangular part; they may only be basic as slacks of leftover rows
    for j in range(1, n + 1):
        if P.col[j].type == glpk.GLP_FX:
            assert P.col[j].stat == glpk.GLP_NS


def test_adv_basis_all_rows_fixed_takes_structurals():
    """Equality rows have no usable slack: the triangle must be built from
    structural columns (a bidiagonal system has a full one)."""
    m = 6
    P = glpk.glp_create_prob()
    glpk.glp_add_rows(P, m)
    glpk.glp_add_cols(P, m)
    for i in range(1, m + 1):
        glpk.glp_set_row_bnds(P, i, glpk.GLP_FX, 1.0, 1.0)
        glpk.glp_set_col_bnds(P, i, glpk.GLP_LO, 0.0, 0.0)
    for j in range(1, m + 1):
        rows = [j] + ([j + 1] if j < m else [])
        glpk.glp_set_mat_col(P, j, len(rows), [0] + rows, [0.0] + [1.0] * len(rows))
    got, size = check_adv_parity(P)
    assert size == m
    assert all(P.col[j].stat == glpk.GLP_BS for j in range(1, m + 1))
    assert all(P.row[i].stat == glpk.GLP_NS for i in range(1, m + 1))


def test_facade_scaling_state():
    """glp_scale_prob writes the factors, marks the device copy stale and
    invalidates the factorisation only when a basic column is touched
    (lib/glpapi04.js:5-13,22-26); no flags = unscale only (test/test.js:76)."""
    P = read_fixture("gap")
    msgs = []
    glpk.glp_set_print_func(msgs.append)
    try:
        glpk.glp_scale_prob(P, glpk.GLP_SF_GM | glpk.GLP_SF_EQ)
    finally:
        glpk.glp_set_print_func(None)
    assert msgs[0] == "Scaling..." and msgs[1].startswith(" A: min|aij| = ") and any(s.startswith("EQ:") for s in msgs)
    assert any(glpk.glp_get_rii(P, i) != 1.0 for i in range(1, P.m + 1))
    assert P._dirty
    P.valid = 1   # all structurals are non-basic after reading: column factors do not matter
    P._dirty = False
    glpk.glp_set_sjj(P, 1, 3.0)
    assert P.valid == 1 and P._dirty
    glpk.glp_set_rii(P, 1, 5.0)
    assert P.valid == 1
    glpk.glp_set_col_stat(P, 1, glpk.GLP_BS)
    P.valid = 1
    glpk.glp_set_sjj(P, 1, 2.0)
    assert P.valid == 0
    glpk.glp_scale_prob(P)
    assert all(glpk.glp_get_rii(P, i) == 1.0 for i in range(1, P.m + 1))
    assert all(glpk.glp_get_sjj(P, j) == 1.0 for j in range(1, P.n + 1))


def test_hostprep_large_is_linear_time():
    """C3-sized covering LP: both steps finish in seconds (O(nnz))."""
    import time
    d = nat.generate("covering", m=16384, n=32768)
    t0 = time.time()
    rii, sjj, rep = nat.scale_prob(d["m"], d["n"], d["A_ptr"], d["A_ind"], d["A_val"],
                                   nat.GLP_SF_GM | nat.GLP_SF_EQ | nat.GLP_SF_2N)
    # rows in index order from the CSC
    order = np.argsort(d["A_ind"], kind="stable")
    cols = np.repeat(np.arange(d["n"]), np.diff(d["A_ptr"])).astype(np.int32)
    R_ptr = np.zeros(d["m"] + 1, np.int32)
    np.cumsum(np.bincount(d["A_ind"], minlength=d["m"]), out=R_ptr[1:])
    stat, size = nat.adv_basis(d["m"], d["n"], d["A_ptr"], d["A_ind"], R_ptr, cols[order],
                               d["type"], d["lb"], d["ub"])
    assert time.time() - t0 < 20.0
    assert rep["2N"][2] <= rep["A"][2] * 4
    assert size == d["m"] and int((stat == nat.GLP_BS).sum()) == d["m"]


# ---- glp_simplex on an LP without constraint coefficients (lib/glpapi06.js:148-255):
# ---- solved on the host, no device involved
def _trivial(dir_, cols, rows=(), meth=None):
    P = glpk.glp_create_prob()
    glpk.glp_set_obj_dir(P, dir_)
    if rows:
        glpk.glp_add_rows(P, len(rows))
        for i, (t, lb, ub) in enumerate(rows, 1):
            glpk.glp_set_row_bnds(P, i, t, lb, ub)
    glpk.glp_add_cols(P, len(cols))
    for j, (t, lb, ub, c) in enumerate(cols, 1):
        glpk.glp_set_col_bnds(P, j, t, lb, ub)
        glpk.glp_set_obj_coef(P, j, c)
    parm = glpk.SMCP()
    if meth:
        parm.meth = meth
    msgs = []
    glpk.glp_set_print_func(msgs.append)
    try:
        ret = glpk.glp_simplex(P, parm)
    finally:
        glpk.glp_set_print_func(None)
    return P, ret, msgs


def test_trivial_lp_optimal():
    P, ret, msgs = _trivial(glpk.GLP_MIN,
                            [(glpk.GLP_LO, 1.0, 0.0, 2.0), (glpk.GLP_UP, 0.0, 4.0, -3.0),
                             (glpk.GLP_DB, -1.0, 5.0, 1.0), (glpk.GLP_DB, -1.0, 5.0, -1.0),
                             (glpk.GLP_DB, -1.0, 5.0, 0.0), (glpk.GLP_DB, -7.0, 5.0, 0.0),
                             (glpk.GLP_FX, 2.5, 2.5, 4.0), (glpk.GLP_FR, 0.0, 0.0, 0.0)],
                            rows=[(glpk.GLP_UP, 0.0, 3.0), (glpk.GLP_FR, 0.0, 0.0)])
    assert ret == 0 and glpk.glp_get_status(P) == glpk.GLP_OPT
    assert [P.col[j].stat for j in range(1, 9)] == [glpk.GLP_NL, glpk.GLP_NU, glpk.GLP_NL, glpk.GLP_NU,
                                                    glpk.GLP_NL, glpk.GLP_NU, glpk.GLP_NS, glpk.GLP_NF]
    assert [P.col[j].prim for j in range(1, 9)] == [1.0, 4.0, -1.0, 5.0, -1.0, 5.0, 2.5, 0.0]
    assert glpk.glp_get_obj_val(P) == 2.0 - 12.0 - 1.0 - 5.0 + 10.0
    assert [P.col[j].dual for j in range(1, 9)] == [2.0, -3.0, 1.0, -1.0, 0.0, 0.0, 4.0, 0.0]
    assert all(P.row[i].stat == glpk.GLP_BS and P.row[i].prim == 0.0 for i in (1, 2))
    assert msgs == ["GLPK Simplex Optimizer, v4.49", "2 rows, 8 columns, 0 non-zeros",
                    "~0: obj = -6  infeas = 0", "OPTIMAL SOLUTION FOUND"]
    assert P.valid == 0


def test_trivial_lp_unbounded_and_infeasible():
    # min -x, x >= 0: dual infeasible; with the primal method `some` names the ray
    P, ret, msgs = _trivial(glpk.GLP_MIN, [(glpk.GLP_LO, 0.0, 0.0, -1.0)])
    assert ret == 0 and glpk.glp_get_status(P) == glpk.GLP_UNBND and glpk.glp_get_unbnd_ray(P) == 1
    assert msgs[-1] == "PROBLEM HAS UNBOUNDED SOLUTION" and msgs[-2] == "~0: obj = 0  infeas = 0"
    P, ret, msgs = _trivial(glpk.GLP_MAX, [(glpk.GLP_UP, 0.0, 0.0, -2.0)], meth=glpk.GLP_DUAL)
    assert glpk.glp_get_dual_stat(P) == glpk.GLP_NOFEAS and P.some == 0
    assert msgs[-1] == "PROBLEM HAS NO DUAL FEASIBLE SOLUTION" and msgs[-2] == "~0: obj = 0  infeas = 1"
    # a row 0 >= 2 cannot hold
    P, ret, msgs = _trivial(glpk.GLP_MIN, [(glpk.GLP_LO, 0.0, 0.0, 1.0)], rows=[(glpk.GLP_LO, 2.0, 0.0)],
                            meth=glpk.GLP_DUAL)
    assert ret == 0 and glpk.glp_get_prim_stat(P) == glpk.GLP_NOFEAS and P.some == 1
    assert msgs[-1] == "PROBLEM HAS NO FEASIBLE SOLUTION"


def test_incorrect_bounds_and_parameter_checks():
    P = read_fixture("test")
    glpk.glp_set_col_bnds(P, 1, glpk.GLP_DB, 0.0, 1.0)
    P.col[1].ub = -1.0
    msgs = []
    glpk.glp_set_print_func(msgs.append)
    try:
        assert glpk.glp_simplex(P, glpk.SMCP()) == glpk.GLP_EBOUND
        assert glpk.glp_intopt(P, glpk.IOCP()) == glpk.GLP_EBOUND
    finally:
        glpk.glp_set_print_func(None)
    assert msgs == ["glp_simplex: column 1: lb = 0, ub = -1; incorrect bounds",
                    "glp_intopt: column 1: lb = 0, ub = -1; incorrect bounds"]
    bad = glpk.IOCP()
    bad.pp_tech = 7
    with pytest.raises(glpk.GlpkError):
        glpk.glp_intopt(P, bad)
    assert glpk._num(1e-7) == "1e-7" and glpk._num(2.5e+30) == "2.5e+30" and glpk._num(3.0) == "3"


# ---- glp_write_lp (lib/glpcpx.js:755-998) and the reader, round trip
def _same_problem(P, Q):
    assert (P.m, P.n, P.nnz, P.dir) == (Q.m, Q.n, Q.nnz, Q.dir)
    for i in range(1, P.m + 1):
        a, b = P.row[i], Q.row[i]
        assert a.type == b.type and glpk.glp_get_row_lb(P, i) == glpk.glp_get_row_lb(Q, i)
        assert glpk.glp_get_row_ub(P, i) == glpk.glp_get_row_ub(Q, i)
        assert sorted(a.elems) == sorted(b.elems)
    for j in range(1, P.n + 1):
        a, b = P.col[j], Q.col[j]
        assert (a.type, a.kind, a.coef) == (b.type, b.kind, b.coef)
        assert glpk.glp_get_col_lb(P, j) == glpk.glp_get_col_lb(Q, j)
        assert glpk.glp_get_col_ub(P, j) == glpk.glp_get_col_ub(Q, j)


@pytest.mark.parametrize("name", ["test", "gap", "todd"])
def test_write_lp_round_trip_fixtures(name):
    P = read_fixture(name)
    lines = []
    assert glpk.glp_write_lp(P, None, lines.append) == 0
    assert lines[0] == "\\* Problem: Unknown *\\" and lines[-1] == "End"
    assert all(len(s) <= 72 for s in lines)
    Q = glpk.glp_create_prob()
    # the reference's callback convention: one character per call, -1 at the end (test/test.js:36-43)
    text, pos = "\n".join(lines) + "\n", [0]

    def getc():
        if pos[0] < len(text):
            pos[0] += 1
            return text[pos[0] - 1]
        return -1

    assert glpk.glp_read_lp(Q, None, getc) == 0
    _same_problem(P, Q)
    assert ([glpk.glp_get_col_name(Q, j) for j in range(1, Q.n + 1)] ==
            [glpk.glp_get_col_name(P, j) for j in range(1, P.n + 1)])
    again = []
    glpk.glp_write_lp(Q, None, again.append)
    assert again == lines                      # write(read(write(P))) is a fixed point


def test_write_lp_round_trip_random_and_special_rows():
    P = random_problem(41, m=12, n=15, density=0.3)
    for i in range(1, P.m + 1):                # ranged rows become an extra column: tested below
        if P.row[i].type in (glpk.GLP_DB, glpk.GLP_FR):
            glpk.glp_set_row_bnds(P, i, glpk.GLP_UP, 0.0, 3.5)
    glpk.glp_set_col_kind(P, 2, glpk.GLP_IV)
    glpk.glp_set_col_name(P, 3, "has blank")   # invalid in LP format -> x_3
    glpk.glp_set_obj_coef(P, 0, 12.5)
    lines = []
    glpk.glp_write_lp(P, None, lines.append)
    assert "\\* constant term = 12.5 *\\" in lines and " x_2" in lines and "Generals" in lines
    Q = glpk.glp_create_prob()
    assert glpk.glp_read_lp_from_string(Q, None, "\n".join(lines)) == 0
    # columns are numbered in order of first appearance when read back: compare by name
    name_p = {("x_%d" % j): j for j in range(1, P.n + 1)}
    assert Q.m == P.m and Q.n == P.n and Q.nnz == P.nnz
    for jq in range(1, Q.n + 1):
        jp = name_p[glpk.glp_get_col_name(Q, jq)]
        a, b = P.col[jp], Q.col[jq]
        assert (a.type, a.kind, a.coef) == (b.type, b.kind, b.coef)
        assert glpk.glp_get_col_lb(P, jp) == glpk.glp_get_col_lb(Q, jq)
        assert glpk.glp_get_col_ub(P, jp) == glpk.glp_get_col_ub(Q, jq)
        assert sorted(a.elems) == sorted(b.elems)
    # a ranged row -> "... - ~r_i = lb" plus "0 <= ~r_i <= ub - lb"; a free row is dropped
    R = read_fixture("test")
    glpk.glp_set_row_bnds(R, 1, glpk.GLP_DB, 10.0, 100.0)
    glpk.glp_set_row_bnds(R, 3, glpk.GLP_FR, 0.0, 0.0)
    lines = []
    glpk.glp_write_lp(R, None, lines.append)
    assert " p: + x1 + x2 + x3 - ~r_1 = 10" in lines and " 0 <= ~r_1 <= 90" in lines
    assert not any(s.startswith(" r:") for s in lines)
    E = glpk.glp_create_prob()
    lines = []
    glpk.glp_write_lp(E, None, lines.append)
    assert lines == ["\\* Problem: Unknown *\\", "", "\\* WARNING: PROBLEM HAS NO ROWS/COLUMNS *\\", "", "End"]


# ---- basis factorisation interface, host-side behaviour (lib/glpapi12.js:1-176)
def test_factorize_host_checks_and_bfcp():
    P = read_fixture("test")
    assert not glpk.glp_bf_exists(P)
    for who in (lambda: glpk.glp_get_bhead(P, 1), lambda: glpk.glp_get_row_bind(P, 1),
                lambda: glpk.glp_ftran(P, [0.0] * 4), lambda: glpk.glp_btran(P, [0.0] * 4)):
        with pytest.raises(glpk.GlpkError, match="basis factorization does not exist"):
            who()
    glpk.glp_set_col_stat(P, 1, glpk.GLP_BS)            # four basic variables for three rows
    assert glpk.glp_factorize(P) == glpk.GLP_EBADB
    for i in (1, 2):
        glpk.glp_set_row_stat(P, i, glpk.GLP_NU)        # two basic variables
    assert glpk.glp_factorize(P) == glpk.GLP_EBADB and P.valid == 0
    parm = {}
    glpk.glp_get_bfcp(P, parm)
    assert parm["nfs_max"] == 100 and parm["piv_tol"] == 0.10 and parm["type"] == glpk.GLP_BF_FT
    glpk.glp_set_bfcp(P, {"nfs_max": 50})
    glpk.glp_get_bfcp(P, parm)
    assert parm["nfs_max"] == 50 and parm["rs_size"] == 2000
    with pytest.raises(glpk.GlpkError, match="nfs_max = 0; invalid parameter"):
        glpk.glp_set_bfcp(P, {"nfs_max": 0})
    with pytest.raises(glpk.GlpkError, match="piv_tol"):
        glpk.glp_set_bfcp(P, {"piv_tol": 1.5})
    glpk.glp_set_bfcp(P, None)
    glpk.glp_get_bfcp(P, parm)
    assert parm["nfs_max"] == 100


# ---- native LP reader (glpb_read_lp, csrc/lpformat.cpp) against the facade's reader
def _facade_arrays(text):
    """the Python restatement of the reader (glpk._read_lp), the native reader's cross-check"""
    P = glpk.glp_create_prob()
    glpk._read_lp(P, text)
    Q = glpk.glp_create_prob()                     # and the facade entry point (native reader inside)
    assert glpk.glp_read_lp_from_string(Q, None, text) == 0
    _same_problem(P, Q)
    assert [(r.name, r.stat) for r in P.row[1:]] == [(r.name, r.stat) for r in Q.row[1:]]
    assert [(c.name, c.stat, c.lb, c.ub) for c in P.col[1:]] == [(c.name, c.stat, c.lb, c.ub) for c in Q.col[1:]]
    assert [r.elems for r in P.row[1:]] == [r.elems for r in Q.row[1:]]
    assert [c.elems for c in P.col[1:]] == [c.elems for c in Q.col[1:]]
    assert P.obj == Q.obj and Q.valid == 0 and Q._dirty
    d, _, _ = glpk._arrays(P)
    names = dict(obj=P.obj, rows=[P.row[i].name for i in range(1, P.m + 1)],
                 cols=[P.col[j].name for j in range(1, P.n + 1)])
    return d, names


def _same_arrays(a, b):
    assert (a["m"], a["n"], a["dir"], a["c0"]) == (b["m"], b["n"], b["dir"], b["c0"])
    for k in ("type", "lb", "ub", "coef", "kind", "A_ptr", "A_ind", "A_val"):
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k


TRICKY_LP = """\\ a comment line
MAXIMIZE
 profit: 3 x + 2.5e0 y - z + 0 w
Subject To
 c1: x + y + z <= 10
 - x + 2 y =< 1.5E+1
 c3: x - y => -4
 end_like: 2 end + x = 3
 x + .5 bin >= 0
such that
Bounds
 -inf <= y <= 8
 z free
 -3 <= w <= 3
 x <= 40
 1 <= end
 2 <= bin <= 2
Generals
 x z
Binaries
 b1 b2
End
"""


def test_native_lp_reader_matches_the_facade_reader():
    for name in ("test", "gap", "todd"):
        text = H.golden_text(name)
        d, names = nat.read_lp(text)
        fd, fnames = _facade_arrays(text)
        _same_arrays(d, fd)
        assert names == fnames
    # the "such that" on its own line is a second constraints keyword in the wrong place: a syntax error for both
    bad = TRICKY_LP
    P = glpk.glp_create_prob()
    with pytest.raises(glpk.GlpkError):
        glpk._read_lp(P, bad)
    msgs = []
    glpk.glp_set_print_func(msgs.append)
    try:
        assert glpk.glp_read_lp_from_string(P, None, bad) == 1 and P.m == 0 and P.n == 0
    finally:
        glpk.glp_set_print_func(None)
    assert msgs == ["Reading problem data", "glp_read_lp: line 10: symbol such in wrong position"]
    with pytest.raises(ValueError, match="line 10: symbol such in wrong position"):
        nat.read_lp(bad)
    good = TRICKY_LP.replace("such that\n", "")
    d, names = nat.read_lp(good)
    fd, fnames = _facade_arrays(good)
    _same_arrays(d, fd)
    assert names == fnames
    assert names["cols"] == ["x", "y", "z", "w", "end", "bin", "b1", "b2"] and names["rows"][1] == "r.6"   # "r." + line number
    t = dict(zip(names["cols"], d["type"][d["m"]:].tolist()))
    assert t == {"x": glpk.GLP_DB, "y": glpk.GLP_UP, "z": glpk.GLP_FR, "w": glpk.GLP_DB, "end": glpk.GLP_LO,
                 "bin": glpk.GLP_FX, "b1": glpk.GLP_DB, "b2": glpk.GLP_DB}
    assert d["kind"].tolist() == [2, 1, 2, 1, 1, 1, 2, 2] and d["dir"] == glpk.GLP_MAX


@pytest.mark.parametrize("seed", range(6))
def test_native_lp_reader_on_written_random_problems(seed):
    P = random_problem(500 + seed, m=6 + 5 * seed, n=8 + 6 * seed, density=0.3, wide=(seed % 2 == 1))
    glpk.glp_set_col_kind(P, 1, glpk.GLP_IV)
    glpk.glp_set_obj_dir(P, glpk.GLP_MAX if seed % 2 else glpk.GLP_MIN)
    lines = []
    glpk.glp_write_lp(P, None, lines.append)
    text = "\n".join(lines) + "\n"
    d, names = nat.read_lp(text)
    fd, fnames = _facade_arrays(text)
    _same_arrays(d, fd)
    assert names == fnames


@pytest.mark.parametrize("text,what", [
    ("x + y\nst\n c: x <= 1\nend\n", "keyword missing"),
    ("min\n x\n", "constraints section missing"),
    ("min\n x + x\nst\n c: x <= 1\nend\n", "multiple use"),
    ("min\n x\nst\n c: x <= 1\n c: x >= 0\nend\n", "multiply defined"),
    ("min\n x\nst\n c: x 1\nend\n", "missing constraint sense"),
    ("min\n x\nst\n c: x <= y\nend\n", "missing right-hand side"),
    ("min\n x\nst\n c: x <= 1\nbounds\n +inf <= x\nend\n", "invalid use of `\\+inf'"),
    ("min\n x\nst\n c: x <= 1\nbounds\n 1 <= x >= 2\nend\n", "invalid bound definition"),
    ("min\n x\nst\n c: x <= 1\nend\n extra\n", "beyond `end'"),
    ("min\n x ^ 2\nst\n c: x <= 1\nend\n", "not recognized"),
])
def test_native_lp_reader_errors_like_the_facade(text, what):
    P = glpk.glp_create_prob()
    with pytest.raises(glpk.GlpkError):
        glpk._read_lp(P, text)
    assert glpk.glp_read_lp_from_string(P, None, text) == 1
    with pytest.raises(ValueError, match=what):
        nat.read_lp(text)


# ---- glp_intopt with a user callback: host-driven slices (no device: a scripted stand-in)
class _ScriptedDevice:
    """the slice interface of native.Problem, replaying a search of five nodes in which
    the incumbent improves after nodes 2 and 4"""
    m, n = 3, 3

    def __init__(self):
        self.node, self.log = 0, []
        self.script = [(False, 0.0), (True, 700.0), (True, 700.0), (True, 720.0), (True, 720.0)]

    def iocp(self, **kw):
        return kw

    def mip_begin(self, ip):
        self.log.append("begin")
        return 0

    def mip_open_count(self):
        return 5 - self.node

    def mip_run(self, k):
        assert k == 1
        self.node += 1
        return (1 if self.node < 5 else 0), 1

    def mip_incumbent(self):
        return self.script[self.node - 1]

    def mip_end(self, ret):
        self.log.append("end %d" % ret)
        return ret

    def intopt(self, ip):
        raise AssertionError("one-call search must not be used when a callback is given")

    def mip(self):
        return dict(mip_stat=glpk.GLP_OPT, mip_obj=720.0, mipx=np.array([100.0, 600.0, 300.0, 30.0, 70.0, 0.0]), nodes=5)

    def solution(self):
        return dict(pbs=glpk.GLP_FEAS, dbs=glpk.GLP_FEAS, obj=733.0, it_cnt=9, some=0, head=np.array([4, 5, 3]),
                    stat=np.array([3, 3, 1, 1, 1, 2]), prim=np.zeros(6), dual=np.zeros(6))


def test_intopt_callback_slices():
    P = read_fixture("test")
    P._dev, P._dirty, P.valid = _ScriptedDevice(), False, 1
    P.pbs_stat = P.dbs_stat = glpk.GLP_FEAS
    seen = []

    def cb(tree, info):
        assert info == "ctx" and glpk.glp_ios_get_prob(tree) is P
        seen.append((glpk.glp_ios_reason(tree), glpk.glp_mip_obj_val(glpk.glp_ios_get_prob(tree))))

    parm = glpk.IOCP({"cb_func": cb, "cb_info": "ctx"})
    parm.msg_lev = glpk.GLP_MSG_OFF
    assert glpk.glp_intopt(P, parm) == 0
    bingo = [(r, v) for r, v in seen if r == glpk.GLP_IBINGO]
    assert bingo == [(glpk.GLP_IBINGO, 700.0), (glpk.GLP_IBINGO, 720.0)]
    assert sum(1 for r, _ in seen if r == glpk.GLP_ISELECT) == 5
    assert P._dev.log == ["begin", "end 0"]
    assert glpk.glp_mip_status(P) == glpk.GLP_OPT and glpk.glp_mip_obj_val(P) == 720.0
    assert glpk.glp_mip_col_val(P, 1) == 30.0


# ---- the prepared inputs are good inputs: oracle solves from the product's scale factors and
# ---- crash basis reach the HiGHS-pinned optimum on the random LPs of the GPU tests
def test_scaled_crash_started_oracle_solves_reach_the_pins():
    import json
    import ctypes as C
    import oracle_lib as O
    with open(os.path.join(H.GOLDEN, "random_pins.json")) as f:
        pins = [p for p in json.load(f)["lp"] if p["highs"] == "optimal"]
    assert len(pins) >= 15
    for pin in pins:
        d = H.random_lp(pin["seed"])
        m, n = d["m"], d["n"]
        rii, sjj, rep = nat.scale_prob(m, n, d["A_ptr"], d["A_ind"], d["A_val"],
                                       nat.GLP_SF_GM | nat.GLP_SF_EQ | nat.GLP_SF_2N)
        # rows of A (ascending columns) for the crash basis
        cols = np.repeat(np.arange(n), np.diff(d["A_ptr"])).astype(np.int32)
        order = np.argsort(d["A_ind"], kind="stable")
        R_ptr = np.zeros(m + 1, np.int32)
        np.cumsum(np.bincount(d["A_ind"], minlength=m), out=R_ptr[1:])
        dn = H.to_native(d)
        stat, size = nat.adv_basis(m, n, d["A_ptr"], d["A_ind"], R_ptr, cols[order], dn["type"], dn["lb"], dn["ub"])
        assert int((stat == nat.GLP_BS).sum()) == m and 0 < size <= m
        for meth in (O.GLP_PRIMAL, O.GLP_DUALP):
            Q = O.Problem.from_arrays(d)
            Q.L.glpo_set_scale.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            Q.L.glpo_set_scale(Q.h, rii.ctypes.data_as(C.c_void_p), sjj.ctypes.data_as(C.c_void_p))
            Q.set_stat(stat)
            assert Q.simplex(meth=meth) == 0, (pin["seed"], meth)
            s = Q.solution()
            assert s["status"] == O.GLP_OPT
            assert abs(s["obj"] - pin["obj"]) <= 1e-9 * max(1.0, abs(pin["obj"])), (pin["seed"], s["obj"], pin["obj"])
            r = H.kkt(dn, s)                  # un-scaled solution against the ORIGINAL problem
            assert max(r.values()) <= 1e-9, (pin["seed"], r)


# ---- problem-object housekeeping (lib/glpapi01.js, lib/glpapi03.js)
def test_check_dup_and_name_index():
    ia, ja = [0, 1, 2, 2, 1, 2], [0, 1, 3, 2, 1, 3]
    assert glpk.glp_check_dup(2, 3, 3, ia, ja) == 0
    assert glpk.glp_check_dup(2, 3, 4, ia, ja) == 4          # (1,1) again at k = 4
    assert glpk.glp_check_dup(2, 3, 5, ia, ja) == 4          # row 1 is examined first
    assert glpk.glp_check_dup(2, 2, 3, ia, ja) == -2         # column 3 out of range
    assert glpk.glp_check_dup(0, 0, 0, None, None) == 0
    with pytest.raises(glpk.GlpkError):
        glpk.glp_check_dup(-1, 0, 0, None, None)
    assert glpk.glp_version() == "4.49"
    P = read_fixture("test")
    with pytest.raises(glpk.GlpkError, match="row name index does not exist"):
        glpk.glp_find_row(P, "p")
    glpk.glp_create_index(P)
    assert glpk.glp_find_row(P, "q") == 2 and glpk.glp_find_col(P, "x3") == 3 and glpk.glp_find_col(P, "nope") == 0
    glpk.glp_set_col_name(P, 3, "z")                         # the index follows a rename
    assert glpk.glp_find_col(P, "x3") == 0 and glpk.glp_find_col(P, "z") == 3
    glpk.glp_delete_index(P)
    with pytest.raises(glpk.GlpkError):
        glpk.glp_find_row(P, "p")
    assert glpk.glp_find_col(P, "z") == 3                    # the reference only drops the row index


def test_del_rows_del_cols_renumber_consistently():
    P = read_fixture("gap")
    m, n, nnz = P.m, P.n, P.nnz
    dense = np.zeros((m + 1, n + 1))
    for j in range(1, n + 1):
        for (i, v) in P.col[j].elems:
            dense[i, j] = v
    names_r = [None] + [P.row[i].name for i in range(1, m + 1)]
    names_c = [None] + [P.col[j].name for j in range(1, n + 1)]
    glpk.glp_create_index(P)
    P.valid = 1
    glpk.glp_del_rows(P, 2, [0, 3, 7])
    keep_r = [i for i in range(1, m + 1) if i not in (3, 7)]
    assert P.m == m - 2 and P.valid == 0 and [P.row[t].name for t in range(1, P.m + 1)] == [names_r[i] for i in keep_r]
    assert glpk.glp_find_row(P, names_r[3]) == 0 and glpk.glp_find_row(P, names_r[8]) == 6
    glpk.glp_del_cols(P, 3, [0, 1, 10, n])
    keep_c = [j for j in range(1, n + 1) if j not in (1, 10, n)]
    assert P.n == n - 3 and [P.col[t].name for t in range(1, P.n + 1)] == [names_c[j] for j in keep_c]
    # both copies of the matrix describe the remaining sub-matrix, with the new numbers
    got = np.zeros((P.m + 1, P.n + 1))
    for j in range(1, P.n + 1):
        assert P.col[j].j == j
        for (i, v) in P.col[j].elems:
            got[i, j] = v
    got_r = np.zeros_like(got)
    for i in range(1, P.m + 1):
        assert P.row[i].i == i
        for (j, v) in P.row[i].elems:
            got_r[i, j] = v
    want = dense[np.ix_([0] + keep_r, [0] + keep_c)]
    assert np.array_equal(got, want) and np.array_equal(got_r, want)
    assert P.nnz == int(np.count_nonzero(want)) and P._dirty
    with pytest.raises(glpk.GlpkError, match="duplicate row numbers"):
        glpk.glp_del_rows(P, 2, [0, 2, 2])
    with pytest.raises(glpk.GlpkError, match="out of range"):
        glpk.glp_del_cols(P, 1, [0, 999])


def test_copy_and_erase_prob():
    P = read_fixture("gap")
    glpk.glp_scale_prob(P, glpk.GLP_SF_EQ)
    glpk.glp_set_bfcp(P, {"nfs_max": 77})
    Q = glpk.glp_create_prob()
    glpk.glp_copy_prob(Q, P, glpk.GLP_ON)
    _same_problem(P, Q)
    assert [c.name for c in Q.col[1:]] == [c.name for c in P.col[1:]] and Q.obj == P.obj
    assert [r.rii for r in Q.row[1:]] == [r.rii for r in P.row[1:]] and [c.sjj for c in Q.col[1:]] == [c.sjj for c in P.col[1:]]
    assert [c.kind for c in Q.col[1:]] == [c.kind for c in P.col[1:]]
    assert Q.col[1].elems == list(reversed(P.col[1].elems))   # glp_set_mat_col prepends
    parm = {}
    glpk.glp_get_bfcp(Q, parm)
    assert parm["nfs_max"] == 77
    R = glpk.glp_create_prob()
    glpk.glp_copy_prob(R, P, glpk.GLP_OFF)
    assert all(c.name is None for c in R.col[1:]) and R.obj is None
    with pytest.raises(glpk.GlpkError, match="itself"):
        glpk.glp_copy_prob(P, P, glpk.GLP_ON)
    glpk.glp_erase_prob(Q)
    assert (Q.m, Q.n, Q.nnz, Q.dir, Q.c0) == (0, 0, 0, glpk.GLP_MIN, 0.0) and Q.row == [None] and Q._dev is None
    glpk.glp_add_rows(Q, 1)                                   # still a usable problem object
    assert Q.m == 1


# ---- glp_warm_up, simplex-table rows/columns, textbook ratio tests (lib/glpapi12.js:244-762):
# ---- facade logic over glp_ftran / glp_btran, checked with a dense stand-in for the device's solves
class _DenseDevice:
    """what glpb_factorize / glpb_ftran / glpb_btran do, in the SCALED space, by dense LU"""

    def __init__(self, P):
        import types
        self.P, self.h = P, None
        self.L = types.SimpleNamespace(glpb_set_it_cnt=lambda h, it: 0)

    def set_bounds(self, *a):
        return 0

    def set_basis(self, stat):
        return 0

    def factorize(self):
        P, m = self.P, self.P.m
        B = np.zeros((m, m))
        for pos in range(1, m + 1):
            k = P.head[pos]
            if k <= m:
                B[k - 1, pos - 1] = 1.0
            else:
                col = P.col[k - m]
                for (i, v) in col.elems:
                    B[i - 1, pos - 1] = -P.row[i].rii * v * col.sjj
        if abs(np.linalg.det(B)) < 1e-12:
            return glpk.GLP_ESING
        self.B = B
        return 0

    def ftran(self, b):
        return np.linalg.solve(self.B, np.asarray(b, float))

    def btran(self, b):
        return np.linalg.solve(self.B.T, np.asarray(b, float))


def _solved_with_oracle(seed):
    import oracle_lib as O
    d = H.random_lp(seed)
    Q = O.Problem.from_arrays(d)
    assert Q.simplex(meth=O.GLP_PRIMAL) == 0
    sol = Q.solution()
    assert sol["status"] == O.GLP_OPT
    P = glpk.glp_create_prob()
    m, n = d["m"], d["n"]
    glpk.glp_add_rows(P, m)
    glpk.glp_add_cols(P, n)
    glpk.glp_set_obj_dir(P, d["dir"])
    glpk.glp_set_obj_coef(P, 0, d["c0"])
    for i in range(1, m + 1):
        glpk.glp_set_row_bnds(P, i, int(d["r_type"][i - 1]), float(d["r_lb"][i - 1]), float(d["r_ub"][i - 1]))
    for j in range(1, n + 1):
        glpk.glp_set_col_bnds(P, j, int(d["c_type"][j - 1]), float(d["c_lb"][j - 1]), float(d["c_ub"][j - 1]))
        glpk.glp_set_obj_coef(P, j, float(d["c_coef"][j - 1]))
        a, b = d["A_ptr"][j - 1], d["A_ptr"][j]
        glpk.glp_set_mat_col(P, j, b - a, [0] + (d["A_ind"][a:b] + 1).tolist(), [0.0] + d["A_val"][a:b].tolist())
    glpk.glp_scale_prob(P, glpk.GLP_SF_GM | glpk.GLP_SF_EQ)
    for i in range(1, m + 1):
        glpk.glp_set_row_stat(P, i, int(sol["stat"][i - 1]))
    for j in range(1, n + 1):
        glpk.glp_set_col_stat(P, j, int(sol["stat"][m + j - 1]))
    P._dev, P._dirty = _DenseDevice(P), False
    return P, d, sol


def test_warm_up_reproduces_the_solution_of_the_basis():
    import json
    with open(os.path.join(H.GOLDEN, "random_pins.json")) as f:
        pins = [p for p in json.load(f)["lp"] if p["highs"] == "optimal"][:8]
    for pin in pins:
        P, d, sol = _solved_with_oracle(pin["seed"])
        assert glpk.glp_warm_up(P) == 0
        assert glpk.glp_get_prim_stat(P) == glpk.GLP_FEAS and glpk.glp_get_dual_stat(P) == glpk.GLP_FEAS
        assert abs(glpk.glp_get_obj_val(P) - pin["obj"]) <= 1e-9 * max(1.0, abs(pin["obj"]))
        m, n = P.m, P.n
        prim = np.array([P.row[i].prim for i in range(1, m + 1)] + [P.col[j].prim for j in range(1, n + 1)])
        dual = np.array([P.row[i].dual for i in range(1, m + 1)] + [P.col[j].dual for j in range(1, n + 1)])
        np.testing.assert_allclose(prim, sol["prim"], rtol=0, atol=1e-8)
        np.testing.assert_allclose(dual, sol["dual"], rtol=0, atol=1e-8)
    # a basis that is not primal feasible is reported as such
    P, d, sol = _solved_with_oracle(pins[0]["seed"])
    for j in range(1, P.n + 1):
        c = P.col[j]
        if c.stat == glpk.GLP_NL and c.type == glpk.GLP_DB:
            glpk.glp_set_col_stat(P, j, glpk.GLP_NU)
    P.valid = 0
    assert glpk.glp_warm_up(P) == 0 and glpk.glp_get_dual_stat(P) in (glpk.GLP_FEAS, glpk.GLP_INFEAS)


def test_simplex_table_rows_columns_and_ratio_tests():
    import json
    with open(os.path.join(H.GOLDEN, "random_pins.json")) as f:
        seed = next(p["seed"] for p in json.load(f)["lp"] if p["highs"] == "optimal" and p["m"] >= 10 and p["n"] >= 12)
    P, d, sol = _solved_with_oracle(seed)
    assert glpk.glp_warm_up(P) == 0
    m, n = P.m, P.n
    A = H.dense_A(d)
    full = np.hstack([np.eye(m), -A])                       # (I | -A), unscaled
    head = [glpk.glp_get_bhead(P, i) for i in range(1, m + 1)]
    nonbasic = [k for k in range(1, m + n + 1) if k not in head]
    T = -np.linalg.solve(full[:, [k - 1 for k in head]], full[:, [k - 1 for k in nonbasic]])   # x_B = T x_N
    ind, val = [0] * (1 + m + n), [0.0] * (1 + m + n)
    for pos in (1, m // 2 + 1, m):
        ln = glpk.glp_eval_tab_row(P, head[pos - 1], ind, val)
        row = np.zeros(m + n + 1)
        row[ind[1:ln + 1]] = val[1:ln + 1]
        np.testing.assert_allclose(row[nonbasic], T[pos - 1], atol=1e-9)
        # dual ratio test on that row against a direct evaluation
        for direction in (+1, -1):
            piv = glpk.glp_dual_rtest(P, ln, ind, val, direction, 1e-9)
            obj = 1.0 if P.dir == glpk.GLP_MIN else -1.0
            best = None
            for t in range(1, ln + 1):
                x = P.row[ind[t]] if ind[t] <= m else P.col[ind[t] - m]
                alfa = direction * val[t]
                if (x.stat == glpk.GLP_NL and alfa >= 1e-9) or (x.stat == glpk.GLP_NU and alfa <= -1e-9):
                    ratio = max(0.0, obj * x.dual / alfa)
                elif x.stat == glpk.GLP_NF and abs(alfa) >= 1e-9:
                    ratio = 0.0
                else:
                    continue
                if best is None or ratio < best[0] or (ratio == best[0] and abs(alfa) > best[1]):
                    best = (ratio, abs(alfa), t)
            assert piv == (best[2] if best else 0)
    for k in nonbasic[:3] + nonbasic[-2:]:
        ln = glpk.glp_eval_tab_col(P, k, ind, val)
        col = np.zeros(m + n + 1)
        col[ind[1:ln + 1]] = val[1:ln + 1]
        np.testing.assert_allclose(col[head], T[:, nonbasic.index(k)], atol=1e-9)
        piv = glpk.glp_prim_rtest(P, ln, ind, val, +1, 1e-9)
        assert 0 <= piv <= ln
        if piv:                                              # the chosen variable really blocks the increase
            x = P.row[ind[piv]] if ind[piv] <= m else P.col[ind[piv] - m]
            assert x.stat == glpk.GLP_BS and x.type != glpk.GLP_FR
    # transform_row of an original row reproduces the table row of its auxiliary variable when that is basic
    for i in range(1, m + 1):
        if P.row[i].stat == glpk.GLP_BS:
            ln0 = len(P.row[i].elems)
            ri, rv = [0] + [j for (j, _) in P.row[i].elems] + [0] * (m + n), [0.0] + [v for (_, v) in P.row[i].elems] + [0.0] * (m + n)
            ln = glpk.glp_transform_row(P, ln0, ri, rv)
            a = np.zeros(m + n + 1)
            a[ri[1:ln + 1]] = rv[1:ln + 1]
            np.testing.assert_allclose(a[nonbasic], T[head.index(i)], atol=1e-9)
            break
    # transform_col of an original column = table column of that (non-basic) variable, sign as in N = -A
    j = next(k - m for k in nonbasic if k > m)
    ci = [0] + [i for (i, _) in P.col[j].elems] + [0] * m
    cv = [0.0] + [v for (_, v) in P.col[j].elems] + [0.0] * m
    ln = glpk.glp_transform_col(P, len(P.col[j].elems), ci, cv)
    c = np.zeros(m + n + 1)
    c[ci[1:ln + 1]] = cv[1:ln + 1]
    np.testing.assert_allclose(c[head], T[:, nonbasic.index(m + j)], atol=1e-9)
    with pytest.raises(glpk.GlpkError, match="must be basic"):
        glpk.glp_eval_tab_row(P, nonbasic[0], ind, val)
    with pytest.raises(glpk.GlpkError, match="must be non-basic"):
        glpk.glp_eval_tab_col(P, head[0], ind, val)


# ---- glpb_read_lp against the REFERENCE'S OWN reader (tests/golden/ref_reader_cases.json: lib/glpcpx.js
#      executed by minijs, oracle/jsref/make_reader_golden.py) ----
def _reader_cases():
    import json
    import os
    with open(os.path.join(H.GOLDEN, "ref_reader_cases.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", sorted(_reader_cases()))
def test_native_reader_equals_the_references_reader(name):
    rec = _reader_cases()[name]
    if rec["rc"] != 0:
        # same message; the reference prefixes the line number ("4: invalid symbol(s) ...")
        line, msg = rec["error"].split(": ", 1)
        with pytest.raises(ValueError) as ei:
            nat.read_lp(rec["text"])
        got = str(ei.value)
        assert msg in got, (got, rec["error"])
        if name not in ("missing_sense", "missing_rhs", "row_twice", "extra_after_end", "plus_inf_lower"):
            assert ("line %s:" % line) in got, (got, rec["error"])      # errors raised while scanning: same line
        return
    d, names = nat.read_lp(rec["text"])
    m, n = rec["m"], rec["n"]
    assert (d["m"], d["n"], d["dir"]) == (m, n, rec["dir"]) and names["obj"] == rec["obj_name"]
    assert names["rows"] == [r["name"] for r in rec["rows"]] and names["cols"] == [c["name"] for c in rec["cols"]]
    assert d["type"][:m].tolist() == [r["type"] for r in rec["rows"]]
    assert d["type"][m:].tolist() == [c["type"] for c in rec["cols"]]
    assert d["lb"][:m].tolist() == [r["lb"] for r in rec["rows"]] and d["ub"][:m].tolist() == [r["ub"] for r in rec["rows"]]
    assert d["lb"][m:].tolist() == [c["lb"] for c in rec["cols"]] and d["ub"][m:].tolist() == [c["ub"] for c in rec["cols"]]
    assert d["coef"].tolist() == [c["coef"] for c in rec["cols"]]
    # glp_get_col_kind reports GLP_BV (3) for an integer column with bounds [0, 1]; the stored kind is GLP_IV
    assert d["kind"].tolist() == [2 if c["kind"] == 3 else c["kind"] for c in rec["cols"]]
    for j in range(n):
        a, b = d["A_ptr"][j], d["A_ptr"][j + 1]
        got = [[int(i) + 1, float(v)] for i, v in zip(d["A_ind"][a:b], d["A_val"][a:b])]
        assert got == rec["columns"][j], (j, got, rec["columns"][j])
    # the facade goes through the same reader
    P = glpk.glp_create_prob()
    assert glpk.glp_read_lp_from_string(P, None, rec["text"]) == 0
    assert [P.row[i].name for i in range(1, P.m + 1)] == names["rows"]


# ---- glp_scale_prob / glp_adv_basis against THE REFERENCE'S OWN runs (tests/golden/ref_hostprep_cases.json:
# lib/glpscl.js and lib/glpini01.js executed by minijs, oracle/jsref/fuzz_hostprep.py --golden): matrices with up
# to twelve decades of dynamic range, every combination of scaling flags
with open(os.path.join(H.GOLDEN, "ref_hostprep_cases.json")) as _f:
    REF_PREP = {k: v for k, v in json.load(_f).items() if not k.startswith("_")}


@pytest.mark.parametrize("name", sorted(REF_PREP))
def test_scaling_and_crash_basis_match_the_reference_bit_for_bit(name):
    import test_presolve as T
    case = REF_PREP[name]
    Q = T.facade_problem(case["problem"])
    for j in range(1, Q.n + 1):             # statuses of a freshly built problem (rows basic, columns by type)
        c = Q.col[j]
        glpk._set_bnds(c, "", j, c.type, c.lb, c.ub)
    lines = []
    glpk.glp_set_print_func(lines.append)
    try:
        glpk.glp_scale_prob(Q, case["flags"])
        glpk.glp_adv_basis(Q, 0)
    finally:
        glpk.glp_set_print_func(None)
    assert [Q.row[i].rii for i in range(1, Q.m + 1)] == case["rii"]
    assert [Q.col[j].sjj for j in range(1, Q.n + 1)] == case["sjj"]
    assert [Q.row[i].stat for i in range(1, Q.m + 1)] == case["row_stat"]
    assert [Q.col[j].stat for j in range(1, Q.n + 1)] == case["col_stat"]
    assert lines == case["lines"]
