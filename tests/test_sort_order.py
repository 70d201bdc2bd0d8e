"""CPU test of the closed form behind k_sort_list (csrc/kernels.cuh): the order the reference's sort_tcol / sort_trow
swap loop leaves (lib/glpspx01.js:795-804, lib/glpspx02.js:780-789), restated here with the same two prefix counts the
kernel computes, against the literal loop -- exhaustively for short vectors, randomly for longer ones -- and against the
lists the reference itself built (tests/golden/ref_vectors.npz, captured from lib/glpspx0[12].js run by minijs)."""
import itertools
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def literal(sig_of_nonzero):
    """the reference's loop on a list of non-zeros given by their significance flags; returns list indices (1-based)"""
    n = len(sig_of_nonzero)
    ind = [0] + list(range(1, n + 1))
    nnz, num = n, 0
    while num < nnz:
        i = ind[nnz]
        if not sig_of_nonzero[i - 1]:
            nnz -= 1
        else:
            num += 1
            ind[nnz] = ind[num]
            ind[num] = i
    return ind[1:num + 1]


def closed_form(sig_of_nonzero):
    """what k_sort_list does: with Ns significant entries, a significant entry with list index j <= Ns - 1 lands at
    place j + 1; the others take the free places (1, and j + 1 for every insignificant j <= Ns - 1) in DESCENDING
    index order"""
    n = len(sig_of_nonzero)
    ns = sum(sig_of_nonzero)
    f = ns - 1
    place = {}
    freerank = {}                       # r-th free place, r >= 2
    back = {}                           # r -> list index
    sgi = 0
    for j in range(1, n + 1):
        sg = sig_of_nonzero[j - 1]
        sgi += 1 if sg else 0
        if sg:
            if j <= f:
                place[j] = j + 1
            else:
                back[ns - (sgi - 1)] = j
        elif j <= f:
            freerank[(j - sgi) + 1] = j + 1
    for r, j in back.items():
        place[j] = 1 if r == 1 else freerank[r]
    out = [0] * ns
    for j, p in place.items():
        out[p - 1] = j
    return out


def test_closed_form_equals_the_swap_loop_exhaustively_up_to_12_entries():
    for n in range(0, 13):
        for flags in itertools.product((False, True), repeat=n):
            assert closed_form(flags) == literal(flags), flags


def test_closed_form_equals_the_swap_loop_on_random_flags():
    rng = np.random.default_rng(5)
    for _ in range(300):
        n = int(rng.integers(13, 400))
        flags = list(rng.random(n) < rng.random())
        assert closed_form(flags) == literal(flags)


def test_closed_form_reproduces_the_references_own_lists():
    z = np.load(os.path.join(HERE, "golden", "ref_vectors.npz"))
    groups = sorted({k.rsplit("/", 1)[0] for k in z.files if "/p_chuzr_" in k or "/d_chuzc_" in k})
    assert len(groups) >= 80
    for g in groups:
        primal = "/p_chuzr_" in g
        vec = z[g + ("/tcol_vec" if primal else "/trow_vec")]
        ind = z[g + ("/tcol_ind" if primal else "/trow_ind")]
        num = int(z[g + ("/tcol_num" if primal else "/trow_num")])
        big = float(np.max(np.abs(vec[1:])))
        eps = (1e-10 if primal else 1e-7) * (1.0 + 0.01 * big)        # smcp.tol_piv / sic tol_bnd (lib/glpspx02.js:1851)
        nz = [i for i in range(1, len(vec)) if vec[i] != 0.0]
        order = closed_form([not (abs(vec[i]) < eps) for i in nz])
        assert [nz[j - 1] for j in order] == [int(x) for x in ind[1:num + 1]], g
