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


def test_row_wise_pivot_row_is_the_column_wise_one_bit_for_bit():
    """lib/glpspx02.js:735-752 switches between eval_trow1 (column dots, :655-693) and eval_trow2 (row
    scatter for sparse rho, :695-733) on the density of rho.  With the columns of A stored in ascending row
    order -- which is how init_csa receives them after glp_sort_matrix / glp_read_lp -- both forms add the
    SAME products in the SAME order for every column (the row scatter merely skips the exact zeros), so the
    switch changes the reference's running time, not one bit of trow: the device, which always walks the
    columns, has nothing to reproduce.  Literal restatement of both loops on random data."""
    import numpy as np
    rng = np.random.default_rng(7)
    for trial in range(30):
        m, n = int(rng.integers(5, 40)), int(rng.integers(5, 60))
        A = np.where(rng.random((m, n)) < 0.3, rng.uniform(-3, 3, (m, n)), 0.0)
        rho = np.where(rng.random(m) < (0.1 if trial % 2 else 0.5), rng.uniform(-2, 2, m), 0.0)
        # a basis header: N = the first n of a random permutation of the m+n variables
        head = rng.permutation(m + n) + 1
        nonbasic = head[m:]                       # x[k] = xN[j], k = head[m+j]
        bind = np.zeros(m + n + 1, dtype=int)
        for pos, k in enumerate(head, 1):
            bind[k] = pos
        fixed = rng.random(n) < 0.1               # stat[j] == GLP_NS
        # eval_trow1
        t1 = np.zeros(n)
        for j in range(n):
            if fixed[j]:
                continue
            k = nonbasic[j]
            if k <= m:
                t1[j] = -rho[k - 1]
            else:
                temp = 0.0
                for i in range(m):                # column k-m, ascending rows
                    if A[i, k - m - 1] != 0.0:
                        temp += rho[i] * A[i, k - m - 1]
                t1[j] = temp
        # eval_trow2
        t2 = np.zeros(n)
        for i in range(m):
            temp = rho[i]
            if temp == 0.0:
                continue
            j = bind[i + 1] - m
            if j >= 1 and not fixed[j - 1]:
                t2[j - 1] -= temp
            for c in range(n):                    # row i, ascending columns
                if A[i, c] != 0.0:
                    j = bind[m + c + 1] - m
                    if j >= 1 and not fixed[j - 1]:
                        t2[j - 1] += temp * A[i, c]
        assert np.array_equal(t1, t2)             # == on every double (0.0 == -0.0: both leave trow_ind alone)
