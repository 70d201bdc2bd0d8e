"""ORACLE (test infrastructure, never on the product path): plain-Python
restatement of the reference's problem scaling and triangular crash basis,
used to check ``glpb_scale_prob`` / ``glpb_adv_basis`` (csrc/hostprep.cpp).

Follows, statement by statement and with the reference's 1-based arrays:
  * ``glp_scale_prob``  lib/glpscl.js:1-225, ``round2n`` lib/glplib03.js:26-31,
    ``glp_unscale_prob`` lib/glpapi04.js:44-50
  * ``glp_adv_basis``   lib/glpini01.js:1-363 (``triang``, ``mat``, ``adv_basis``)

Parity unpinned by the reference (it ships no asserted outputs and cannot be
run here: no JS engine); pinned instead by the properties tested in
tests/test_hostprep.py (triangularity of the chosen basis, scale-factor
identities) and by the hand-checked 3x3 case there.

A problem is given as ``rows[i] = [(j, a_ij), ...]`` and ``cols[j] = [(i, a_ij),
...]`` (1-based, slot 0 unused) in the reference's LIST order, i.e. what
walking ``row.ptr -> r_next`` / ``col.ptr -> c_next`` yields.
"""
import math

GLP_FR, GLP_LO, GLP_UP, GLP_DB, GLP_FX = 1, 2, 3, 4, 5
GLP_BS, GLP_NL, GLP_NU, GLP_NF, GLP_NS = 1, 2, 3, 4, 5
GLP_SF_GM, GLP_SF_EQ, GLP_SF_2N, GLP_SF_SKIP, GLP_SF_AUTO = 0x01, 0x10, 0x20, 0x40, 0x80


def round2n(x):  # glplib03.js:26-31
    assert x > 0.0
    e = math.floor(math.log(x) / math.log(2)) + 1
    f = x / math.pow(2, e)
    return math.pow(2, e - 1 if f <= 0.75 else e)


def scale_prob(m, n, rows, cols, flags):
    """Returns (rii[1+m], sjj[1+n], report) -- report = list of (tag, min, max, ratio)."""
    rii = [1.0] * (1 + m)  # glp_unscale_prob
    sjj = [1.0] * (1 + n)

    def min_row_aij(i):  # glpscl.js:2-15 (scaled = 1 everywhere on this path)
        min_aij, first = 1.0, True
        for (j, v) in rows[i]:
            temp = abs(v)
            temp *= (rii[i] * sjj[j])
            if first or min_aij > temp:
                min_aij = temp
            first = False
        return min_aij

    def max_row_aij(i):  # :17-30
        max_aij, first = 1.0, True
        for (j, v) in rows[i]:
            temp = abs(v)
            temp *= (rii[i] * sjj[j])
            if first or max_aij < temp:
                max_aij = temp
            first = False
        return max_aij

    def min_col_aij(j):  # :32-45
        min_aij, first = 1.0, True
        for (i, v) in cols[j]:
            temp = abs(v)
            temp *= (rii[i] * sjj[j])
            if first or min_aij > temp:
                min_aij = temp
            first = False
        return min_aij

    def max_col_aij(j):  # :47-60
        max_aij, first = 1.0, True
        for (i, v) in cols[j]:
            temp = abs(v)
            temp *= (rii[i] * sjj[j])
            if first or max_aij < temp:
                max_aij = temp
            first = False
        return max_aij

    def min_mat_aij():  # :62-72
        min_aij = 1.0
        for i in range(1, m + 1):
            temp = min_row_aij(i)
            if i == 1 or min_aij > temp:
                min_aij = temp
        return min_aij

    def max_mat_aij():  # :74-84
        max_aij = 1.0
        for i in range(1, m + 1):
            temp = max_row_aij(i)
            if i == 1 or max_aij < temp:
                max_aij = temp
        return max_aij

    def eq_scaling(flag):  # :86-103
        for pas in (0, 1):
            if pas == flag:
                for i in range(1, m + 1):
                    rii[i] = rii[i] / max_row_aij(i)
            else:
                for j in range(1, n + 1):
                    sjj[j] = sjj[j] / max_col_aij(j)

    def gm_scaling(flag):  # :105-124
        for pas in (0, 1):
            if pas == flag:
                for i in range(1, m + 1):
                    temp = min_row_aij(i) * max_row_aij(i)
                    rii[i] = rii[i] / math.sqrt(temp)
            else:
                for j in range(1, n + 1):
                    temp = min_col_aij(j) * max_col_aij(j)
                    sjj[j] = sjj[j] / math.sqrt(temp)

    def max_row_ratio():  # :126-135
        ratio = 1.0
        for i in range(1, m + 1):
            temp = max_row_aij(i) / min_row_aij(i)
            if i == 1 or ratio < temp:
                ratio = temp
        return ratio

    def max_col_ratio():  # :137-146
        ratio = 1.0
        for j in range(1, n + 1):
            temp = max_col_aij(j) / min_col_aij(j)
            if j == 1 or ratio < temp:
                ratio = temp
        return ratio

    def gm_iterate(it_max, tau):  # :148-165
        ratio = 0.0
        flag = int(max_row_ratio() > max_col_ratio())
        for k in range(1, it_max + 1):
            r_old = ratio
            ratio = max_mat_aij() / min_mat_aij()
            if k > 1 and ratio > tau * r_old:
                break
            gm_scaling(flag)

    report = []

    def note(tag):
        lo, hi = min_mat_aij(), max_mat_aij()
        report.append((tag, lo, hi, hi / lo))
        return lo, hi

    # glp_scale_prob, :216-225
    if flags & ~(GLP_SF_GM | GLP_SF_EQ | GLP_SF_2N | GLP_SF_SKIP | GLP_SF_AUTO):
        raise ValueError("glp_scale_prob: flags = %d; invalid scaling options" % flags)
    if flags & GLP_SF_AUTO:
        flags = GLP_SF_GM | GLP_SF_EQ | GLP_SF_SKIP
    # scale_prob, :167-214
    lo, hi = note("A")
    if lo >= 0.10 and hi <= 10.0:
        if flags & GLP_SF_SKIP:
            report.append(("skipped",))
            return rii, sjj, report
    if flags & GLP_SF_GM:
        gm_iterate(15, 0.90)
        note("GM")
    if flags & GLP_SF_EQ:
        eq_scaling(int(max_row_ratio() > max_col_ratio()))
        note("EQ")
    if flags & GLP_SF_2N:
        for i in range(1, m + 1):
            rii[i] = round2n(rii[i])
        for j in range(1, n + 1):
            sjj[j] = round2n(sjj[j])
        note("2N")
    return rii, sjj, report


def triang(m, n, mat):
    """glpini01.js:2-225.  mat(k, ndx) fills ndx[1..len] with the pattern of row k
    (k > 0) or column -k (k < 0) and returns len.  Returns (size, rn, cn)."""
    ndx = [0] * (1 + max(m, n))
    rs_len = [0] * (1 + m)
    rs_head = [0] * (1 + n)
    rs_prev = [0] * (1 + m)
    rs_next = [0] * (1 + m)
    cs_prev = [0] * (1 + n)
    cs_next = [0] * (1 + n)
    rn = [0] * (1 + m)
    cn = [0] * (1 + n)
    size = 0
    head = [0] * (1 + m)  # the reference borrows rs_len for this
    for j in range(1, n + 1):
        ln = mat(-j, ndx)
        assert 0 <= ln <= m
        cs_prev[j] = head[ln]
        head[ln] = j
    cs_head = 0
    for ln in range(0, m + 1):
        j = head[ln]
        while j != 0:
            cs_next[j] = cs_head
            cs_head = j
            j = cs_prev[j]
    jj = 0
    j = cs_head
    while j != 0:
        cs_prev[j] = jj
        jj = j
        j = cs_next[j]
    for i in range(1, m + 1):
        rs_len[i] = ln = mat(+i, ndx)
        assert 0 <= ln <= n
        rs_prev[i] = 0
        rs_next[i] = rs_head[ln]
        if rs_next[i] != 0:
            rs_prev[rs_next[i]] = i
        rs_head[ln] = i
    k1, k2 = 1, n
    while k1 <= k2:
        i = rs_head[1]
        if i != 0:
            assert rs_len[i] == 1
            j = 0
            t = mat(+i, ndx)
            while t >= 1:
                jj = ndx[t]
                if cn[jj] == 0:
                    assert j == 0
                    j = jj
                t -= 1
            assert j != 0
            rn[i] = cn[j] = k1
            k1 += 1
            size += 1
        else:
            j = cs_head
            assert j != 0
            cn[j] = k2
            k2 -= 1
        if cs_prev[j] == 0:
            cs_head = cs_next[j]
        else:
            cs_next[cs_prev[j]] = cs_next[j]
        if cs_next[j] != 0:
            cs_prev[cs_next[j]] = cs_prev[j]
        t = mat(-j, ndx)
        while t >= 1:
            i = ndx[t]
            ln = rs_len[i]
            assert ln >= 1
            if rs_prev[i] == 0:
                rs_head[ln] = rs_next[i]
            else:
                rs_next[rs_prev[i]] = rs_next[i]
            if rs_next[i] != 0:
                rs_prev[rs_next[i]] = rs_prev[i]
            ln -= 1
            rs_len[i] = ln
            rs_prev[i] = 0
            rs_next[i] = rs_head[ln]
            if rs_next[i] != 0:
                rs_prev[rs_next[i]] = i
            rs_head[ln] = i
            t -= 1
    for i in range(1, m + 1):
        if rn[i] == 0:
            rn[i] = k1
            k1 += 1
    for j in range(1, n + 1):
        assert cn[j] != 0
    # the reference's own optional checks (:176-222): permutations, lower triangle
    rn_inv = [0] * (1 + m)
    for i in range(1, m + 1):
        assert 1 <= rn[i] <= m and rn_inv[rn[i]] == 0
        rn_inv[rn[i]] = i
    cn_inv = [0] * (1 + n)
    for j in range(1, n + 1):
        assert 1 <= cn[j] <= n and cn_inv[cn[j]] == 0
        cn_inv[cn[j]] = j
    for ii in range(1, size + 1):
        diag = 0
        i = rn_inv[ii]
        t = mat(+i, ndx)
        while t >= 1:
            j = ndx[t]
            jj = cn[j]
            if jj <= size:
                assert jj <= ii
            if jj == ii:
                assert not diag
                diag = 1
            t -= 1
        assert diag
    return size, rn, cn


def adv_basis(m, n, rows, cols, r_type, r_lb, r_ub, c_type, c_lb, c_ub):
    """glpini01.js:281-363.  All arrays 1-based.  Returns (stat[1+m+n], size);
    for m == 0 or n == 0 the reference falls back to glp_std_basis (:358-361),
    which the caller handles."""
    assert m > 0 and n > 0

    def typx(k):
        return r_type[k] if k <= m else c_type[k - m]

    def mat(k, ndx):  # :227-279
        ln = 0
        if k > 0:
            i = k
            for (j, _) in rows[i]:
                if c_type[j] != GLP_FX:
                    ln += 1
                    ndx[ln] = m + j
            if r_type[i] != GLP_FX:
                ln += 1
                ndx[ln] = i
        else:
            j = -k
            if typx(j) != GLP_FX:
                if j <= m:
                    ln += 1
                    ndx[ln] = j
                else:
                    for (i, _) in cols[j - m]:
                        ln += 1
                        ndx[ln] = i
        return ln

    size, rn, cn = triang(m, m + n, mat)
    rn_inv = [0] * (1 + m)
    cn_inv = [0] * (1 + m + n)
    for i in range(1, m + 1):
        rn_inv[rn[i]] = i
    for j in range(1, m + n + 1):
        cn_inv[cn[j]] = j
    tagx = [-1] * (1 + m + n)
    for jj in range(1, size + 1):
        tagx[cn_inv[jj]] = GLP_BS
    for jj in range(size + 1, m + 1):
        i = rn_inv[jj]
        assert 1 <= i <= m and cn[i] > size
        tagx[i] = GLP_BS
    for k in range(1, m + n + 1):
        if tagx[k] != GLP_BS:
            t = typx(k)
            lb = r_lb[k] if k <= m else c_lb[k - m]
            ub = r_ub[k] if k <= m else c_ub[k - m]
            tagx[k] = {GLP_FR: GLP_NF, GLP_LO: GLP_NL, GLP_UP: GLP_NU,
                       GLP_DB: (GLP_NL if abs(lb) <= abs(ub) else GLP_NU), GLP_FX: GLP_NS}[t]
    return tagx, size
