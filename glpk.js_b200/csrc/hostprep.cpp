/* hostprep.cpp -- host-side callers on the input side of the simplex path
 * (SURVEY.md 8f rank 3): problem scaling and the triangular crash basis.
 *
 * Both produce INPUTS of the hot path -- the scale factors rii/sjj that
 * glpb_create applies to A, and the initial statuses glpb_set_basis takes --
 * so a host binding can reproduce the reference's
 *     glp_scale_prob(lp, flags); glp_adv_basis(lp, 0); glp_simplex(lp, parm)
 * flow (lib/glpapi06.js:108-128, lib/glpapi09.js:204-212, test/test.js:76)
 * without keeping its own copy of the matrix in linked lists.  O(nnz) symbolic
 * / min-max work, once per solve: it stays on the host by design (there is
 * nothing for the GPU to win here).
 *
 * Algorithms follow lib/glpscl.js:1-225, lib/glplib03.js:26-31 (round2n) and
 * lib/glpini01.js:1-363, written against flat 0-based CSC/CSR arrays.
 */
#include "../../include/glpb200.h"
#include <cmath>
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

namespace {

/* ---------------------------------------------------------------- scaling */

struct ScaleWork {
    int m, n;
    const int *cp, *ci;      /* columns: ptr[n+1], row index            */
    const double *cv;
    std::vector<int> rp, rj; /* rows:    ptr[m+1], column index          */
    std::vector<double> rv;  /* |a| by rows                              */
    std::vector<double> ca;  /* |a| by columns                           */
    double *rii, *sjj;

    void build_rows()
    {
        const int nnz = cp[n];
        rp.assign(m + 1, 0);
        for (int e = 0; e < nnz; e++) rp[ci[e] + 1]++;
        for (int i = 0; i < m; i++) rp[i + 1] += rp[i];
        rj.resize(nnz);
        rv.resize(nnz);
        ca.resize(nnz);
        std::vector<int> fill(rp.begin(), rp.end() - 1);
        for (int j = 0; j < n; j++)
            for (int e = cp[j]; e < cp[j + 1]; e++) {
                double a = std::fabs(cv[e]);
                ca[e] = a;
                int pos = fill[ci[e]]++;
                rj[pos] = j;
                rv[pos] = a;
            }
    }
    /* scaled magnitude exactly as the reference forms it:
     * temp = |a|; temp *= (rii * sjj)   (glpscl.js:9-10) */
    void row_minmax(int i, double &lo, double &hi) const
    {
        lo = hi = 1.0; /* value for an empty row (glpscl.js:6,20) */
        const double r = rii[i];
        for (int e = rp[i]; e < rp[i + 1]; e++) {
            double t = rv[e] * (r * sjj[rj[e]]);
            if (e == rp[i]) lo = hi = t;
            else {
                if (lo > t) lo = t;
                if (hi < t) hi = t;
            }
        }
    }
    void col_minmax(int j, double &lo, double &hi) const
    {
        lo = hi = 1.0;
        const double s = sjj[j];
        for (int e = cp[j]; e < cp[j + 1]; e++) {
            double t = ca[e] * (rii[ci[e]] * s);
            if (e == cp[j]) lo = hi = t;
            else {
                if (lo > t) lo = t;
                if (hi < t) hi = t;
            }
        }
    }
    /* min_mat_aij / max_mat_aij (glpscl.js:60-84): over ROWS, an empty row
     * counting as 1.0, and 1.0 for m == 0 */
    void mat_minmax(double &lo, double &hi) const
    {
        lo = hi = 1.0;
        for (int i = 0; i < m; i++) {
            double a, b;
            row_minmax(i, a, b);
            if (i == 0 || lo > a) lo = a;
            if (i == 0 || hi < b) hi = b;
        }
    }
    double max_row_ratio() const /* glpscl.js:126-135 */
    {
        double ratio = 1.0;
        for (int i = 0; i < m; i++) {
            double a, b;
            row_minmax(i, a, b);
            double t = b / a;
            if (i == 0 || ratio < t) ratio = t;
        }
        return ratio;
    }
    double max_col_ratio() const /* glpscl.js:137-146 */
    {
        double ratio = 1.0;
        for (int j = 0; j < n; j++) {
            double a, b;
            col_minmax(j, a, b);
            double t = b / a;
            if (j == 0 || ratio < t) ratio = t;
        }
        return ratio;
    }
    /* one sweep over the rows or over the columns; geometric = divide by
     * sqrt(min*max) (glpscl.js:105-124), else by max (glpscl.js:86-103).
     * A row sweep only changes rii and row i only reads rii[i], so the
     * sequential loop of the reference has no carried dependence. */
    void sweep_rows(bool geometric)
    {
        for (int i = 0; i < m; i++) {
            double a, b;
            row_minmax(i, a, b);
            rii[i] = geometric ? rii[i] / std::sqrt(a * b) : rii[i] / b;
        }
    }
    void sweep_cols(bool geometric)
    {
        for (int j = 0; j < n; j++) {
            double a, b;
            col_minmax(j, a, b);
            sjj[j] = geometric ? sjj[j] / std::sqrt(a * b) : sjj[j] / b;
        }
    }
    /* `flag` = rows are scaled worse than columns: then the columns go
     * first (pass == flag selects the rows; glpscl.js:90-91,109-110) */
    void two_pass(bool geometric, bool flag)
    {
        if (flag) { sweep_cols(geometric); sweep_rows(geometric); }
        else      { sweep_rows(geometric); sweep_cols(geometric); }
    }
};

/* lib/glplib03.js:26-31 -- the JS port takes the exponent from a quotient of
 * logarithms (not frexp); restated the same way so that a value sitting on a
 * rounding edge of log() falls the same side */
double round2n(double x)
{
    double e = std::floor(std::log(x) / std::log(2.0)) + 1.0;
    double f = x / std::pow(2.0, e);
    return std::pow(2.0, f <= 0.75 ? e - 1.0 : e);
}

/* ------------------------------------------------------------ crash basis */

/* triang (lib/glpini01.js:2-225) on the augmented matrix (I | -A) with the
 * columns of fixed variables emptied (mat, :227-279).  Rows live in buckets
 * by active length, every bucket a LIFO list; columns in one list by
 * descending initial length, ascending index inside a length.  The order in
 * which rows re-enter the buckets follows the column patterns read
 * backwards, which is why the caller hands the columns in list order. */
struct Triang {
    int m, N; /* N = m + n columns */
    const int *cp, *ci, *rp, *rj;
    const int *type; /* [m+n] GLP_* types */
    std::vector<int> rlen, rhead, rprev, rnext, cprev, cnext;
    int chead = -1;

    bool fixed(int k) const { return type[k] == 5 /* GLP_FX */; }

    template <class F> void for_col_backwards(int k, F f) const
    {
        if (fixed(k)) return;
        if (k < m) { f(k); return; }
        int j = k - m;
        for (int e = cp[j + 1] - 1; e >= cp[j]; e--) f(ci[e]);
    }
    int col_len(int k) const
    {
        if (fixed(k)) return 0;
        return k < m ? 1 : cp[k - m + 1] - cp[k - m];
    }
    int row_len(int i) const
    {
        int len = 0;
        for (int e = rp[i]; e < rp[i + 1]; e++)
            if (!fixed(m + rj[e])) len++;
        if (!fixed(i)) len++;
        return len;
    }
    void bucket_push(int i, int len)
    {
        rprev[i] = -1;
        rnext[i] = rhead[len];
        if (rnext[i] >= 0) rprev[rnext[i]] = i;
        rhead[len] = i;
    }
    void bucket_drop(int i, int len)
    {
        if (rprev[i] < 0) rhead[len] = rnext[i];
        else rnext[rprev[i]] = rnext[i];
        if (rnext[i] >= 0) rprev[rnext[i]] = rprev[i];
    }
    /* rn[m], cn[N]: 1-based positions in P*A~*Q as in the reference */
    int run(std::vector<int> &rn, std::vector<int> &cn)
    {
        rlen.assign(m, 0);
        rhead.assign(N + 1, -1);
        rprev.assign(m, -1);
        rnext.assign(m, -1);
        cprev.assign(N, -1);
        cnext.assign(N, -1);
        /* columns: counting sort by length, longest first, index ascending
         * within a length (glpini01.js:52-75) */
        {
            std::vector<int> cnt(m + 2, 0);
            for (int k = 0; k < N; k++) cnt[col_len(k)]++;
            std::vector<int> start(m + 2, 0);
            int pos = 0;
            for (int len = m; len >= 0; len--) { start[len] = pos; pos += cnt[len]; }
            std::vector<int> order(N);
            for (int k = 0; k < N; k++) order[start[col_len(k)]++] = k;
            chead = N ? order[0] : -1;
            for (int t = 0; t < N; t++) {
                cprev[order[t]] = t ? order[t - 1] : -1;
                cnext[order[t]] = t + 1 < N ? order[t + 1] : -1;
            }
        }
        for (int i = 0; i < m; i++) bucket_push(i, rlen[i] = row_len(i)); /* :78-88 */
        rn.assign(m, 0);
        cn.assign(N, 0);
        int k1 = 1, k2 = N, size = 0;
        while (k1 <= k2) {
            int j, i = rhead[1];
            if (i >= 0) {
                /* row singleton: its one active column (glpini01.js:97-121) */
                j = -1;
                if (!fixed(i) && cn[i] == 0) j = i;
                for (int e = rp[i]; e < rp[i + 1] && j < 0; e++) {
                    int k = m + rj[e];
                    if (!fixed(k) && cn[k] == 0) j = k;
                }
                if (j < 0) return -1; /* cannot happen */
                rn[i] = cn[j] = k1++;
                size++;
            } else {
                j = chead; /* longest remaining column leaves (:123-132) */
                if (j < 0) return -1;
                cn[j] = k2--;
            }
            if (cprev[j] < 0) chead = cnext[j];
            else cnext[cprev[j]] = cnext[j];
            if (cnext[j] >= 0) cprev[cnext[j]] = cprev[j];
            for_col_backwards(j, [&](int r) { /* :143-166 */
                int len = rlen[r];
                bucket_drop(r, len);
                bucket_push(r, rlen[r] = len - 1);
            });
        }
        for (int i = 0; i < m; i++)
            if (rn[i] == 0) rn[i] = k1++; /* :169 */
        return size;
    }
};

} // namespace

extern "C" {

/* replaces glp_scale_prob (lib/glpscl.js:216-225) */
int glpb_scale_prob(int m, int n, const int *A_ptr, const int *A_ind, const double *A_val,
                    int flags, double *rii, double *sjj, double *report)
try {
    if (m < 0 || n < 0 || !rii || !sjj || (n > 0 && !A_ptr)) return GLPB_EINVAL;
    if (flags & ~(GLPB_SF_GM | GLPB_SF_EQ | GLPB_SF_2N | GLPB_SF_SKIP | GLPB_SF_AUTO)) return GLPB_EINVAL;
    static const int zero_ptr[1] = {0};
    const int nnz = n ? A_ptr[n] : 0;
    for (int e = 0; e < nnz; e++)
        if (A_ind[e] < 0 || A_ind[e] >= m) return GLPB_EINVAL;
    if (flags & GLPB_SF_AUTO) flags = GLPB_SF_GM | GLPB_SF_EQ | GLPB_SF_SKIP;
    ScaleWork w;
    w.m = m; w.n = n;
    w.cp = n ? A_ptr : zero_ptr; w.ci = A_ind; w.cv = A_val;
    w.rii = rii; w.sjj = sjj;
    w.build_rows();
    double rep[13];
    for (double &x : rep) x = 0.0;
    auto note = [&](int slot) {
        double lo, hi;
        w.mat_minmax(lo, hi);
        rep[1 + 3 * slot] = lo; rep[2 + 3 * slot] = hi; rep[3 + 3 * slot] = hi / lo;
        rep[0] += (double)(1 << slot);
    };
    auto done = [&]() { if (report) std::memcpy(report, rep, sizeof rep); return 0; };
    for (int i = 0; i < m; i++) rii[i] = 1.0; /* glp_unscale_prob, glpscl.js:176 */
    for (int j = 0; j < n; j++) sjj[j] = 1.0;
    note(0);
    if (rep[1] >= 0.10 && rep[2] <= 10.0 && (flags & GLPB_SF_SKIP)) { /* :183-187 */
        rep[0] += 16.0;
        return done();
    }
    if (flags & GLPB_SF_GM) { /* gm_iterate(lp, 15, 0.90), glpscl.js:148-165 */
        bool flag = w.max_row_ratio() > w.max_col_ratio();
        double ratio = 0.0;
        for (int k = 1; k <= 15; k++) {
            double r_old = ratio, lo, hi;
            w.mat_minmax(lo, hi);
            ratio = hi / lo;
            if (k > 1 && ratio > 0.90 * r_old) break;
            w.two_pass(true, flag);
        }
        note(1);
    }
    if (flags & GLPB_SF_EQ) { /* :197-203 */
        w.two_pass(false, w.max_row_ratio() > w.max_col_ratio());
        note(2);
    }
    if (flags & GLPB_SF_2N) { /* :205-213 */
        for (int i = 0; i < m; i++) rii[i] = round2n(rii[i]);
        for (int j = 0; j < n; j++) sjj[j] = round2n(sjj[j]);
        note(3);
    }
    return done();
} catch (const std::bad_alloc &) {
    return GLPB_ENOMEM;   /* nothing crosses the C ABI as an exception */
}

/* replaces glp_adv_basis(lp, 0) (lib/glpini01.js:281-363) */
int glpb_adv_basis(int m, int n, const int *A_ptr, const int *A_ind, const int *R_ptr,
                   const int *R_ind, const int *type, const double *lb, const double *ub,
                   int *stat, int *tri_size)
try {
    if (m <= 0 || n <= 0 || !A_ptr || !R_ptr || !type || !lb || !ub || !stat) return GLPB_EINVAL;
    if (A_ptr[n] != R_ptr[m]) return GLPB_EINVAL;
    for (int e = 0; e < A_ptr[n]; e++)
        if (A_ind[e] < 0 || A_ind[e] >= m || R_ind[e] < 0 || R_ind[e] >= n) return GLPB_EINVAL;
    for (int k = 0; k < m + n; k++)
        if (type[k] < 1 || type[k] > 5) return GLPB_EINVAL;
    Triang t;
    t.m = m; t.N = m + n;
    t.cp = A_ptr; t.ci = A_ind; t.rp = R_ptr; t.rj = R_ind; t.type = type;
    std::vector<int> rn, cn;
    int size = t.run(rn, cn);
    if (size < 0) return GLPB_ESTATE;
    if (tri_size) *tri_size = size;
    std::vector<char> basic(m + n, 0);
    for (int k = 0; k < m + n; k++)
        if (cn[k] <= size) basic[k] = 1; /* glpini01.js:310-314 */
    for (int i = 0; i < m; i++)
        if (rn[i] > size) basic[i] = 1;  /* rows outside the triangle keep their slack, :317-330 */
    for (int k = 0; k < m + n; k++) {
        if (basic[k]) { stat[k] = 1; continue; } /* GLP_BS */
        switch (type[k]) {                          /* :332-352 */
        case 1: stat[k] = 4; break;                 /* FR -> NF */
        case 2: stat[k] = 2; break;                 /* LO -> NL */
        case 3: stat[k] = 3; break;                 /* UP -> NU */
        case 4: stat[k] = std::fabs(lb[k]) <= std::fabs(ub[k]) ? 2 : 3; break;
        default: stat[k] = 5; break;                /* FX -> NS */
        }
    }
    return 0;
} catch (const std::bad_alloc &) {
    return GLPB_ENOMEM;   /* nothing crosses the C ABI as an exception */
}

} // extern "C"
