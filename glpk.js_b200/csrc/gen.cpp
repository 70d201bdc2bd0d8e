/* gen.cpp -- synthetic LP/MIP generators for the benchmark configurations
 * (C2 packing LP, C3 covering LP, C5 multi-dimensional knapsack).
 * The random stream is Knuth's subtractive generator as restated from
 * lib/glprng01.js:1-60 and lib/glprng02.js, integer-only and therefore
 * reproducible bit for bit in any language.
 */
#include "../../include/glpb200.h"
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

struct Rng {
    int A[56];
    int fptr;
    static int md(int x, int y) { return (int)(((unsigned)x - (unsigned)y) & 0x7FFFFFFFu); }
    int flip()
    {
        int ii, jj;
        for (ii = 1, jj = 32; jj <= 55; ii++, jj++) A[ii] = md(A[ii], A[jj]);
        for (jj = 1; ii <= 55; ii++, jj++) A[ii] = md(A[ii], A[jj]);
        fptr = 54;
        return A[55];
    }
    explicit Rng(int seed)
    {
        A[0] = -1;
        for (int i = 1; i <= 55; i++) A[i] = 0;
        fptr = 0;
        int prev = seed, next = 1;
        seed = prev = md(prev, 0);
        A[55] = prev;
        for (int i = 21; i; i = (i + 21) % 55) {
            A[i] = next;
            next = md(prev, next);
            if (seed & 1) seed = 0x40000000 + (seed >> 1); else seed >>= 1;
            next = md(next, seed);
            prev = A[i];
        }
        for (int t = 0; t < 5; t++) flip();
    }
    int next() { return A[fptr] >= 0 ? A[fptr--] : flip(); }
    int unif(int m)
    {
        const unsigned two31 = 0x80000000u;
        unsigned t = two31 - (two31 % (unsigned)m);
        int r;
        do { r = next(); } while (t <= (unsigned)r);
        return r % m;
    }
    double u01() { return (double)next() / 2147483647.0; }
};

enum { GLP_MIN = 1, GLP_MAX = 2, GLP_CV = 1, GLP_IV = 2,
       GLP_FR = 1, GLP_LO = 2, GLP_UP = 3, GLP_DB = 4, GLP_FX = 5 };

template <class T> T *dup(const std::vector<T> &v)
{
    T *p = (T *)malloc((v.size() ? v.size() : 1) * sizeof(T));
    if (v.size()) memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}

void finish(glpb_problem_data *out, int m, int n, int dir,
            const std::vector<int> &type, const std::vector<double> &lb,
            const std::vector<double> &ub, const std::vector<double> &coef,
            const std::vector<int> &kind, const std::vector<int> &ptr,
            const std::vector<int> &ind, const std::vector<double> &val)
{
    out->m = m; out->n = n; out->nnz = (int)ind.size(); out->dir = dir; out->c0 = 0.0;
    out->type = dup(type); out->lb = dup(lb); out->ub = dup(ub);
    out->coef = dup(coef); out->kind = dup(kind);
    out->A_ptr = dup(ptr); out->A_ind = dup(ind); out->A_val = dup(val);
}

} /* namespace */

extern "C" {

void glpb_rng_fill(int seed, int count, int *out)
{
    Rng r(seed);
    for (int i = 0; i < count; i++) out[i] = r.next();
}

/* C2: maximize c'x, Ax <= b, x >= 0; slack basis primal feasible */
int glpb_gen_packing(int m, int n, double density, int seed, glpb_problem_data *out)
{
    if (m < 1 || n < 1 || !out) return GLPB_EINVAL;
    Rng r(seed);
    std::vector<int> ptr(n + 1, 0), ind;
    std::vector<double> val, rowsum(m, 0.0);
    for (int j = 0; j < n; j++) {
        ptr[j] = (int)ind.size();
        for (int i = 0; i < m; i++) {
            double u = r.u01();
            if (u < density) {
                double a = 0.1 + 0.9 * r.u01();
                ind.push_back(i); val.push_back(a);
                rowsum[i] += a;
            }
        }
    }
    ptr[n] = (int)ind.size();
    std::vector<double> coef(n);
    for (int j = 0; j < n; j++) coef[j] = 1.0 + 99.0 * r.u01();
    std::vector<int> type(m + n), kind(n, GLP_CV);
    std::vector<double> lb(m + n, 0.0), ub(m + n, 0.0);
    for (int i = 0; i < m; i++) { type[i] = GLP_UP; ub[i] = 0.25 * rowsum[i]; }
    for (int j = 0; j < n; j++) type[m + j] = GLP_LO;
    finish(out, m, n, GLP_MAX, type, lb, ub, coef, kind, ptr, ind, val);
    return 0;
}

/* C3: minimize c'x, Ax >= 1, x >= 0, c > 0; slack basis dual feasible.
 * Column j gets kmin + unif(kspan) distinct rows by rejection. */
int glpb_gen_covering(int m, int n, int kmin, int kspan, int seed, glpb_problem_data *out)
{
    if (m < 1 || n < 1 || kmin < 1 || kspan < 1 || kmin + kspan - 1 > m || !out) return GLPB_EINVAL;
    Rng r(seed);
    std::vector<std::vector<std::pair<int, double>>> cols(n);
    std::vector<int> mark(m, -1), rowcnt(m, 0);
    for (int j = 0; j < n; j++) {
        int kj = kmin + r.unif(kspan);
        while ((int)cols[j].size() < kj) {
            int i = r.unif(m);
            if (mark[i] == j) continue;
            mark[i] = j;
            cols[j].push_back({i, 0.1 + 0.9 * r.u01()});
            rowcnt[i]++;
        }
    }
    /* cover empty rows with the least-index column not containing them */
    for (int i = 0; i < m; i++) {
        if (rowcnt[i] > 0) continue;
        cols[0].push_back({i, 0.1 + 0.9 * r.u01()});
        rowcnt[i]++;
    }
    std::vector<int> ptr(n + 1, 0), ind;
    std::vector<double> val;
    for (int j = 0; j < n; j++) {
        ptr[j] = (int)ind.size();
        /* ascending row order = state after glp_sort_matrix */
        std::vector<std::pair<int, double>> &c = cols[j];
        for (size_t a = 1; a < c.size(); a++)
            for (size_t b = a; b > 0 && c[b - 1].first > c[b].first; b--) std::swap(c[b - 1], c[b]);
        for (auto &e : c) { ind.push_back(e.first); val.push_back(e.second); }
    }
    ptr[n] = (int)ind.size();
    std::vector<double> coef(n);
    for (int j = 0; j < n; j++) coef[j] = 1.0 + 9.0 * r.u01();
    std::vector<int> type(m + n), kind(n, GLP_CV);
    std::vector<double> lb(m + n, 0.0), ub(m + n, 0.0);
    for (int i = 0; i < m; i++) { type[i] = GLP_LO; lb[i] = 1.0; }
    for (int j = 0; j < n; j++) type[m + j] = GLP_LO;
    finish(out, m, n, GLP_MIN, type, lb, ub, coef, kind, ptr, ind, val);
    return 0;
}

/* C5: multi-dimensional knapsack, n binaries, m capacity rows (dense) */
int glpb_gen_mkp(int m, int n, int seed, glpb_problem_data *out)
{
    if (m < 1 || n < 1 || !out) return GLPB_EINVAL;
    Rng r(seed);
    std::vector<double> w((size_t)m * n);
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) w[(size_t)i * n + j] = 1.0 + r.unif(1000);
    std::vector<double> coef(n);
    for (int j = 0; j < n; j++) {
        double s = 0.0;
        for (int i = 0; i < m; i++) s += w[(size_t)i * n + j];
        coef[j] = std::floor(s / m + 0.5) + r.unif(500);
    }
    std::vector<int> ptr(n + 1, 0), ind;
    std::vector<double> val;
    for (int j = 0; j < n; j++) {
        ptr[j] = (int)ind.size();
        for (int i = 0; i < m; i++) { ind.push_back(i); val.push_back(w[(size_t)i * n + j]); }
    }
    ptr[n] = (int)ind.size();
    std::vector<int> type(m + n), kind(n, GLP_IV);
    std::vector<double> lb(m + n, 0.0), ub(m + n, 0.0);
    for (int i = 0; i < m; i++) {
        double s = 0.0;
        for (int j = 0; j < n; j++) s += w[(size_t)i * n + j];
        type[i] = GLP_UP; ub[i] = std::floor(0.5 * s);
    }
    for (int j = 0; j < n; j++) { type[m + j] = GLP_DB; lb[m + j] = 0.0; ub[m + j] = 1.0; }
    finish(out, m, n, GLP_MAX, type, lb, ub, coef, kind, ptr, ind, val);
    return 0;
}

void glpb_free_problem(glpb_problem_data *d)
{
    if (!d) return;
    free(d->type); free(d->lb); free(d->ub); free(d->coef); free(d->kind);
    free(d->A_ptr); free(d->A_ind); free(d->A_val);
    memset(d, 0, sizeof *d);
}

} /* extern "C" */
