/* gen.cpp -- CPU ORACLE (test infrastructure only; see glpo.h).
 * The oracle's own builder of the synthetic benchmark problems (SURVEY.md
 * 8d: C2 packing LP, C3 covering LP, C5 multi-dimensional knapsack), written
 * against the specification and the oracle's restatement of the reference's
 * random stream (lib/glprng01.js:17-52, lib/glprng02.js:1-5), so that
 * `bench.py --impl reference`, the cpu_baseline leg and tests/golden scripts
 * get their inputs without loading the product library.  tests/ check that
 * the product's generator (csrc/gen.cpp) yields the same arrays.
 * Output layout = glpo_load's input: 0-based CSC, rows ascending per column.
 */
#include "glpo.h"
#include <algorithm>
#include <cmath>
#include <cstring>

using namespace glpo;

#define API extern "C" __attribute__((visibility("default")))

namespace {
struct Built {
    int m = 0, n = 0, dir = GLP_MIN;
    std::vector<int> r_type, c_type, c_kind, A_ptr, A_ind;
    std::vector<double> r_lb, r_ub, c_lb, c_ub, c_coef, A_val;
};
Built g_last;

void shape(Built &b, int m, int n, int dir)
{
    b = Built();
    b.m = m; b.n = n; b.dir = dir;
    b.r_type.assign(m, GLP_FR); b.r_lb.assign(m, 0.0); b.r_ub.assign(m, 0.0);
    b.c_type.assign(n, GLP_LO); b.c_lb.assign(n, 0.0); b.c_ub.assign(n, 0.0);
    b.c_kind.assign(n, GLP_CV); b.c_coef.assign(n, 0.0);
    b.A_ptr.assign(n + 1, 0);
}
}

/* C2: max c'x, Ax <= b, x >= 0.  Column-major sweep; a cell is
   kept when its uniform draw is below `density`, its value is the next draw. */
API int glpo_gen_packing(int m, int n, double density, int seed)
{
    RNG r; rng_init(r, seed);
    Built &b = g_last; shape(b, m, n, GLP_MAX);
    std::vector<double> act(m, 0.0);
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < m; i++) {
            if (!(rng_unif_01(r) < density)) continue;
            double a = 0.1 + 0.9 * rng_unif_01(r);
            b.A_ind.push_back(i); b.A_val.push_back(a);
            act[i] += a;
        }
        b.A_ptr[j + 1] = (int)b.A_ind.size();
    }
    for (int j = 0; j < n; j++) b.c_coef[j] = 1.0 + 99.0 * rng_unif_01(r);
    for (int i = 0; i < m; i++) { b.r_type[i] = GLP_UP; b.r_ub[i] = 0.25 * act[i]; }
    return (int)b.A_ind.size();
}

/* C3: min c'x, Ax >= 1, x >= 0, c > 0.  Column j draws
   kmin + unif_rand(kspan) distinct rows by rejection; uncovered rows are
   appended to the first column. */
API int glpo_gen_covering(int m, int n, int kmin, int kspan, int seed)
{
    RNG r; rng_init(r, seed);
    Built &b = g_last; shape(b, m, n, GLP_MIN);
    std::vector<std::vector<Elem>> col(n);
    std::vector<int> owner(m, -1);
    std::vector<char> covered(m, 0);
    for (int j = 0; j < n; j++) {
        int want = kmin + rng_unif_rand(r, kspan);
        int have = 0;
        while (have < want) {
            int i = rng_unif_rand(r, m);
            if (owner[i] == j) continue;
            owner[i] = j;
            double a = 0.1 + 0.9 * rng_unif_01(r);
            col[j].push_back(Elem{i, a});
            covered[i] = 1;
            have++;
        }
    }
    for (int i = 0; i < m; i++)
        if (!covered[i]) col[0].push_back(Elem{i, 0.1 + 0.9 * rng_unif_01(r)});
    for (int j = 0; j < n; j++) {
        std::stable_sort(col[j].begin(), col[j].end(),
                         [](const Elem &x, const Elem &y) { return x.idx < y.idx; });
        for (const Elem &e : col[j]) { b.A_ind.push_back(e.idx); b.A_val.push_back(e.val); }
        b.A_ptr[j + 1] = (int)b.A_ind.size();
    }
    for (int j = 0; j < n; j++) b.c_coef[j] = 1.0 + 9.0 * rng_unif_01(r);
    for (int i = 0; i < m; i++) { b.r_type[i] = GLP_LO; b.r_lb[i] = 1.0; }
    return (int)b.A_ind.size();
}

/* C5: max p'x, Wx <= cap, x binary; weights row by row */
API int glpo_gen_mkp(int m, int n, int seed)
{
    RNG r; rng_init(r, seed);
    Built &b = g_last; shape(b, m, n, GLP_MAX);
    std::vector<double> W((size_t)m * n), colsum(n, 0.0), rowsum(m, 0.0);
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) {
            double w = 1.0 + rng_unif_rand(r, 1000);
            W[(size_t)i * n + j] = w;
        }
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < m; i++) colsum[j] += W[(size_t)i * n + j];
        b.c_coef[j] = std::floor(colsum[j] / m + 0.5) + rng_unif_rand(r, 500);
    }
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) rowsum[i] += W[(size_t)i * n + j];
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < m; i++) { b.A_ind.push_back(i); b.A_val.push_back(W[(size_t)i * n + j]); }
        b.A_ptr[j + 1] = (int)b.A_ind.size();
        b.c_type[j] = GLP_DB; b.c_ub[j] = 1.0; b.c_kind[j] = GLP_IV;
    }
    for (int i = 0; i < m; i++) { b.r_type[i] = GLP_UP; b.r_ub[i] = std::floor(0.5 * rowsum[i]); }
    return (int)b.A_ind.size();
}

/* copy the problem built by the last glpo_gen_* call (arrays sized by the
   caller from m, n and the returned nnz) */
API void glpo_gen_fetch(int *dir, int *r_type, double *r_lb, double *r_ub, int *c_type,
                        double *c_lb, double *c_ub, double *c_coef, int *c_kind,
                        int *A_ptr, int *A_ind, double *A_val)
{
    const Built &b = g_last;
    *dir = b.dir;
    auto cp = [](auto *dst, const auto &v) { if (!v.empty()) memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
    cp(r_type, b.r_type); cp(r_lb, b.r_lb); cp(r_ub, b.r_ub);
    cp(c_type, b.c_type); cp(c_lb, b.c_lb); cp(c_ub, b.c_ub); cp(c_coef, b.c_coef); cp(c_kind, b.c_kind);
    cp(A_ptr, b.A_ptr); cp(A_ind, b.A_ind); cp(A_val, b.A_val);
}
