/* spx_common.h -- CPU ORACLE (test infrastructure only; see glpo.h).
 * Routines that are textually duplicated between the reference's primal and
 * dual loops; restated once.  Citations give the primal file first.
 */
#ifndef GLPO_SPX_COMMON_H
#define GLPO_SPX_COMMON_H
#include "csa.h"
#include <cassert>
#include <cmath>
#include <cstring>

namespace glpo {

static const double kappa = 0.10; /* lib/glpspx01.js:3, lib/glpspx02.js:3 */

/* lib/glpspx01.js:147-175 / lib/glpspx02.js:192-220 inv_col: column i of B,
   i.e. column head[i] of (I | -A) */
static int spx_inv_col(void *info, int i, int *ind, double *val)
{
    CSA &c = *(CSA *)info;
    int k = c.head[i];
    if (k <= c.m) { ind[1] = k; val[1] = 1.0; return 1; }
    int ptr = c.A_ptr[k - c.m], len = c.A_ptr[k - c.m + 1] - ptr;
    for (int t = 1; t <= len; t++) { ind[t] = c.A_ind[ptr + t - 1]; val[t] = -c.A_val[ptr + t - 1]; }
    return len;
}

/* lib/glpspx01.js:177-181 invert_B */
static inline int spx_invert_B(CSA &c)
{
    int ret = bfd_factorize(*c.bfd, c.m, spx_inv_col, &c);
    c.valid = (ret == 0);
    return ret;
}

/* lib/glpspx01.js:183-217 update_B: column i of B becomes column k of (I|-A) */
static inline int spx_update_B(CSA &c, int i, int k)
{
    int ret;
    if (k <= c.m) {
        int ind[2] = {0, k};
        double val[2] = {0.0, 1.0};
        ret = bfd_update_it(*c.bfd, i, 1, ind, 0, val);
    } else {
        double *val = c.work1.data();
        int beg = c.A_ptr[k - c.m], end = c.A_ptr[k - c.m + 1], len = 0;
        for (int ptr = beg; ptr < end; ptr++) val[++len] = -c.A_val[ptr];
        ret = bfd_update_it(*c.bfd, i, len, c.A_ind.data(), beg - 1, val);
    }
    c.valid = (ret == 0);
    return ret;
}

/* lib/glpspx01.js:219-249 error_ftran: r = h - B*x */
static inline void spx_error_ftran(CSA &c, const double *h, const double *x, double *r)
{
    const int m = c.m;
    memmove(&r[1], &h[1], m * sizeof(double));
    for (int i = 1; i <= m; i++) {
        double temp = x[i];
        if (temp == 0.0) continue;
        int k = c.head[i];
        if (k <= m) r[k] -= temp;
        else {
            int beg = c.A_ptr[k - m], end = c.A_ptr[k - m + 1];
            for (int ptr = beg; ptr < end; ptr++) r[c.A_ind[ptr]] += c.A_val[ptr] * temp;
        }
    }
}

/* lib/glpspx01.js:251-263 refine_ftran (r and d alias work1) */
static inline void spx_refine_ftran(CSA &c, const double *h, double *x)
{
    double *r = c.work1.data();
    spx_error_ftran(c, h, x, r);
    bfd_ftran(*c.bfd, r);
    for (int i = 1; i <= c.m; i++) x[i] += r[i];
}

/* lib/glpspx01.js:265-293 error_btran: r = h - B'*x */
static inline void spx_error_btran(CSA &c, const double *h, const double *x, double *r)
{
    const int m = c.m;
    for (int i = 1; i <= m; i++) {
        int k = c.head[i];
        double temp = h[i];
        if (k <= m) temp -= x[k];
        else {
            int beg = c.A_ptr[k - m], end = c.A_ptr[k - m + 1];
            for (int ptr = beg; ptr < end; ptr++) temp += c.A_val[ptr] * x[c.A_ind[ptr]];
        }
        r[i] = temp;
    }
}

/* lib/glpspx01.js:295-307 refine_btran */
static inline void spx_refine_btran(CSA &c, const double *h, double *x)
{
    double *r = c.work1.data();
    spx_error_btran(c, h, x, r);
    bfd_btran(*c.bfd, r);
    for (int i = 1; i <= c.m; i++) x[i] += r[i];
}

/* lib/glpspx01.js:442-471 get_xN */
static inline double spx_get_xN(const CSA &c, int j)
{
    int k = c.head[c.m + j];
    switch (c.stat[j]) {
    case GLP_NL: return c.lb[k];
    case GLP_NU: return c.ub[k];
    case GLP_NF: return 0.0;
    case GLP_NS: return c.lb[k];
    default: assert(!"bad stat"); return 0.0;
    }
}

/* lib/glpspx01.js:473-512 eval_beta: beta = inv(B) * (-N xN), one refinement */
static inline void spx_eval_beta(CSA &c, double *beta)
{
    const int m = c.m, n = c.n;
    double *h = c.work2.data();
    for (int i = 1; i <= m; i++) h[i] = 0.0;
    for (int j = 1; j <= n; j++) {
        int k = c.head[m + j];
        double xN = spx_get_xN(c, j);
        if (xN == 0.0) continue;
        if (k <= m) h[k] -= xN;
        else {
            int beg = c.A_ptr[k - m], end = c.A_ptr[k - m + 1];
            for (int ptr = beg; ptr < end; ptr++) h[c.A_ind[ptr]] += xN * c.A_val[ptr];
        }
    }
    memmove(&beta[1], &h[1], m * sizeof(double));
    bfd_ftran(*c.bfd, beta);
    spx_refine_ftran(c, h, beta);
}

/* lib/glpspx01.js:514-529 eval_pi */
static inline void spx_eval_pi(CSA &c, double *pi)
{
    const int m = c.m;
    double *cB = c.work2.data();
    for (int i = 1; i <= m; i++) cB[i] = c.coef[c.head[i]];
    memmove(&pi[1], &cB[1], m * sizeof(double));
    bfd_btran(*c.bfd, pi);
    spx_refine_btran(c, cB, pi);
}

/* lib/glpspx01.js:531-558 eval_cost: d_j = c_k - N_j' pi */
static inline double spx_eval_cost(const CSA &c, const double *pi, int j)
{
    const int m = c.m;
    int k = c.head[m + j];
    double dj = c.coef[k];
    if (k <= m) dj -= pi[k];
    else {
        int beg = c.A_ptr[k - m], end = c.A_ptr[k - m + 1];
        for (int ptr = beg; ptr < end; ptr++) dj += c.A_val[ptr] * pi[c.A_ind[ptr]];
    }
    return dj;
}

/* lib/glpspx01.js:560-563 eval_bbar */
static inline void spx_eval_bbar(CSA &c) { spx_eval_beta(c, c.bbar.data()); }

/* lib/glpspx01.js:565-584 eval_cbar */
static inline void spx_eval_cbar(CSA &c)
{
    double *pi = c.work3.data();
    spx_eval_pi(c, pi);
    for (int j = 1; j <= c.n; j++) c.cbar[j] = spx_eval_cost(c, pi, j);
}

/* build h = -N[q] in a dense vector (lib/glpspx01.js:702-719) */
static inline void spx_neg_Nq(const CSA &c, int q, double *h)
{
    const int m = c.m;
    int k = c.head[m + q];
    for (int i = 1; i <= m; i++) h[i] = 0.0;
    if (k <= m) h[k] = -1.0;
    else {
        int beg = c.A_ptr[k - m], end = c.A_ptr[k - m + 1];
        for (int ptr = beg; ptr < end; ptr++) h[c.A_ind[ptr]] = c.A_val[ptr];
    }
}

static inline void spx_tcol_pattern(CSA &c)
{
    int nnz = 0;
    for (int i = 1; i <= c.m; i++)
        if (c.tcol_vec[i] != 0.0) c.tcol_ind[++nnz] = i;
    c.tcol_nnz = nnz;
}

/* lib/glpspx01.js:690-730 / lib/glpspx02.js:937-977 eval_tcol */
static inline void spx_eval_tcol(CSA &c)
{
    spx_neg_Nq(c, c.q, c.tcol_vec.data());
    bfd_ftran(*c.bfd, c.tcol_vec.data());
    spx_tcol_pattern(c);
}

/* lib/glpspx01.js:732-771 / lib/glpspx02.js:979-1018 refine_tcol */
static inline void spx_refine_tcol(CSA &c)
{
    double *h = c.work3.data();
    spx_neg_Nq(c, c.q, h);
    spx_refine_ftran(c, h, c.tcol_vec.data());
    spx_tcol_pattern(c);
}

/* lib/glpspx01.js:1524-1548 / lib/glpspx02.js:1424-1450 eval_obj */
static inline double spx_eval_obj(const CSA &c)
{
    const int m = c.m, n = c.n;
    double sum = c.obj[0];
    for (int i = 1; i <= m; i++) {
        int k = c.head[i];
        if (k > m) sum += c.obj[k - m] * c.bbar[i];
    }
    for (int j = 1; j <= n; j++) {
        int k = c.head[m + j];
        if (k > m) sum += c.obj[k - m] * spx_get_xN(c, j);
    }
    return sum;
}

/* the part of init_csa both loops share (lib/glpspx01.js:42-132,
   lib/glpspx02.js:89-180): scaled bounds/costs, zeta, CSC copy of A in list
   order, basis header with non-basic rows first then non-basic columns */
static inline void spx_init_common(CSA &c, Prob &lp, bool dual)
{
    const int m = c.m = lp.m, n = c.n = lp.n, nnz = lp.nnz;
    assert(m > 0 && n > 0);
    c.type.assign(1 + m + n, 0); c.lb.assign(1 + m + n, 0.0);
    c.ub.assign(1 + m + n, 0.0); c.coef.assign(1 + m + n, 0.0);
    c.obj.assign(1 + n, 0.0);
    c.A_ptr.assign(1 + n + 1, 0); c.A_ind.assign(1 + nnz, 0); c.A_val.assign(1 + nnz, 0.0);
    c.head.assign(1 + m + n, 0); c.stat.assign(1 + n, 0);
    c.bbar.assign(1 + m, 0.0); c.cbar.assign(1 + n, 0.0);
    c.refsp.assign(1 + m + n, 0);
    c.gamma.assign(1 + (dual ? m : n), 0.0);
    c.tcol_ind.assign(1 + m, 0); c.tcol_vec.assign(1 + m, 0.0);
    c.trow_ind.assign(1 + n, 0); c.trow_vec.assign(1 + n, 0.0);
    c.work1.assign(1 + m, 0.0); c.work2.assign(1 + m, 0.0);
    c.work3.assign(1 + m, 0.0); c.work4.assign(1 + m, 0.0);
    for (int i = 1; i <= m; i++) {
        c.type[i] = (signed char)lp.r_type[i];
        c.lb[i] = lp.r_lb[i] * lp.r_rii[i];
        c.ub[i] = lp.r_ub[i] * lp.r_rii[i];
        c.coef[i] = 0.0;
    }
    for (int j = 1; j <= n; j++) {
        c.type[m + j] = (signed char)lp.c_type[j];
        c.lb[m + j] = lp.c_lb[j] / lp.c_sjj[j];
        c.ub[m + j] = lp.c_ub[j] / lp.c_sjj[j];
        c.coef[m + j] = lp.c_coef[j] * lp.c_sjj[j];
    }
    c.obj[0] = lp.c0;
    for (int j = 1; j <= n; j++) c.obj[j] = c.coef[m + j];
    double cmax = 0.0;
    for (int j = 1; j <= n; j++)
        if (cmax < fabs(c.obj[j])) cmax = fabs(c.obj[j]);
    if (cmax == 0.0) cmax = 1.0;
    c.zeta = (lp.dir == GLP_MIN ? +1.0 : -1.0) / cmax;
    if (fabs(c.zeta) < 1.0) c.zeta *= 1000.0;
    int loc = 1;
    for (int j = 1; j <= n; j++) {
        c.A_ptr[j] = loc;
        for (const Elem &e : lp.col_list[j]) {
            c.A_ind[loc] = e.idx;
            c.A_val[loc] = lp.r_rii[e.idx] * e.val * lp.c_sjj[j];
            loc++;
        }
    }
    c.A_ptr[n + 1] = loc;
    assert(loc == nnz + 1);
    assert(lp.valid);
    for (int i = 1; i <= m; i++) c.head[i] = lp.head[i];
    int k = 0;
    for (int i = 1; i <= m; i++)
        if (lp.r_stat[i] != GLP_BS) { k++; c.head[m + k] = i; c.stat[k] = (signed char)lp.r_stat[i]; }
    for (int j = 1; j <= n; j++)
        if (lp.c_stat[j] != GLP_BS) { k++; c.head[m + k] = m + j; c.stat[k] = (signed char)lp.c_stat[j]; }
    assert(k == n);
    c.valid = 1; lp.valid = 0;
    c.bfd = lp.bfd; lp.bfd = nullptr;
    c.phase = 0;
    c.tm_beg = xtime_ms();
    c.it_beg = c.it_cnt = lp.it_cnt;
    c.it_dpy = -1;
    c.refct = 0;
}

/* lib/glpspx01.js:1591-1681 / lib/glpspx02.js:1499-1590 store_sol */
static inline void spx_store_sol(CSA &c, Prob &lp, int p_stat, int d_stat, int ray)
{
    const int m = c.m, n = c.n;
    lp.valid = 1; c.valid = 0;
    lp.bfd = c.bfd; c.bfd = nullptr;
    for (int i = 1; i <= m; i++) lp.head[i] = c.head[i];
    lp.pbs_stat = p_stat;
    lp.dbs_stat = d_stat;
    lp.obj_val = spx_eval_obj(c);
    lp.it_cnt = c.it_cnt;
    lp.some = ray;
    for (int i = 1; i <= m; i++) {
        int k = c.head[i];
        if (k <= m) {
            lp.r_stat[k] = GLP_BS; lp.r_bind[k] = i;
            lp.r_prim[k] = c.bbar[i] / lp.r_rii[k]; lp.r_dual[k] = 0.0;
        } else {
            int j = k - m;
            lp.c_stat[j] = GLP_BS; lp.c_bind[j] = i;
            lp.c_prim[j] = c.bbar[i] * lp.c_sjj[j]; lp.c_dual[j] = 0.0;
        }
    }
    for (int j = 1; j <= n; j++) {
        int k = c.head[m + j];
        if (k <= m) {
            lp.r_stat[k] = c.stat[j]; lp.r_bind[k] = 0;
            switch (c.stat[j]) {
            case GLP_NL: lp.r_prim[k] = lp.r_lb[k]; break;
            case GLP_NU: lp.r_prim[k] = lp.r_ub[k]; break;
            case GLP_NF: lp.r_prim[k] = 0.0; break;
            case GLP_NS: lp.r_prim[k] = lp.r_lb[k]; break;
            }
            lp.r_dual[k] = (c.cbar[j] * lp.r_rii[k]) / c.zeta;
        } else {
            int jj = k - m;
            lp.c_stat[jj] = c.stat[j]; lp.c_bind[jj] = 0;
            switch (c.stat[j]) {
            case GLP_NL: lp.c_prim[jj] = lp.c_lb[jj]; break;
            case GLP_NU: lp.c_prim[jj] = lp.c_ub[jj]; break;
            case GLP_NF: lp.c_prim[jj] = 0.0; break;
            case GLP_NS: lp.c_prim[jj] = lp.c_lb[jj]; break;
            }
            lp.c_dual[jj] = (c.cbar[j] / lp.c_sjj[jj]) / c.zeta;
        }
    }
}

/* give the factorisation back without a solution
   (lib/glpspx01.js:1715-1721, 1943-1949) */
static inline int spx_fail(CSA &c, Prob &lp)
{
    lp.bfd = c.bfd; c.bfd = nullptr;
    lp.pbs_stat = lp.dbs_stat = GLP_UNDEF;
    lp.obj_val = 0.0;
    lp.it_cnt = c.it_cnt;
    lp.some = 0;
    return GLP_EFAIL;
}

/* ---- stateless selection routines (also exported for kernel parity) ---- */
int o_chuzc_primal(int n, const signed char *stat, const double *cbar,
                   const double *gamma, double tol_dj);
void o_sort_list(int *ind, const double *vec, int nnz, double tol_piv,
                 int *num_out, double *max_out);
void o_chuzr_primal(int m, const signed char *type, const double *lb,
                    const double *ub, const double *coef, const int *head,
                    int phase, const double *bbar, double cbar_q, int q,
                    const int *tcol_ind, const double *tcol_vec, int tcol_num,
                    double rtol, int *p_out, int *p_stat_out, double *teta_out);
int o_chuzr_dual(int m, const signed char *type, const double *lb,
                 const double *ub, const int *head, const double *bbar,
                 const double *gamma, double tol_bnd, double *delta_out);
void o_chuzc_dual(const signed char *stat, const double *cbar, double delta,
                  const int *trow_ind, const double *trow_vec, int trow_num,
                  double rtol, int *q_out, double *new_dq_out);

} /* namespace glpo */
#endif
