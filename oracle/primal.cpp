/* primal.cpp -- CPU ORACLE (test infrastructure only; see glpo.h).
 * Restatement of the two-phase primal revised simplex, lib/glpspx01.js.
 */
#include "spx_common.h"

namespace glpo {

/* lib/glpspx01.js:646-688 chuzc: argmax d_j^2/gamma_j over eligible j;
   strict '<' keeps the lowest index on ties.  Arrays are 1-based. */
int o_chuzc_primal(int n, const signed char *stat, const double *cbar,
                   const double *gamma, double tol_dj)
{
    int q = 0;
    double best = 0.0;
    for (int j = 1; j <= n; j++) {
        double dj = cbar[j];
        switch (stat[j]) {
        case GLP_NL: if (dj >= -tol_dj) continue; break;
        case GLP_NU: if (dj <= +tol_dj) continue; break;
        case GLP_NF: if (-tol_dj <= dj && dj <= +tol_dj) continue; break;
        case GLP_NS: continue;
        default: assert(!"bad stat");
        }
        double temp = (dj * dj) / gamma[j];
        if (best < temp) { q = j; best = temp; }
    }
    return q;
}

/* lib/glpspx01.js:773-806 sort_tcol / lib/glpspx02.js:754-791 sort_trow:
   infinity norm, eps = tol*(1+0.01*big), then the sequential swap partition
   that moves significant entries to the front (it defines the list order the
   ratio tests break ties in) */
void o_sort_list(int *ind, const double *vec, int nnz, double tol_piv,
                 int *num_out, double *max_out)
{
    double big = 0.0;
    for (int pos = 1; pos <= nnz; pos++) {
        double temp = fabs(vec[ind[pos]]);
        if (big < temp) big = temp;
    }
    double eps = tol_piv * (1.0 + 0.01 * big);
    int num = 0;
    while (num < nnz) {
        int i = ind[nnz];
        if (fabs(vec[i]) < eps)
            nnz--;
        else {
            num++;
            ind[nnz] = ind[num];
            ind[num] = i;
        }
    }
    *num_out = num;
    *max_out = big;
}

/* lib/glpspx01.js:808-1028 chuzr: textbook (rtol == 0) or Harris two-pass
   ratio test.  p = 0 none, p = -1 bound flip of xN[q]. */
void o_chuzr_primal(int m, const signed char *type, const double *lb,
                    const double *ub, const double *coef, const int *head,
                    int phase, const double *bbar, double cbar_q, int q,
                    const int *tcol_ind, const double *tcol_vec, int tcol_num,
                    double rtol, int *p_out, int *p_stat_out, double *teta_out)
{
    int p, p_stat, i_stat = 0;
    double teta, big, tmax;
    const double s = (cbar_q > 0.0 ? -1.0 : +1.0);
    int k = head[m + q];
    if (type[k] == GLP_DB) { p = -1; p_stat = 0; teta = ub[k] - lb[k]; big = 1.0; }
    else { p = 0; p_stat = 0; teta = DBL_MAX; big = 0.0; }
    for (int pass = 1; pass <= 2; pass++) {
        const double r = (pass == 1 ? rtol : 0.0);
        if (pass == 2) {
            if (rtol == 0.0 || p <= 0 || teta == 0.0) break;
            tmax = teta;
            p = 0; p_stat = 0; teta = DBL_MAX; big = 0.0;
        }
        for (int pos = 1; pos <= tcol_num; pos++) {
            int i = tcol_ind[pos];
            k = head[i];
            double alfa = s * tcol_vec[i], t, delta;
            if (alfa > 0.0) {
                if (phase == 1 && coef[k] < 0.0) {
                    delta = r * (1.0 + kappa * fabs(lb[k]));
                    t = ((lb[k] + delta) - bbar[i]) / alfa;
                    i_stat = GLP_NL;
                } else if (phase == 1 && coef[k] > 0.0)
                    continue;
                else if (type[k] == GLP_UP || type[k] == GLP_DB || type[k] == GLP_FX) {
                    delta = r * (1.0 + kappa * fabs(ub[k]));
                    t = ((ub[k] + delta) - bbar[i]) / alfa;
                    i_stat = GLP_NU;
                } else
                    continue;
            } else {
                if (phase == 1 && coef[k] > 0.0) {
                    delta = r * (1.0 + kappa * fabs(ub[k]));
                    t = ((ub[k] - delta) - bbar[i]) / alfa;
                    i_stat = GLP_NU;
                } else if (phase == 1 && coef[k] < 0.0)
                    continue;
                else if (type[k] == GLP_LO || type[k] == GLP_DB || type[k] == GLP_FX) {
                    delta = r * (1.0 + kappa * fabs(lb[k]));
                    t = ((lb[k] - delta) - bbar[i]) / alfa;
                    i_stat = GLP_NL;
                } else
                    continue;
            }
            if (t < 0.0) t = 0.0;
            if (pass == 1) {
                if (teta > t || (teta == t && big < fabs(alfa))) {
                    p = i; p_stat = i_stat; teta = t; big = fabs(alfa);
                }
            } else {
                if (t <= tmax && big < fabs(alfa)) {
                    p = i; p_stat = i_stat; teta = t; big = fabs(alfa);
                }
            }
        }
        if (pass == 2) assert(p != 0);
    }
    *p_out = p;
    *p_stat_out = (p > 0 && type[head[p]] == GLP_FX) ? GLP_NS : p_stat;
    *teta_out = s * teta;
}

/* NOTE on pass 2 above: the reference's second pass computes
   t = (bound - bbar)/alfa with no delta (lib/glpspx01.js:959,970,983,994);
   with r = 0 the expression ((bound +/- 0*(...)) - bbar)/alfa is the same
   floating-point value because x + 0.0 == x and x - 0.0 == x exactly. */

namespace {

struct Primal {
    CSA c;
    Prob &lp;
    const SMCP &parm;
    const Hook *hook;
    Primal(Prob &lp_, const SMCP &parm_, const Hook *h) : lp(lp_), parm(parm_), hook(h) {}

    void fire(int ev) { if (hook && hook->fn) hook->fn(hook->user, ev, &c); }

    /* lib/glpspx01.js:340-375 add_N_col */
    void add_N_col(int j, int k)
    {
        const int m = c.m;
        if (k <= m) {
            int pos = c.N_ptr[k] + (c.N_len[k]++);
            c.N_ind[pos] = j; c.N_val[pos] = 1.0;
        } else {
            int beg = c.A_ptr[k - m], end = c.A_ptr[k - m + 1];
            for (int ptr = beg; ptr < end; ptr++) {
                int i = c.A_ind[ptr];
                int pos = c.N_ptr[i] + (c.N_len[i]++);
                c.N_ind[pos] = j; c.N_val[pos] = -c.A_val[ptr];
            }
        }
    }

    void del_N_elem(int i, int j)
    {
        int head = c.N_ptr[i], pos;
        for (pos = head; c.N_ind[pos] != j; pos++) {}
        int tail = head + (--c.N_len[i]);
        c.N_ind[pos] = c.N_ind[tail];
        c.N_val[pos] = c.N_val[tail];
    }

    /* lib/glpspx01.js:377-419 del_N_col */
    void del_N_col(int j, int k)
    {
        const int m = c.m;
        if (k <= m) del_N_elem(k, j);
        else {
            int beg = c.A_ptr[k - m], end = c.A_ptr[k - m + 1];
            for (int ptr = beg; ptr < end; ptr++) del_N_elem(c.A_ind[ptr], j);
        }
    }

    /* lib/glpspx01.js:309-338 alloc_N + 421-440 build_N */
    void alloc_build_N()
    {
        const int m = c.m, n = c.n;
        c.N_ptr.assign(1 + m + 1, 0); c.N_len.assign(1 + m, 0);
        for (int i = 1; i <= m; i++) c.N_len[i] = 1;
        for (int j = 1; j <= n; j++)
            for (int ptr = c.A_ptr[j]; ptr < c.A_ptr[j + 1]; ptr++) c.N_len[c.A_ind[ptr]]++;
        c.N_ptr[1] = 1;
        for (int i = 1; i <= m; i++) {
            if (c.N_len[i] > n) c.N_len[i] = n;
            c.N_ptr[i + 1] = c.N_ptr[i] + c.N_len[i];
        }
        c.N_ind.assign(c.N_ptr[m + 1], 0);
        c.N_val.assign(c.N_ptr[m + 1], 0.0);
        for (int i = 1; i <= m; i++) c.N_len[i] = 0;
        for (int j = 1; j <= n; j++)
            if (c.stat[j] != GLP_NS) add_N_col(j, c.head[m + j]);
    }

    /* lib/glpspx01.js:42-145 init_csa */
    void init()
    {
        spx_init_common(c, lp, false);
        alloc_build_N();
        for (int j = 1; j <= c.n; j++) c.gamma[j] = 1.0;
    }

    /* lib/glpspx01.js:586-601 reset_refsp */
    void reset_refsp()
    {
        assert(c.refct == 0);
        c.refct = 1000;
        std::fill(c.refsp.begin(), c.refsp.end(), 0);
        for (int j = 1; j <= c.n; j++) { c.refsp[c.head[c.m + j]] = 1; c.gamma[j] = 1.0; }
    }

    void chuzc(double tol_dj)
    {
        c.hook_tol = tol_dj;
        c.q = o_chuzc_primal(c.n, c.stat.data(), c.cbar.data(), c.gamma.data(), tol_dj);
    }

    void sort_tcol(double tol_piv)
    {
        o_sort_list(c.tcol_ind.data(), c.tcol_vec.data(), c.tcol_nnz, tol_piv,
                    &c.tcol_num, &c.tcol_max);
    }

    void chuzr(double rtol)
    {
        c.hook_tol = rtol;
        o_chuzr_primal(c.m, c.type.data(), c.lb.data(), c.ub.data(), c.coef.data(),
                       c.head.data(), c.phase, c.bbar.data(), c.cbar[c.q], c.q,
                       c.tcol_ind.data(), c.tcol_vec.data(), c.tcol_num, rtol,
                       &c.p, &c.p_stat, &c.teta);
    }

    /* lib/glpspx01.js:1030-1042 eval_rho */
    void eval_rho(double *rho)
    {
        for (int i = 1; i <= c.m; i++) rho[i] = 0.0;
        rho[c.p] = 1.0;
        bfd_btran(*c.bfd, rho);
    }

    /* lib/glpspx01.js:1044-1056 refine_rho */
    void refine_rho(double *rho)
    {
        double *e = c.work3.data();
        for (int i = 1; i <= c.m; i++) e[i] = 0.0;
        e[c.p] = 1.0;
        spx_refine_btran(c, e, rho);
    }

    /* lib/glpspx01.js:1058-1098 eval_trow: trow = -sum_i rho_i N'[i] */
    void eval_trow(const double *rho)
    {
        const int m = c.m, n = c.n;
        for (int j = 1; j <= n; j++) c.trow_vec[j] = 0.0;
        for (int i = 1; i <= m; i++) {
            double temp = rho[i];
            if (temp == 0.0) continue;
            int beg = c.N_ptr[i], end = beg + c.N_len[i];
            for (int ptr = beg; ptr < end; ptr++) c.trow_vec[c.N_ind[ptr]] -= temp * c.N_val[ptr];
        }
        int nnz = 0;
        for (int j = 1; j <= n; j++)
            if (c.trow_vec[j] != 0.0) c.trow_ind[++nnz] = j;
        c.trow_nnz = nnz;
    }

    /* lib/glpspx01.js:1100-1131 update_bbar */
    void update_bbar()
    {
        if (c.p > 0) c.bbar[c.p] = spx_get_xN(c, c.q) + c.teta;
        if (c.teta == 0.0) return;
        for (int pos = 1; pos <= c.tcol_nnz; pos++) {
            int i = c.tcol_ind[pos];
            if (i == c.p) continue;
            c.bbar[i] += c.tcol_vec[i] * c.teta;
        }
    }

    /* lib/glpspx01.js:1133-1152 reeval_cost */
    double reeval_cost()
    {
        double dq = c.coef[c.head[c.m + c.q]];
        for (int pos = 1; pos <= c.tcol_nnz; pos++) {
            int i = c.tcol_ind[pos];
            dq += c.coef[c.head[i]] * c.tcol_vec[i];
        }
        return dq;
    }

    /* lib/glpspx01.js:1154-1176 update_cbar */
    void update_cbar()
    {
        double new_dq = (c.cbar[c.q] /= c.trow_vec[c.q]);
        for (int pos = 1; pos <= c.trow_nnz; pos++) {
            int j = c.trow_ind[pos];
            if (j == c.q) continue;
            c.cbar[j] -= c.trow_vec[j] * new_dq;
        }
    }

    /* lib/glpspx01.js:1178-1255 update_gamma (projected steepest edge) */
    void update_gamma()
    {
        const int m = c.m;
        double *u = c.work3.data();
        assert(c.refct > 0);
        c.refct--;
        double gamma_q, delta_q;
        gamma_q = delta_q = (c.refsp[c.head[m + c.q]] ? 1.0 : 0.0);
        for (int i = 1; i <= m; i++) u[i] = 0.0;
        for (int pos = 1; pos <= c.tcol_nnz; pos++) {
            int i = c.tcol_ind[pos];
            if (c.refsp[c.head[i]]) {
                double t = c.tcol_vec[i];
                u[i] = t;
                gamma_q += t * t;
            } else
                u[i] = 0.0;
        }
        bfd_btran(*c.bfd, u);
        double pivot = c.trow_vec[c.q];
        for (int pos = 1; pos <= c.trow_nnz; pos++) {
            int j = c.trow_ind[pos];
            if (j == c.q) continue;
            double t = c.trow_vec[j] / pivot, s;
            int k = c.head[m + j];
            if (k <= m) s = u[k];
            else {
                s = 0.0;
                for (int ptr = c.A_ptr[k - m]; ptr < c.A_ptr[k - m + 1]; ptr++)
                    s -= c.A_val[ptr] * u[c.A_ind[ptr]];
            }
            double t1 = c.gamma[j] + t * t * gamma_q + 2.0 * t * s;
            double t2 = (c.refsp[k] ? 1.0 : 0.0) + delta_q * t * t;
            c.gamma[j] = (t1 >= t2 ? t1 : t2);
            if (c.gamma[j] < DBL_EPSILON) c.gamma[j] = DBL_EPSILON;
        }
        if (c.type[c.head[c.p]] == GLP_FX)
            c.gamma[c.q] = 1.0;
        else {
            c.gamma[c.q] = gamma_q / (pivot * pivot);
            if (c.gamma[c.q] < DBL_EPSILON) c.gamma[c.q] = DBL_EPSILON;
        }
    }

    /* lib/glpspx01.js:1310-1371 change_basis */
    void change_basis()
    {
        if (c.p < 0) {
            switch (c.stat[c.q]) {
            case GLP_NL: c.stat[c.q] = GLP_NU; break;
            case GLP_NU: c.stat[c.q] = GLP_NL; break;
            default: assert(!"bad flip");
            }
        } else {
            int k = c.head[c.p];
            c.head[c.p] = c.head[c.m + c.q];
            c.head[c.m + c.q] = k;
            c.stat[c.q] = (signed char)c.p_stat;
        }
    }

    /* lib/glpspx01.js:1373-1414 set_aux_obj */
    int set_aux_obj(double tol_bnd)
    {
        const int m = c.m, n = c.n;
        int cnt = 0;
        tol_bnd *= 0.90;
        for (int k = 1; k <= m + n; k++) c.coef[k] = 0.0;
        for (int i = 1; i <= m; i++) {
            int k = c.head[i];
            if (c.type[k] == GLP_LO || c.type[k] == GLP_DB || c.type[k] == GLP_FX) {
                double eps = tol_bnd * (1.0 + kappa * fabs(c.lb[k]));
                if (c.bbar[i] < c.lb[k] - eps) { c.coef[k] = -1.0; cnt++; }
            }
            if (c.type[k] == GLP_UP || c.type[k] == GLP_DB || c.type[k] == GLP_FX) {
                double eps = tol_bnd * (1.0 + kappa * fabs(c.ub[k]));
                if (c.bbar[i] > c.ub[k] + eps) { c.coef[k] = +1.0; cnt++; }
            }
        }
        return cnt;
    }

    /* lib/glpspx01.js:1416-1427 set_orig_obj */
    void set_orig_obj()
    {
        for (int i = 1; i <= c.m; i++) c.coef[i] = 0.0;
        for (int j = 1; j <= c.n; j++) c.coef[c.m + j] = c.zeta * c.obj[j];
    }

    /* lib/glpspx01.js:1429-1481 check_stab */
    int check_stab(double tol_bnd)
    {
        for (int i = 1; i <= c.m; i++) {
            int k = c.head[i];
            double eps;
            if (c.phase == 1 && c.coef[k] < 0.0) {
                eps = tol_bnd * (1.0 + kappa * fabs(c.lb[k]));
                if (c.bbar[i] > c.lb[k] + eps) return 1;
            } else if (c.phase == 1 && c.coef[k] > 0.0) {
                eps = tol_bnd * (1.0 + kappa * fabs(c.ub[k]));
                if (c.bbar[i] < c.ub[k] - eps) return 1;
            } else {
                if (c.type[k] == GLP_LO || c.type[k] == GLP_DB || c.type[k] == GLP_FX) {
                    eps = tol_bnd * (1.0 + kappa * fabs(c.lb[k]));
                    if (c.bbar[i] < c.lb[k] - eps) return 1;
                }
                if (c.type[k] == GLP_UP || c.type[k] == GLP_DB || c.type[k] == GLP_FX) {
                    eps = tol_bnd * (1.0 + kappa * fabs(c.ub[k]));
                    if (c.bbar[i] > c.ub[k] + eps) return 1;
                }
            }
        }
        return 0;
    }

    /* lib/glpspx01.js:1483-1522 check_feas (phase 1 only) */
    int check_feas(double tol_bnd)
    {
        assert(c.phase == 1);
        for (int i = 1; i <= c.m; i++) {
            int k = c.head[i];
            double eps;
            if (c.coef[k] < 0.0) {
                eps = tol_bnd * (1.0 + kappa * fabs(c.lb[k]));
                if (c.bbar[i] < c.lb[k] - eps) return 1;
            } else if (c.coef[k] > 0.0) {
                eps = tol_bnd * (1.0 + kappa * fabs(c.ub[k]));
                if (c.bbar[i] > c.ub[k] + eps) return 1;
            }
        }
        return 0;
    }

    /* termination on it_lim / tm_lim: lib/glpspx01.js:1813-1832 */
    int stop_on_limit(int code)
    {
        int p_stat, d_stat;
        if (c.phase == 1) { p_stat = GLP_INFEAS; set_orig_obj(); spx_eval_cbar(c); }
        else p_stat = GLP_FEAS;
        chuzc(parm.tol_dj);
        d_stat = (c.q == 0 ? GLP_FEAS : GLP_INFEAS);
        spx_store_sol(c, lp, p_stat, d_stat, 0);
        return code;
    }

    /* main loop: lib/glpspx01.js:1684-2056 */
    int run()
    {
        int binv_st = 2, bbar_st = 0, cbar_st = 0, rigorous = 0;
        int p_stat, d_stat, ret;
        init();
        for (;;) {
            if (binv_st == 0) {
                ret = spx_invert_B(c);
                if (ret != 0) return spx_fail(c, lp);
                c.valid = 1;
                binv_st = 1;
                bbar_st = cbar_st = 0;
            }
            if (bbar_st == 0) {
                spx_eval_bbar(c);
                bbar_st = 1;
                if (c.phase == 0) {
                    if (set_aux_obj(parm.tol_bnd) > 0)
                        c.phase = 1;
                    else { set_orig_obj(); c.phase = 2; }
                    assert(check_stab(parm.tol_bnd) == 0);
                    cbar_st = 0;
                }
                if (check_stab(parm.tol_bnd)) {
                    c.phase = 0; binv_st = 0; rigorous = 5;
                    continue;
                }
            }
            assert(c.phase == 1 || c.phase == 2);
            if (c.phase == 1 && !check_feas(parm.tol_bnd)) {
                c.phase = 2;
                set_orig_obj();
                cbar_st = 0;
            }
            if (cbar_st == 0) { spx_eval_cbar(c); cbar_st = 1; }
            if (parm.pricing == GLP_PT_PSE && c.refct == 0) reset_refsp();
            if (parm.it_lim < INT_MAX && c.it_cnt - c.it_beg >= parm.it_lim) {
                if (bbar_st != 1 || (c.phase == 2 && cbar_st != 1)) {
                    if (bbar_st != 1) bbar_st = 0;
                    if (c.phase == 2 && cbar_st != 1) cbar_st = 0;
                    continue;
                }
                return stop_on_limit(GLP_EITLIM);
            }
            if (parm.tm_lim < INT_MAX && (xtime_ms() - c.tm_beg) >= parm.tm_lim) {
                if (bbar_st != 1 || (c.phase == 2 && cbar_st != 1)) {
                    if (bbar_st != 1) bbar_st = 0;
                    if (c.phase == 2 && cbar_st != 1) cbar_st = 0;
                    continue;
                }
                return stop_on_limit(GLP_ETMLIM);
            }
            chuzc(parm.tol_dj);
            fire(EV_P_CHUZC);
            if (c.q == 0) {
                if (bbar_st != 1 || cbar_st != 1) {
                    if (bbar_st != 1) bbar_st = 0;
                    if (cbar_st != 1) cbar_st = 0;
                    continue;
                }
                if (c.phase == 1) {
                    p_stat = GLP_NOFEAS;
                    set_orig_obj();
                    spx_eval_cbar(c);
                    chuzc(parm.tol_dj);
                    d_stat = (c.q == 0 ? GLP_FEAS : GLP_INFEAS);
                } else
                    p_stat = d_stat = GLP_FEAS;
                spx_store_sol(c, lp, p_stat, d_stat, 0);
                return 0;
            }
            spx_eval_tcol(c);
            if (rigorous) spx_refine_tcol(c);
            sort_tcol(parm.tol_piv);
            {   /* lib/glpspx01.js:1901-1919 accuracy of d_q */
                double d1 = c.cbar[c.q], d2 = reeval_cost();
                assert(d1 != 0.0);
                if (fabs(d1 - d2) > 1e-5 * (1.0 + fabs(d2)) ||
                    !((d1 < 0.0 && d2 < 0.0) || (d1 > 0.0 && d2 > 0.0))) {
                    if (cbar_st != 1 || !rigorous) {
                        if (cbar_st != 1) cbar_st = 0;
                        rigorous = 5;
                        continue;
                    }
                }
                if (d1 > 0.0) c.cbar[c.q] = (d2 > 0.0 ? d2 : +DBL_EPSILON);
                else c.cbar[c.q] = (d2 < 0.0 ? d2 : -DBL_EPSILON);
            }
            chuzr(parm.r_test == GLP_RT_STD ? 0.0 : 0.30 * parm.tol_bnd);
            fire(EV_P_CHUZR);
            if (c.p == 0) {
                if (bbar_st != 1 || cbar_st != 1 || !rigorous) {
                    if (bbar_st != 1) bbar_st = 0;
                    if (cbar_st != 1) cbar_st = 0;
                    rigorous = 1;
                    continue;
                }
                if (c.phase == 1) return spx_fail(c, lp);
                spx_store_sol(c, lp, GLP_FEAS, GLP_NOFEAS, c.head[c.m + c.q]);
                return 0;
            }
            if (c.p > 0) {
                double piv = c.tcol_vec[c.p];
                double eps = 1e-5 * (1.0 + 0.01 * c.tcol_max);
                if (fabs(piv) < eps && !rigorous) { rigorous = 5; continue; }
            }
            if (c.p > 0) {
                double *rho = c.work4.data();
                eval_rho(rho);
                if (rigorous) refine_rho(rho);
                eval_trow(rho);
                fire(EV_P_TROW);
            }
            if (c.p > 0) {
                double piv1 = c.tcol_vec[c.p], piv2 = c.trow_vec[c.q];
                assert(piv1 != 0.0);
                if (fabs(piv1 - piv2) > 1e-8 * (1.0 + fabs(piv1)) ||
                    !((piv1 > 0.0 && piv2 > 0.0) || (piv1 < 0.0 && piv2 < 0.0))) {
                    if (binv_st != 1 || !rigorous) {
                        if (binv_st != 1) binv_st = 0;
                        rigorous = 5;
                        continue;
                    }
                    if (c.trow_vec[c.q] == 0.0) {
                        c.trow_nnz++;
                        c.trow_ind[c.trow_nnz] = c.q;
                    }
                    c.trow_vec[c.q] = piv1;
                }
            }
            update_bbar();
            bbar_st = 2;
            if (c.p > 0) {
                update_cbar();
                cbar_st = 2;
                if (c.phase == 1) {
                    int k = c.head[c.p];
                    c.cbar[c.q] -= c.coef[k];
                    c.coef[k] = 0.0;
                }
            }
            if (c.p > 0 && parm.pricing == GLP_PT_PSE && c.refct > 0) {
                update_gamma();
                fire(EV_P_GAMMA);
            }
            if (c.p > 0) {
                ret = spx_update_B(c, c.p, c.head[c.m + c.q]);
                if (ret == 0) binv_st = 2;
                else { c.valid = 0; binv_st = 0; }
            }
            if (c.p > 0) {
                del_N_col(c.q, c.head[c.m + c.q]);
                if (c.type[c.head[c.p]] != GLP_FX) add_N_col(c.q, c.head[c.p]);
            }
            change_basis();
            c.it_cnt++;
            if (rigorous > 0) rigorous--;
            fire(EV_P_ITER);
        }
    }
};

} /* anonymous namespace */

int spx_primal(Prob &lp, const SMCP &parm, const Hook *hook)
{
    Primal s(lp, parm, hook);
    return s.run();
}

} /* namespace glpo */
