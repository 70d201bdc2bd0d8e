/* dual.cpp -- CPU ORACLE (test infrastructure only; see glpo.h).
 * Restatement of the two-phase dual revised simplex, lib/glpspx02.js.
 */
#include "spx_common.h"

namespace glpo {

/* lib/glpspx02.js:572-625 chuzr: basic variable with the largest weighted
   bound violation r_i^2/max(gamma_i, eps); strict '<' keeps the lowest i */
int o_chuzr_dual(int m, const signed char *type, const double *lb,
                 const double *ub, const int *head, const double *bbar,
                 const double *gamma, double tol_bnd, double *delta_out)
{
    int p = 0;
    double delta = 0.0, best = 0.0;
    for (int i = 1; i <= m; i++) {
        int k = head[i];
        double ri = 0.0, eps;
        if (type[k] == GLP_LO || type[k] == GLP_DB || type[k] == GLP_FX) {
            eps = tol_bnd * (1.0 + kappa * fabs(lb[k]));
            if (bbar[i] < lb[k] - eps) ri = lb[k] - bbar[i];
        }
        if (type[k] == GLP_UP || type[k] == GLP_DB || type[k] == GLP_FX) {
            eps = tol_bnd * (1.0 + kappa * fabs(ub[k]));
            if (bbar[i] > ub[k] + eps) ri = ub[k] - bbar[i];
        }
        if (ri == 0.0) continue;
        double temp = gamma[i];
        if (temp < DBL_EPSILON) temp = DBL_EPSILON;
        temp = (ri * ri) / temp;
        if (best < temp) { p = i; delta = ri; best = temp; }
    }
    *delta_out = delta;
    return p;
}

/* lib/glpspx02.js:793-935 chuzc: dual ratio test, textbook or Harris */
void o_chuzc_dual(const signed char *stat, const double *cbar, double delta,
                  const int *trow_ind, const double *trow_vec, int trow_num,
                  double rtol, int *q_out, double *new_dq_out)
{
    const double s = (delta > 0.0 ? +1.0 : -1.0);
    int q = 0;
    double teta = DBL_MAX, big = 0.0, tmax = 0.0;
    for (int pass = 1; pass <= 2; pass++) {
        if (pass == 2) {
            if (rtol == 0.0 || q == 0 || teta == 0.0) break;
            tmax = teta;
            q = 0; teta = DBL_MAX; big = 0.0;
        }
        for (int pos = 1; pos <= trow_num; pos++) {
            int j = trow_ind[pos];
            double alfa = s * trow_vec[j], t;
            if (alfa > 0.0) {
                if (stat[j] == GLP_NL || stat[j] == GLP_NF)
                    t = (pass == 1 ? (cbar[j] + rtol) : cbar[j]) / alfa;
                else
                    continue;
            } else {
                if (stat[j] == GLP_NU || stat[j] == GLP_NF)
                    t = (pass == 1 ? (cbar[j] - rtol) : cbar[j]) / alfa;
                else
                    continue;
            }
            if (t < 0.0) t = 0.0;
            if (pass == 1) {
                if (teta > t || (teta == t && big < fabs(alfa))) { q = j; teta = t; big = fabs(alfa); }
            } else {
                if (t <= tmax && big < fabs(alfa)) { q = j; teta = t; big = fabs(alfa); }
            }
        }
        if (pass == 2) assert(q != 0);
    }
    *q_out = q;
    *new_dq_out = s * teta;
}

namespace {

struct Dual {
    CSA c;
    Prob &lp;
    const SMCP &parm;
    const Hook *hook;
    Dual(Prob &lp_, const SMCP &parm_, const Hook *h) : lp(lp_), parm(parm_), hook(h) {}

    void fire(int ev) { if (hook && hook->fn) hook->fn(hook->user, ev, &c); }

    /* lib/glpspx02.js:89-190 init_csa */
    void init()
    {
        spx_init_common(c, lp, true);
        const int m = c.m, n = c.n, nnz = lp.nnz;
        c.orig_type = c.type; c.orig_lb = c.lb; c.orig_ub = c.ub;
        /* lib/glpspx02.js:148: working costs are pre-scaled by zeta */
        for (int j = 1; j <= n; j++) c.coef[m + j] *= c.zeta;
        /* row-wise copy of A in row-list order (lib/glpspx02.js:66-87) */
        c.AT_ptr.assign(1 + m + 1, 0); c.AT_ind.assign(1 + nnz, 0); c.AT_val.assign(1 + nnz, 0.0);
        int loc = 1;
        for (int i = 1; i <= m; i++) {
            c.AT_ptr[i] = loc;
            for (const Elem &e : lp.row_list[i]) {
                c.AT_ind[loc] = e.idx;
                c.AT_val[loc] = lp.r_rii[i] * e.val * lp.c_sjj[e.idx];
                loc++;
            }
        }
        c.AT_ptr[m + 1] = loc;
        assert(loc - 1 == nnz);
        c.bind.assign(1 + m + n, 0);
        for (int k = 1; k <= m + n; k++) c.bind[c.head[k]] = k;
        for (int i = 1; i <= m; i++) c.gamma[i] = 1.0;
    }

    /* lib/glpspx02.js:497-512 reset_refsp */
    void reset_refsp()
    {
        assert(c.refct == 0);
        c.refct = 1000;
        std::fill(c.refsp.begin(), c.refsp.end(), 0);
        for (int i = 1; i <= c.m; i++) { c.refsp[c.head[i]] = 1; c.gamma[i] = 1.0; }
    }

    void chuzr(double tol_bnd)
    {
        c.hook_tol = tol_bnd;
        c.p = o_chuzr_dual(c.m, c.type.data(), c.lb.data(), c.ub.data(), c.head.data(),
                           c.bbar.data(), c.gamma.data(), tol_bnd, &c.delta);
    }

    /* lib/glpspx02.js:627-639 eval_rho (the reference solves in the outer
       'rho', which aliases the argument -- SURVEY 8a row a9) */
    void eval_rho(double *rho)
    {
        for (int i = 1; i <= c.m; i++) rho[i] = 0.0;
        rho[c.p] = 1.0;
        bfd_btran(*c.bfd, rho);
    }

    /* lib/glpspx02.js:641-653 refine_rho */
    void refine_rho(double *rho)
    {
        double *e = c.work3.data();
        for (int i = 1; i <= c.m; i++) e[i] = 0.0;
        e[c.p] = 1.0;
        spx_refine_btran(c, e, rho);
    }

    /* lib/glpspx02.js:655-693 eval_trow1: column dots */
    void eval_trow1(const double *rho)
    {
        const int m = c.m, n = c.n;
        int nnz = 0;
        for (int j = 1; j <= n; j++) {
            if (c.stat[j] == GLP_NS) { c.trow_vec[j] = 0.0; continue; }
            int k = c.head[m + j];
            double temp;
            if (k <= m) temp = -rho[k];
            else {
                temp = 0.0;
                for (int ptr = c.A_ptr[k - m]; ptr < c.A_ptr[k - m + 1]; ptr++)
                    temp += rho[c.A_ind[ptr]] * c.A_val[ptr];
            }
            if (temp != 0.0) c.trow_ind[++nnz] = j;
            c.trow_vec[j] = temp;
        }
        c.trow_nnz = nnz;
    }

    /* lib/glpspx02.js:695-733 eval_trow2: row scatter */
    void eval_trow2(const double *rho)
    {
        const int m = c.m, n = c.n;
        for (int j = 1; j <= n; j++) c.trow_vec[j] = 0.0;
        for (int i = 1; i <= m; i++) {
            double temp = rho[i];
            if (temp == 0.0) continue;
            int j = c.bind[i] - m;
            if (j >= 1 && c.stat[j] != GLP_NS) c.trow_vec[j] -= temp;
            for (int ptr = c.AT_ptr[i]; ptr < c.AT_ptr[i + 1]; ptr++) {
                j = c.bind[m + c.AT_ind[ptr]] - m;
                if (j >= 1 && c.stat[j] != GLP_NS) c.trow_vec[j] += temp * c.AT_val[ptr];
            }
        }
        int nnz = 0;
        for (int j = 1; j <= n; j++)
            if (c.trow_vec[j] != 0.0) c.trow_ind[++nnz] = j;
        c.trow_nnz = nnz;
    }

    /* lib/glpspx02.js:735-752 eval_trow: 20 % density switch */
    void eval_trow(const double *rho)
    {
        int nnz = 0;
        for (int i = 1; i <= c.m; i++)
            if (rho[i] != 0.0) nnz++;
        double dens = (double)nnz / (double)c.m;
        if (dens >= 0.20) eval_trow1(rho); else eval_trow2(rho);
    }

    void sort_trow(double tol_piv)
    {
        o_sort_list(c.trow_ind.data(), c.trow_vec.data(), c.trow_nnz, tol_piv,
                    &c.trow_num, &c.trow_max);
    }

    void chuzc(double rtol)
    {
        c.hook_tol = rtol;
        o_chuzc_dual(c.stat.data(), c.cbar.data(), c.delta, c.trow_ind.data(),
                     c.trow_vec.data(), c.trow_num, rtol, &c.q, &c.new_dq);
    }

    /* lib/glpspx02.js:1020-1040 update_cbar */
    void update_cbar()
    {
        c.cbar[c.q] = c.new_dq;
        if (c.new_dq == 0.0) return;
        for (int pos = 1; pos <= c.trow_nnz; pos++) {
            int j = c.trow_ind[pos];
            if (j != c.q) c.cbar[j] -= c.trow_vec[j] * c.new_dq;
        }
    }

    /* lib/glpspx02.js:1042-1073 update_bbar */
    void update_bbar()
    {
        double teta = c.delta / c.tcol_vec[c.p];
        c.bbar[c.p] = spx_get_xN(c, c.q) + teta;
        if (teta == 0.0) return;
        for (int pos = 1; pos <= c.tcol_nnz; pos++) {
            int i = c.tcol_ind[pos];
            if (i != c.p) c.bbar[i] += c.tcol_vec[i] * teta;
        }
    }

    /* lib/glpspx02.js:1075-1188 update_gamma (dual projected steepest edge) */
    void update_gamma()
    {
        const int m = c.m;
        double *u = c.work3.data();
        assert(c.refct > 0);
        c.refct--;
        double gamma_p, eta_p;
        gamma_p = eta_p = (c.refsp[c.head[c.p]] ? 1.0 : 0.0);
        for (int i = 1; i <= m; i++) u[i] = 0.0;
        for (int pos = 1; pos <= c.trow_nnz; pos++) {
            int j = c.trow_ind[pos];
            int k = c.head[m + j];
            if (!c.refsp[k]) continue;
            double t = c.trow_vec[j];
            gamma_p += t * t;
            if (k <= m) u[k] += t;
            else
                for (int ptr = c.A_ptr[k - m]; ptr < c.A_ptr[k - m + 1]; ptr++)
                    u[c.A_ind[ptr]] -= t * c.A_val[ptr];
        }
        bfd_ftran(*c.bfd, u);
        double pivot = c.tcol_vec[c.p];
        for (int pos = 1; pos <= c.tcol_nnz; pos++) {
            int i = c.tcol_ind[pos];
            int k = c.head[i];
            if (i == c.p) continue;
            if (c.type[k] == GLP_FR) continue;
            double t = c.tcol_vec[i] / pivot;
            double t1 = c.gamma[i] + t * t * gamma_p + 2.0 * t * u[i];
            double t2 = (c.refsp[k] ? 1.0 : 0.0) + eta_p * t * t;
            c.gamma[i] = (t1 >= t2 ? t1 : t2);
            if (c.gamma[i] < DBL_EPSILON) c.gamma[i] = DBL_EPSILON;
        }
        if (c.type[c.head[m + c.q]] == GLP_FR)
            c.gamma[c.p] = 1.0;
        else {
            c.gamma[c.p] = gamma_p / (pivot * pivot);
            if (c.gamma[c.p] < DBL_EPSILON) c.gamma[c.p] = DBL_EPSILON;
        }
        /* a fixed leaving variable drops out of the reference space */
        int k = c.head[c.p];
        if (c.type[k] == GLP_FX && c.refsp[k]) {
            c.refsp[k] = 0;
            for (int pos = 1; pos <= c.tcol_nnz; pos++) {
                int i = c.tcol_ind[pos];
                double t;
                if (i == c.p) {
                    if (c.type[c.head[m + c.q]] == GLP_FR) continue;
                    t = 1.0 / c.tcol_vec[c.p];
                } else {
                    if (c.type[c.head[i]] == GLP_FR) continue;
                    t = c.tcol_vec[i] / c.tcol_vec[c.p];
                }
                c.gamma[i] -= t * t;
                if (c.gamma[i] < DBL_EPSILON) c.gamma[i] = DBL_EPSILON;
            }
        }
    }

    /* lib/glpspx02.js:1259-1294 change_basis */
    void change_basis()
    {
        const int m = c.m;
        int k = c.head[c.p];
        c.head[c.p] = c.head[m + c.q];
        c.head[m + c.q] = k;
        c.bind[c.head[c.p]] = c.p;
        c.bind[c.head[m + c.q]] = m + c.q;
        if (c.type[k] == GLP_FX) c.stat[c.q] = GLP_NS;
        else if (c.delta > 0.0) c.stat[c.q] = GLP_NL;
        else c.stat[c.q] = GLP_NU;
    }

    /* lib/glpspx02.js:1296-1315 check_feas */
    int check_feas(double tol_dj)
    {
        for (int j = 1; j <= c.n; j++) {
            int k = c.head[c.m + j];
            if (c.cbar[j] < -tol_dj)
                if (c.orig_type[k] == GLP_LO || c.orig_type[k] == GLP_FR) return 1;
            if (c.cbar[j] > +tol_dj)
                if (c.orig_type[k] == GLP_UP || c.orig_type[k] == GLP_FR) return 1;
        }
        return 0;
    }

    /* lib/glpspx02.js:1317-1359 set_aux_bnds */
    void set_aux_bnds()
    {
        const int m = c.m, n = c.n;
        for (int k = 1; k <= m + n; k++) {
            switch (c.orig_type[k]) {
            case GLP_FR: c.type[k] = GLP_DB; c.lb[k] = -1e3; c.ub[k] = +1e3; break;
            case GLP_LO: c.type[k] = GLP_DB; c.lb[k] = 0.0; c.ub[k] = +1.0; break;
            case GLP_UP: c.type[k] = GLP_DB; c.lb[k] = -1.0; c.ub[k] = 0.0; break;
            case GLP_DB: case GLP_FX: c.type[k] = GLP_FX; c.lb[k] = c.ub[k] = 0.0; break;
            default: assert(!"bad type");
            }
        }
        for (int j = 1; j <= n; j++) {
            int k = c.head[m + j];
            if (c.type[k] == GLP_FX) c.stat[j] = GLP_NS;
            else if (c.cbar[j] >= 0.0) c.stat[j] = GLP_NL;
            else c.stat[j] = GLP_NU;
        }
    }

    /* lib/glpspx02.js:1361-1408 set_orig_bnds */
    void set_orig_bnds()
    {
        const int m = c.m, n = c.n;
        c.type = c.orig_type; c.lb = c.orig_lb; c.ub = c.orig_ub;
        for (int j = 1; j <= n; j++) {
            int k = c.head[m + j];
            switch (c.type[k]) {
            case GLP_FR: c.stat[j] = GLP_NF; break;
            case GLP_LO: c.stat[j] = GLP_NL; break;
            case GLP_UP: c.stat[j] = GLP_NU; break;
            case GLP_DB:
                if (c.cbar[j] >= +DBL_EPSILON) c.stat[j] = GLP_NL;
                else if (c.cbar[j] <= -DBL_EPSILON) c.stat[j] = GLP_NU;
                else if (fabs(c.lb[k]) <= fabs(c.ub[k])) c.stat[j] = GLP_NL;
                else c.stat[j] = GLP_NU;
                break;
            case GLP_FX: c.stat[j] = GLP_NS; break;
            default: assert(!"bad type");
            }
        }
    }

    /* lib/glpspx02.js:1410-1422 check_stab */
    int check_stab(double tol_dj)
    {
        for (int j = 1; j <= c.n; j++) {
            if (c.cbar[j] < -tol_dj)
                if (c.stat[j] == GLP_NL || c.stat[j] == GLP_NF) return 1;
            if (c.cbar[j] > +tol_dj)
                if (c.stat[j] == GLP_NU || c.stat[j] == GLP_NF) return 1;
        }
        return 0;
    }

    int stop_on_limit(int code)
    {
        int d_stat;
        if (c.phase == 1) { d_stat = GLP_INFEAS; set_orig_bnds(); spx_eval_bbar(c); }
        else d_stat = GLP_FEAS;
        spx_store_sol(c, lp, GLP_INFEAS, d_stat, 0);
        return code;
    }

    /* main loop: lib/glpspx02.js:1592-1966 */
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
            if (cbar_st == 0) {
                spx_eval_cbar(c);
                cbar_st = 1;
                if (c.phase == 0) {
                    if (check_feas(0.90 * parm.tol_dj) != 0) { c.phase = 1; set_aux_bnds(); }
                    else { c.phase = 2; set_orig_bnds(); }
                    assert(check_stab(parm.tol_dj) == 0);
                    c.refct = 0;
                    bbar_st = 0;
                }
                if (check_stab(parm.tol_dj) != 0) {
                    if (parm.meth == GLP_DUALP) {
                        spx_store_sol(c, lp, GLP_UNDEF, GLP_UNDEF, 0);
                        return GLP_EFAIL;
                    }
                    c.phase = 0; binv_st = 0; rigorous = 5;
                    continue;
                }
            }
            assert(c.phase == 1 || c.phase == 2);
            if (c.phase == 1 && check_feas(parm.tol_dj) == 0) {
                c.phase = 2;
                if (cbar_st != 1) { spx_eval_cbar(c); cbar_st = 1; }
                set_orig_bnds();
                c.refct = 0;
                bbar_st = 0;
            }
            if (bbar_st == 0) {
                spx_eval_bbar(c);
                if (c.phase == 2) c.bbar[0] = spx_eval_obj(c);
                bbar_st = 1;
            }
            if (parm.pricing == GLP_PT_PSE && c.refct == 0) reset_refsp();
            if (c.phase == 2 && c.zeta < 0.0 && parm.obj_ll > -DBL_MAX && c.bbar[0] <= parm.obj_ll) {
                if (bbar_st != 1 || cbar_st != 1) {
                    if (bbar_st != 1) bbar_st = 0;
                    if (cbar_st != 1) cbar_st = 0;
                    continue;
                }
                spx_store_sol(c, lp, GLP_INFEAS, GLP_FEAS, 0);
                return GLP_EOBJLL;
            }
            if (c.phase == 2 && c.zeta > 0.0 && parm.obj_ul < +DBL_MAX && c.bbar[0] >= parm.obj_ul) {
                if (bbar_st != 1 || cbar_st != 1) {
                    if (bbar_st != 1) bbar_st = 0;
                    if (cbar_st != 1) cbar_st = 0;
                    continue;
                }
                spx_store_sol(c, lp, GLP_INFEAS, GLP_FEAS, 0);
                return GLP_EOBJUL;
            }
            if (parm.it_lim < INT_MAX && c.it_cnt - c.it_beg >= parm.it_lim) {
                if ((c.phase == 2 && bbar_st != 1) || cbar_st != 1) {
                    if (c.phase == 2 && bbar_st != 1) bbar_st = 0;
                    if (cbar_st != 1) cbar_st = 0;
                    continue;
                }
                return stop_on_limit(GLP_EITLIM);
            }
            if (parm.tm_lim < INT_MAX && (xtime_ms() - c.tm_beg) >= parm.tm_lim) {
                if ((c.phase == 2 && bbar_st != 1) || cbar_st != 1) {
                    if (c.phase == 2 && bbar_st != 1) bbar_st = 0;
                    if (cbar_st != 1) cbar_st = 0;
                    continue;
                }
                return stop_on_limit(GLP_ETMLIM);
            }
            chuzr(parm.tol_bnd);
            fire(EV_D_CHUZR);
            if (c.p == 0) {
                if (bbar_st != 1 || cbar_st != 1) {
                    if (bbar_st != 1) bbar_st = 0;
                    if (cbar_st != 1) cbar_st = 0;
                    continue;
                }
                if (c.phase == 1) {
                    set_orig_bnds();
                    spx_eval_bbar(c);
                    p_stat = GLP_INFEAS; d_stat = GLP_NOFEAS;
                } else
                    p_stat = d_stat = GLP_FEAS;
                spx_store_sol(c, lp, p_stat, d_stat, 0);
                return 0;
            }
            {
                double *rho = c.work4.data();
                eval_rho(rho);
                if (rigorous) refine_rho(rho);
                eval_trow(rho);
                fire(EV_D_TROW);
                sort_trow(parm.tol_bnd); /* sic: tol_bnd, lib/glpspx02.js:1851 */
            }
            chuzc(parm.r_test == GLP_RT_STD ? 0.0 : 0.30 * parm.tol_dj);
            fire(EV_D_CHUZC);
            if (c.q == 0) {
                if (bbar_st != 1 || cbar_st != 1 || !rigorous) {
                    if (bbar_st != 1) bbar_st = 0;
                    if (cbar_st != 1) cbar_st = 0;
                    rigorous = 1;
                    continue;
                }
                if (c.phase == 1) return spx_fail(c, lp);
                spx_store_sol(c, lp, GLP_NOFEAS, GLP_FEAS, c.head[c.p]);
                return 0;
            }
            {
                double piv = c.trow_vec[c.q];
                double eps = 1e-5 * (1.0 + 0.01 * c.trow_max);
                if (fabs(piv) < eps && !rigorous) { rigorous = 5; continue; }
            }
            spx_eval_tcol(c);
            if (rigorous) spx_refine_tcol(c);
            {
                double piv1 = c.tcol_vec[c.p], piv2 = c.trow_vec[c.q];
                assert(piv1 != 0.0);
                if (fabs(piv1 - piv2) > 1e-8 * (1.0 + fabs(piv1)) ||
                    !((piv1 > 0.0 && piv2 > 0.0) || (piv1 < 0.0 && piv2 < 0.0))) {
                    if (binv_st != 1 || !rigorous) {
                        if (binv_st != 1) binv_st = 0;
                        rigorous = 5;
                        continue;
                    }
                    if (c.tcol_vec[c.p] == 0.0) {
                        c.tcol_nnz++;
                        c.tcol_ind[c.tcol_nnz] = c.p;
                    }
                    c.tcol_vec[c.p] = piv2;
                }
            }
            update_bbar();
            if (c.phase == 2)
                c.bbar[0] += (c.cbar[c.q] / c.zeta) * (c.delta / c.tcol_vec[c.p]);
            bbar_st = 2;
            update_cbar();
            cbar_st = 2;
            if (parm.pricing == GLP_PT_PSE && c.refct > 0) {
                update_gamma();
                fire(EV_D_GAMMA);
            }
            ret = spx_update_B(c, c.p, c.head[c.m + c.q]);
            if (ret == 0) binv_st = 2;
            else { c.valid = 0; binv_st = 0; }
            change_basis();
            c.it_cnt++;
            if (rigorous > 0) rigorous--;
            fire(EV_D_ITER);
        }
    }
};

} /* anonymous namespace */

int spx_dual(Prob &lp, const SMCP &parm, const Hook *hook)
{
    Dual s(lp, parm, hook);
    return s.run();
}

} /* namespace glpo */
