/* capi.cpp -- CPU ORACLE (test infrastructure only; see glpo.h).
 * Plain C entry points so tests/, smoke() and bench.py's cpu_baseline leg
 * can drive the oracle through ctypes.  Arrays crossing this boundary are
 * 0-based C arrays unless a comment says "CSA layout" (1-based, as the
 * reference keeps them, so kernel parity tests can hand them on verbatim).
 */
#include "csa.h"
#include "spx_common.h"
#include <cstring>

using namespace glpo;

namespace {
struct Handle {
    Prob P;
    Hook hook;
    std::string text;
    long nodes = 0;
};
}

#define API extern "C" __attribute__((visibility("default")))

API void *glpo_create(void) { return new Handle(); }
API void glpo_delete(void *h) { delete (Handle *)h; }

API int glpo_read_lp(void *h, const char *text, char *err, int errcap)
{
    Handle *H = (Handle *)h;
    std::string e;
    H->P.~Prob();
    new (&H->P) Prob();
    int ret = read_lp(H->P, text, e);
    if (err && errcap > 0) { strncpy(err, e.c_str(), errcap - 1); err[errcap - 1] = 0; }
    return ret;
}

API int glpo_write_lp(void *h, char *buf, int cap)
{
    Handle *H = (Handle *)h;
    H->text = write_lp(H->P);
    if (buf && cap > 0) { strncpy(buf, H->text.c_str(), cap - 1); buf[cap - 1] = 0; }
    return (int)H->text.size();
}

/* load a problem from 0-based CSC arrays; each column must be ascending in
   row index (the state glp_read_lp leaves after glp_sort_matrix) */
API void glpo_load(void *h, int m, int n, int dir, double c0,
                   const int *r_type, const double *r_lb, const double *r_ub,
                   const int *c_type, const double *c_lb, const double *c_ub,
                   const double *c_coef, const int *c_kind,
                   const int *A_ptr, const int *A_ind, const double *A_val)
{
    Handle *H = (Handle *)h;
    H->P.~Prob();
    new (&H->P) Prob();
    Prob &P = H->P;
    prob_add_rows(P, m);
    prob_add_cols(P, n);
    P.dir = dir; P.c0 = c0;
    for (int i = 1; i <= m; i++) prob_set_row_bnds(P, i, r_type[i - 1], r_lb[i - 1], r_ub[i - 1]);
    for (int j = 1; j <= n; j++) {
        prob_set_col_bnds(P, j, c_type[j - 1], c_lb[j - 1], c_ub[j - 1]);
        P.c_coef[j] = c_coef[j - 1];
        P.c_kind[j] = c_kind ? c_kind[j - 1] : GLP_CV;
    }
    for (int j = 1; j <= n; j++)
        for (int t = A_ptr[j - 1]; t < A_ptr[j]; t++) {
            if (A_val[t] == 0.0) continue;
            P.col_list[j].push_back(Elem{A_ind[t] + 1, A_val[t]});
            P.row_list[A_ind[t] + 1].push_back(Elem{j, A_val[t]});
            P.nnz++;
        }
}

API void glpo_dims(void *h, int *m, int *n, int *nnz, int *dir, double *c0)
{
    Prob &P = ((Handle *)h)->P;
    *m = P.m; *n = P.n; *nnz = P.nnz; *dir = P.dir; *c0 = P.c0;
}

/* export the problem as 0-based arrays (CSC in column-list order) */
API void glpo_export(void *h, int *r_type, double *r_lb, double *r_ub,
                     int *c_type, double *c_lb, double *c_ub, double *c_coef,
                     int *c_kind, int *A_ptr, int *A_ind, double *A_val,
                     double *rii, double *sjj)
{
    Prob &P = ((Handle *)h)->P;
    for (int i = 1; i <= P.m; i++) {
        r_type[i - 1] = P.r_type[i]; r_lb[i - 1] = P.r_lb[i]; r_ub[i - 1] = P.r_ub[i];
        if (rii) rii[i - 1] = P.r_rii[i];
    }
    int loc = 0;
    for (int j = 1; j <= P.n; j++) {
        c_type[j - 1] = P.c_type[j]; c_lb[j - 1] = P.c_lb[j]; c_ub[j - 1] = P.c_ub[j];
        c_coef[j - 1] = P.c_coef[j]; c_kind[j - 1] = P.c_kind[j];
        if (sjj) sjj[j - 1] = P.c_sjj[j];
        A_ptr[j - 1] = loc;
        for (const Elem &e : P.col_list[j]) { A_ind[loc] = e.idx - 1; A_val[loc] = e.val; loc++; }
    }
    A_ptr[P.n] = loc;
}

API void glpo_set_scale(void *h, const double *rii, const double *sjj)
{
    Prob &P = ((Handle *)h)->P;
    for (int i = 1; i <= P.m; i++) P.r_rii[i] = rii[i - 1];
    for (int j = 1; j <= P.n; j++) P.c_sjj[j] = sjj[j - 1];
    P.valid = 0;
}

API void glpo_std_basis(void *h) { prob_std_basis(((Handle *)h)->P); }

/* stat[0..m+n): rows then columns */
API void glpo_set_stat(void *h, const int *stat)
{
    Prob &P = ((Handle *)h)->P;
    for (int i = 1; i <= P.m; i++) prob_set_row_stat(P, i, stat[i - 1]);
    for (int j = 1; j <= P.n; j++) prob_set_col_stat(P, j, stat[P.m + j - 1]);
}

API void glpo_set_col_bnds(void *h, int j, int type, double lb, double ub)
{
    prob_set_col_bnds(((Handle *)h)->P, j, type, lb, ub);
}

API void glpo_set_row_bnds(void *h, int i, int type, double lb, double ub)
{
    prob_set_row_bnds(((Handle *)h)->P, i, type, lb, ub);
}

API void glpo_set_bfcp(void *h, int nfs_max, double piv_tol, int piv_lim, double upd_tol)
{
    Prob &P = ((Handle *)h)->P;
    delete P.bfd;
    P.bfd = new BFD();
    P.bfd->nfs_max = nfs_max; P.bfd->piv_tol = piv_tol; P.bfd->piv_lim = piv_lim;
    P.bfd->upd_tol = upd_tol;
    P.valid = 0;
}

struct glpo_smcp {
    int msg_lev, meth, pricing, r_test;
    double tol_bnd, tol_dj, tol_piv, obj_ll, obj_ul;
    int it_lim, tm_lim, out_frq, out_dly, presolve;
};

static SMCP to_smcp(const glpo_smcp *s)
{
    SMCP p;
    if (!s) return p;
    p.msg_lev = s->msg_lev; p.meth = s->meth; p.pricing = s->pricing; p.r_test = s->r_test;
    p.tol_bnd = s->tol_bnd; p.tol_dj = s->tol_dj; p.tol_piv = s->tol_piv;
    p.obj_ll = s->obj_ll; p.obj_ul = s->obj_ul;
    p.it_lim = s->it_lim; p.tm_lim = s->tm_lim; p.out_frq = s->out_frq; p.out_dly = s->out_dly;
    p.presolve = s->presolve;
    return p;
}

API void glpo_init_smcp(glpo_smcp *s)
{
    SMCP p;
    s->msg_lev = p.msg_lev; s->meth = p.meth; s->pricing = p.pricing; s->r_test = p.r_test;
    s->tol_bnd = p.tol_bnd; s->tol_dj = p.tol_dj; s->tol_piv = p.tol_piv;
    s->obj_ll = p.obj_ll; s->obj_ul = p.obj_ul;
    s->it_lim = p.it_lim; s->tm_lim = p.tm_lim; s->out_frq = p.out_frq; s->out_dly = p.out_dly;
    s->presolve = p.presolve;
}

API int glpo_factorize(void *h) { return prob_factorize(((Handle *)h)->P); }

API int glpo_simplex(void *h, const glpo_smcp *s)
{
    Handle *H = (Handle *)h;
    SMCP p = to_smcp(s);
    return simplex(H->P, p, H->hook.fn ? &H->hook : nullptr);
}

API void glpo_set_hook(void *h, hook_fn fn, void *user)
{
    Handle *H = (Handle *)h;
    H->hook.fn = fn; H->hook.user = user;
}

/* solution: stat/prim/dual are [m+n] rows-then-columns, head is [m] 1-based k */
API void glpo_get_solution(void *h, int *stat, double *prim, double *dual, int *head,
                           int *pbs, int *dbs, double *obj, int *it_cnt, int *some)
{
    Prob &P = ((Handle *)h)->P;
    for (int i = 1; i <= P.m; i++) {
        if (stat) stat[i - 1] = P.r_stat[i];
        if (prim) prim[i - 1] = P.r_prim[i];
        if (dual) dual[i - 1] = P.r_dual[i];
        if (head) head[i - 1] = P.head[i];
    }
    for (int j = 1; j <= P.n; j++) {
        if (stat) stat[P.m + j - 1] = P.c_stat[j];
        if (prim) prim[P.m + j - 1] = P.c_prim[j];
        if (dual) dual[P.m + j - 1] = P.c_dual[j];
    }
    if (pbs) *pbs = P.pbs_stat;
    if (dbs) *dbs = P.dbs_stat;
    if (obj) *obj = P.obj_val;
    if (it_cnt) *it_cnt = P.it_cnt;
    if (some) *some = P.some;
}

API int glpo_get_status(void *h) { return prob_get_status(((Handle *)h)->P); }

API void glpo_get_bfd_stats(void *h, long *out4)
{
    Prob &P = ((Handle *)h)->P;
    if (!P.bfd) { out4[0] = out4[1] = out4[2] = out4[3] = 0; return; }
    out4[0] = P.bfd->n_factorize; out4[1] = P.bfd->n_update;
    out4[2] = P.bfd->n_ftran; out4[3] = P.bfd->n_btran;
}

struct glpo_iocp {
    int msg_lev, br_tech, bt_tech;
    double tol_int, tol_obj;
    int tm_lim, out_frq, out_dly, pp_tech;
    double mip_gap;
    int presolve;
    long node_lim;
};

API void glpo_init_iocp(glpo_iocp *s)
{
    IOCP p;
    s->msg_lev = p.msg_lev; s->br_tech = p.br_tech; s->bt_tech = p.bt_tech;
    s->tol_int = p.tol_int; s->tol_obj = p.tol_obj; s->tm_lim = p.tm_lim;
    s->out_frq = p.out_frq; s->out_dly = p.out_dly; s->pp_tech = p.pp_tech;
    s->mip_gap = p.mip_gap; s->presolve = p.presolve; s->node_lim = p.node_lim;
}

API int glpo_intopt(void *h, const glpo_iocp *s)
{
    Handle *H = (Handle *)h;
    IOCP p;
    if (s) {
        p.msg_lev = s->msg_lev; p.br_tech = s->br_tech; p.bt_tech = s->bt_tech;
        p.tol_int = s->tol_int; p.tol_obj = s->tol_obj; p.tm_lim = s->tm_lim;
        p.out_frq = s->out_frq; p.out_dly = s->out_dly; p.pp_tech = s->pp_tech;
        p.mip_gap = s->mip_gap; p.presolve = s->presolve; p.node_lim = s->node_lim;
    }
    return intopt(H->P, p, &H->nodes);
}

API void glpo_get_mip(void *h, int *mip_stat, double *mip_obj, double *mipx, long *nodes)
{
    Handle *H = (Handle *)h;
    Prob &P = H->P;
    if (mip_stat) *mip_stat = P.mip_stat;
    if (mip_obj) *mip_obj = P.mip_obj;
    if (mipx) {
        for (int i = 1; i <= P.m; i++) mipx[i - 1] = P.r_mipx[i];
        for (int j = 1; j <= P.n; j++) mipx[P.m + j - 1] = P.c_mipx[j];
    }
    if (nodes) *nodes = H->nodes;
}

/* ---- live CSA access from inside a hook (CSA layout: 1-based arrays) ---- */
API void glpo_csa_scalars(void *csa, int *iv, double *dv)
{
    CSA &c = *(CSA *)csa;
    iv[0] = c.m; iv[1] = c.n; iv[2] = c.phase; iv[3] = c.p; iv[4] = c.q; iv[5] = c.p_stat;
    iv[6] = c.tcol_nnz; iv[7] = c.tcol_num; iv[8] = c.trow_nnz; iv[9] = c.trow_num;
    iv[10] = c.it_cnt; iv[11] = c.refct; iv[12] = (int)c.A_val.size() - 1;
    dv[0] = c.teta; dv[1] = c.delta; dv[2] = c.new_dq; dv[3] = c.zeta;
    dv[4] = c.tcol_max; dv[5] = c.trow_max; dv[6] = c.hook_tol;
}

template <class T> static int copy_out(const std::vector<T> &v, void *out, int cap_bytes)
{
    int bytes = (int)(v.size() * sizeof(T));
    if (out && bytes <= cap_bytes) memcpy(out, v.data(), bytes);
    return bytes;
}

/* copies the named CSA array (including slot 0) into out; returns its size
   in bytes, or -1 for an unknown name */
API int glpo_csa_get(void *csa, const char *name, void *out, int cap_bytes)
{
    CSA &c = *(CSA *)csa;
    std::string s(name);
#define F(x) if (s == #x) return copy_out(c.x, out, cap_bytes)
    F(type); F(lb); F(ub); F(coef); F(orig_type); F(orig_lb); F(orig_ub); F(obj);
    F(A_ptr); F(A_ind); F(A_val); F(AT_ptr); F(AT_ind); F(AT_val);
    F(head); F(bind); F(stat); F(N_ptr); F(N_len); F(N_ind); F(N_val);
    F(bbar); F(cbar); F(refsp); F(gamma);
    F(tcol_ind); F(tcol_vec); F(trow_ind); F(trow_vec);
    F(work1); F(work2); F(work3); F(work4);
#undef F
    return -1;
}

/* size of the factors the oracle's basis code holds right now (called from a
   hook): out = {nnz_f, nnz_v, nnz_h, hh_nfs, luf.n}; SURVEY 8d counts the bytes
   of FTRAN/BTRAN as 12*(nnz_f + nnz_h + nnz_v) + 16 m */
API void glpo_csa_lu_stats(void *csa, long *out5)
{
    CSA &c = *(CSA *)csa;
    out5[0] = out5[1] = out5[2] = out5[3] = out5[4] = 0;
    if (!c.bfd) return;
    out5[0] = c.bfd->luf.nnz_f; out5[1] = c.bfd->luf.nnz_v; out5[2] = c.bfd->nnz_h;
    out5[3] = c.bfd->hh_nfs; out5[4] = c.bfd->luf.n;
}

/* ---- stateless selection routines, CSA layout ---- */
API int glpo_chuzc_primal(int n, const signed char *stat, const double *cbar,
                          const double *gamma, double tol_dj)
{
    return o_chuzc_primal(n, stat, cbar, gamma, tol_dj);
}

API void glpo_sort_list(int *ind, const double *vec, int nnz, double tol, int *num, double *vmax)
{
    o_sort_list(ind, vec, nnz, tol, num, vmax);
}

API void glpo_chuzr_primal(int m, const signed char *type, const double *lb, const double *ub,
                           const double *coef, const int *head, int phase, const double *bbar,
                           double cbar_q, int q, const int *tcol_ind, const double *tcol_vec,
                           int tcol_num, double rtol, int *p, int *p_stat, double *teta)
{
    o_chuzr_primal(m, type, lb, ub, coef, head, phase, bbar, cbar_q, q, tcol_ind, tcol_vec,
                   tcol_num, rtol, p, p_stat, teta);
}

API int glpo_chuzr_dual(int m, const signed char *type, const double *lb, const double *ub,
                        const int *head, const double *bbar, const double *gamma,
                        double tol_bnd, double *delta)
{
    return o_chuzr_dual(m, type, lb, ub, head, bbar, gamma, tol_bnd, delta);
}

API void glpo_chuzc_dual(const signed char *stat, const double *cbar, double delta,
                         const int *trow_ind, const double *trow_vec, int trow_num,
                         double rtol, int *q, double *new_dq)
{
    o_chuzc_dual(stat, cbar, delta, trow_ind, trow_vec, trow_num, rtol, q, new_dq);
}

/* RNG cross-check for the product's synthetic generators */
API void glpo_rng_fill(int seed, int count, int *out)
{
    RNG r;
    rng_init(r, seed);
    for (int i = 0; i < count; i++) out[i] = rng_next(r);
}
