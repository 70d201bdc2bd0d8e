/* nodeengine.cuh -- ONE BRANCH-AND-BOUND NODE PER CTA ("many
 * node LPs concurrently, one CTA per node, LP resident in shared memory").
 *
 * A CTA carries one open node through everything the reference does between
 * picking it and freezing it, i.e. the whole `more:` state of ios_driver
 * (lib/glpios03.js:615-905):
 *
 *   ios_preprocess_node        lib/glpios02.js           ne_preprocess
 *   ios_solve_node             lib/glpios01.js:866-910   ne_solve_lp
 *     = glp_simplex(DUALP): glp_factorize + spx_dual     (lib/glpspx02.js:1592-1966:
 *       eval_cbar, check_feas/check_stab, set_aux/orig_bnds, eval_bbar, chuzr,
 *       eval_rho, eval_trow, sort_trow, Harris chuzc, eval_tcol, update_bbar/
 *       cbar/gamma, basis change, objective cut-off, store_sol)
 *   ios_round_bound            lib/glpios01.js:730-787   ne_round_prepare/ne_round_bound
 *   check_integrality          lib/glpios03.js:56-116
 *   fix_by_red_cost            lib/glpios03.js:307-377
 *   branch_drtom / branch_mostf / first / last           lib/glpios09.js:28-270
 *     = glp_eval_tab_row + 2 x glp_dual_rtest per fractional column
 *       (lib/glpapi12.js:401-453, 687-762)
 *   branch_on + ios_eval_degrad lib/glpios03.js:141-305, lib/glpios01.js:615-728
 *
 * and leaves either a verdict (fathomed / integral) or the two children, each
 * a complete node state (type, lb, ub, stat of all m+n variables) written to
 * the device-resident node slab.  The host keeps only the tree (bounds and
 * slab numbers).  The dense scaled matrix lives in shared memory when it fits
 * (30 x 500 doubles = 120 KB for BASELINE.json configs[4]), the basis inverse
 * is explicit (m x m, shared memory): every solve is a short chain of dense
 * m x n passes and block reductions, no global traffic but the node state.
 *
 * The file is plain C++ apart from the NE_* macros: with -DNE_EMUL it compiles
 * for the host with one "thread" per CTA, which is how tests/ exercise the
 * node logic without a GPU (test infrastructure, never part of the product).
 *
 * Ties: every arg-reduction carries the element index and prefers the lowest
 * one, like the sequential loops of the reference in ascending index order.
 */
#ifndef GLPB_NODEENGINE_CUH
#define GLPB_NODEENGINE_CUH

#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>

#ifdef NE_EMUL
#define NE_D static inline
#define NE_HD static inline
#define NE_M inline
#define NE_SYNC() ((void)0)
#else
#define NE_D __device__ __forceinline__
#define NE_HD __host__ __device__ __forceinline__
#define NE_M __device__ __forceinline__
#define NE_SYNC() __syncthreads()
#endif

#ifdef NE_EMUL
#include <cstdio>
#include <cstdlib>
#define NE_TRACE(...) do { static int on_ = -1; if (on_ < 0) on_ = getenv("NE_TRACE") ? 1 : 0; if (on_) fprintf(stderr, __VA_ARGS__); } while (0)
#else
#define NE_TRACE(...) ((void)0)
#endif

#define NE_MAXM 64
#define NE_MAXN 1024
#define NE_NT 256          /* threads per CTA */

enum { NE_FR = 1, NE_LO = 2, NE_UP = 3, NE_DB = 4, NE_FX = 5 };
enum { NE_BS = 1, NE_NL = 2, NE_NU = 3, NE_NF = 4, NE_NS = 5 };
enum { NE_UNDEF = 1, NE_FEAS = 2, NE_INFEAS = 3, NE_NOFEAS = 4 };
enum { NE_MIN = 1, NE_MAX = 2 };
enum { NE_BR_FFV = 1, NE_BR_LFV = 2, NE_BR_MFV = 3, NE_BR_DTH = 4 };
enum { NE_PP_NONE = 0, NE_PP_ROOT = 1, NE_PP_ALL = 2 };
enum { NE_EFAIL = 5, NE_EOBJLL = 6, NE_EOBJUL = 7, NE_EITLIM = 8 };

/* verdict of one task */
enum {
    NE_R_FATHOM = 1,     /* infeasible, cut off or hopeless: delete the node        */
    NE_R_INTEGRAL = 2,   /* LP optimum is integer feasible: candidate incumbent      */
    NE_R_BRANCH = 3,     /* two children written (slab[node] = down, slab[child] = up) */
    NE_R_FAIL = 4,       /* the LP could not be solved (GLP_EFAIL of ios_driver)     */
};
enum { NE_NO_BRNCH = 0, NE_DN_BRNCH = 1, NE_UP_BRNCH = 2 };

/* static problem data, shared by all nodes (device global memory) */
struct NeProb {
    int m, n, lda, ldb;         /* lda: row stride of As; ldb: row stride of the inverse */
    int dir, a_in_smem;
    double c0, zeta;
    const double *As;           /* [m][lda] scaled matrix rii*a*sjj, row-major          */
    const double *cw;           /* [n] working costs coef*sjj*zeta (lib/glpspx02.js:148) */
    const double *obj;          /* [n] scaled costs coef*sjj                             */
    const double *ucoef;        /* [n] costs as given                                    */
    const double *rii, *sjj;    /* [m], [n]                                              */
    const signed char *kind;    /* [n] 1 = integer column                                */
    double tol_bnd, tol_dj, tol_piv, tol_int, tol_obj;
    int pp_tech, br_tech, it_max, refac_period;
    int unit_scale, pad1;       /* 1: rii = sjj = 1 (no division in the preprocessing) */
    /* node slab */
    double *slab_lb, *slab_ub;          /* [cap][mn]  */
    signed char *slab_type, *slab_stat; /* [cap][mn]  */
    double *slab_bi;                    /* [cap][m*ldb] inverse of the node's starting basis (= the parent's final one) */
    int *slab_upd;                      /* [cap] updates that inverse has seen since it was computed afresh; -1: none stored */
    int *slab_head;                     /* [cap][m] basis header (variable per position) the stored inverse belongs to */
    /* incumbent shared by the CTAs of a launch: inc[0] objective, inc_have flag */
    double *inc;
    int *inc_have;
};

struct NeTask {
    int node, child, level, pad;
    double bound, lp_obj;
};

struct NeResult {
    int code, jv, next, ii_cnt;
    int iters, solves, refacs, ret;
    double obj, bound, ii_sum;
    double dn_lp, dn_bnd, up_lp, up_bnd;
    double x_jv;
    long long cyc[6];    /* SM cycles of thread 0: preprocessing, basis set-up, fresh vectors, iterations, branching, state I/O */
};

struct NeT { int tid, nt, lane, warp, nwarps, wsize; mutable int rc; };

NE_D long long ne_clock()
{
#ifdef NE_EMUL
    return 0;
#else
    return clock64();
#endif
}

/* ------------------------------------------------------------------ */
/* reductions                                                         */
/* ------------------------------------------------------------------ */
struct NeKey { double a, b; int idx, aux; };

#ifndef NE_EMUL
template <class K> __device__ __forceinline__ K ne_shfl_down(const K &v, int off)
{
    static_assert(sizeof(K) % 4 == 0, "key size");
    K r;
    const unsigned *s = reinterpret_cast<const unsigned *>(&v);
    unsigned *d = reinterpret_cast<unsigned *>(&r);
#pragma unroll
    for (int w = 0; w < (int)(sizeof(K) / 4); w++) d[w] = __shfl_down_sync(0xffffffffu, s[w], off);
    return r;
}
template <class K> __device__ __forceinline__ K ne_shfl_idx(const K &v, int src)
{
    K r;
    const unsigned *s = reinterpret_cast<const unsigned *>(&v);
    unsigned *d = reinterpret_cast<unsigned *>(&r);
#pragma unroll
    for (int w = 0; w < (int)(sizeof(K) / 4); w++) d[w] = __shfl_sync(0xffffffffu, s[w], src);
    return r;
}
#endif

/* op(a, b): a comes from the lower thread index.  All threads of the CTA
   must call; every thread gets the same result.  scratch: >= nwarps keys. */
template <class K, class Op> NE_D K ne_block_reduce(const NeT &t, K v, Op op, K *scratch)
{
#ifdef NE_EMUL
    (void)t; (void)op; (void)scratch;
    return v;
#else
    for (int off = 16; off; off >>= 1) {
        K o = ne_shfl_down(v, off);
        if (t.lane + off < 32) v = op(v, o);
    }
    /* two scratch halves used alternately: a half is rewritten only after the barrier of
       the next reduction, which every thread passes after it has read this one */
    K *buf = (K *)((unsigned char *)scratch + (t.rc & 1) * 16 * 32);
    t.rc++;
    if (t.lane == 0) buf[t.warp] = v;
    __syncthreads();
    K r = buf[0];
    for (int w = 1; w < t.nwarps; w++) r = op(r, buf[w]);
    return r;
#endif
}

/* block-wide OR of a predicate: one barrier */
NE_D int ne_block_or(const NeT &t, int v)
{
#ifdef NE_EMUL
    (void)t;
    return v;
#else
    return __syncthreads_or(v);
#endif
}

template <class K, class Op> NE_D K ne_warp_reduce(const NeT &t, K v, Op op)
{
#ifdef NE_EMUL
    (void)t; (void)op;
    return v;
#else
    for (int off = 16; off; off >>= 1) {
        K o = ne_shfl_down(v, off);
        if (t.lane + off < 32) v = op(v, o);
    }
    return ne_shfl_idx(v, 0);
#endif
}

struct NeSum { NE_M double operator()(double a, double b) const { return a + b; } };
struct NeMax { NE_M double operator()(double a, double b) const { return a >= b ? a : b; } };
struct NeOrI { NE_M int operator()(int a, int b) const { return a | b; } };
struct NeAddI { NE_M int operator()(int a, int b) const { return a + b; } };

/* ------------------------------------------------------------------ */
/* per-CTA state (pointers into shared memory)                        */
/* ------------------------------------------------------------------ */
struct NeS {
    const double *A;            /* [m][lda] */
    double *Bi;                 /* [m][ldb] explicit inverse of the basis           */
    double *lb, *ub;            /* [mn] node bounds as the API sees them (unscaled) */
    double *sl, *su;            /* [mn] working bounds of the simplex (scaled)      */
    double *prim, *dual;        /* [mn] solution as store_sol leaves it (unscaled)  */
    double *cbar, *trow;        /* [n]  */
    double *bbar, *gamma, *rho, *tcol, *u, *w;   /* [m] */
    double *Lr, *Ur;            /* [m+1] row bounds of the preprocessing            */
    double *dz;                 /* [4][m] Driebeck-Tomlin estimates per fractional column */
    int *head, *bind;           /* [mn] */
    int *frac;                  /* [m+1] fractional basic structurals               */
    int *list;                  /* [m+2] work list of the preprocessing             */
    signed char *type, *wtype, *kstat, *refsp;   /* [mn] */
    signed char *nstat;         /* [n] status by non-basic position                 */
    signed char *flag;          /* [max(n, m+1)] scratch flags                      */
    signed char *mark, *pass;   /* [m+1] */
    NeKey *red;                 /* 2 x 16 x 32 bytes: reduction scratch, two halves (<= 16 warps, keys <= 32 B) */
    double *sc;                 /* [16] broadcast scalars                           */
    int *si;                    /* [16] broadcast ints                              */
};

NE_HD size_t ne_align(size_t x) { return (x + 15) & ~(size_t)15; }

/* bytes of shared memory the state needs (without the matrix) */
NE_HD size_t ne_state_bytes(int m, int n, int ldb)
{
    const int mn = m + n;
    size_t b = 0;
    b += ne_align((size_t)m * ldb * 8);
    b += 6 * ne_align((size_t)mn * 8);
    b += 2 * ne_align((size_t)n * 8);
    b += 6 * ne_align((size_t)m * 8);
    b += 2 * ne_align((size_t)(m + 1) * 8);
    b += ne_align((size_t)4 * m * 8);
    b += 2 * ne_align((size_t)mn * 4);
    b += ne_align((size_t)(m + 1) * 4) + ne_align((size_t)(m + 2) * 4);
    b += 4 * ne_align((size_t)mn);
    b += ne_align((size_t)n);
    b += ne_align((size_t)(n > m + 1 ? n : m + 1));
    b += 2 * ne_align((size_t)(m + 1));
    b += ne_align(2 * 16 * 32);
    b += ne_align(16 * 8) + ne_align(16 * 4);
    return b;
}

NE_D void ne_carve(NeS &S, unsigned char *p, int m, int n, int ldb)
{
    const int mn = m + n;
    auto take = [&](size_t bytes) { unsigned char *r = p; p += ne_align(bytes); return r; };
    S.Bi = (double *)take((size_t)m * ldb * 8);
    S.lb = (double *)take((size_t)mn * 8); S.ub = (double *)take((size_t)mn * 8);
    S.sl = (double *)take((size_t)mn * 8); S.su = (double *)take((size_t)mn * 8);
    S.prim = (double *)take((size_t)mn * 8); S.dual = (double *)take((size_t)mn * 8);
    S.cbar = (double *)take((size_t)n * 8); S.trow = (double *)take((size_t)n * 8);
    S.bbar = (double *)take((size_t)m * 8); S.gamma = (double *)take((size_t)m * 8);
    S.rho = (double *)take((size_t)m * 8); S.tcol = (double *)take((size_t)m * 8);
    S.u = (double *)take((size_t)m * 8); S.w = (double *)take((size_t)m * 8);
    S.Lr = (double *)take((size_t)(m + 1) * 8); S.Ur = (double *)take((size_t)(m + 1) * 8);
    S.dz = (double *)take((size_t)4 * m * 8);
    S.head = (int *)take((size_t)mn * 4); S.bind = (int *)take((size_t)mn * 4);
    S.frac = (int *)take((size_t)(m + 1) * 4); S.list = (int *)take((size_t)(m + 2) * 4);
    S.type = (signed char *)take(mn); S.wtype = (signed char *)take(mn);
    S.kstat = (signed char *)take(mn); S.refsp = (signed char *)take(mn);
    S.nstat = (signed char *)take(n);
    S.flag = (signed char *)take(n > m + 1 ? n : m + 1);
    S.mark = (signed char *)take(m + 1); S.pass = (signed char *)take(m + 1);
    S.red = (NeKey *)take(2 * 16 * 32);
    S.sc = (double *)take(16 * 8); S.si = (int *)take(16 * 4);
}

/* sum_i x[i] * A[i][col] over the m rows of the dense matrix (row stride lda) */
NE_D double ne_coldot(const double *A, int lda, int m, int col, const double *x)
{
    const double *a = A + col;
    double s0 = 0.0, s1 = 0.0;
    int i = 0;
    for (; i + 1 < m; i += 2) { s0 += x[i] * a[(size_t)i * lda]; s1 += x[i + 1] * a[(size_t)(i + 1) * lda]; }
    if (i < m) s0 += x[i] * a[(size_t)i * lda];
    return s0 + s1;
}

/* scale of variable k: unscaled value = scaled value * ne_scale(k)
   (rows: 1/rii, columns: sjj; lib/glpspx01.js:61-75, 1629-1678) */
NE_D double ne_scale(const NeProb &P, int k) { return k < P.m ? 1.0 / P.rii[k] : P.sjj[k - P.m]; }

/* glp_set_row_bnds / glp_set_col_bnds: lib/glpapi01.js:217-281 */
NE_D void ne_set_bnds(NeS &S, int k, int type, double l, double u)
{
    switch (type) {
    case NE_FR: l = u = 0.0; break;
    case NE_LO: u = 0.0; break;
    case NE_UP: l = 0.0; break;
    case NE_FX: u = l; break;
    default: break;
    }
    S.type[k] = (signed char)type; S.lb[k] = l; S.ub[k] = u;
    int st = S.kstat[k];
    if (st != NE_BS) {
        switch (type) {
        case NE_FR: st = NE_NF; break;
        case NE_LO: st = NE_NL; break;
        case NE_UP: st = NE_NU; break;
        case NE_DB: if (!(st == NE_NL || st == NE_NU)) st = (fabs(l) <= fabs(u) ? NE_NL : NE_NU); break;
        default: st = NE_NS;
        }
        S.kstat[k] = (signed char)st;
    }
}

/* glp_get_row_lb/ub, glp_get_col_lb/ub: -DBL_MAX / +DBL_MAX when absent */
NE_D double ne_get_lb(const NeS &S, int k)
{
    int t = S.type[k];
    return (t == NE_FR || t == NE_UP) ? -DBL_MAX : S.lb[k];
}
NE_D double ne_get_ub(const NeS &S, int k)
{
    int t = S.type[k];
    return (t == NE_FR || t == NE_LO) ? +DBL_MAX : (t == NE_FX ? S.lb[k] : S.ub[k]);
}

/* ------------------------------------------------------------------ */
/* node preprocessing: lib/glpios02.js                                */
/* ------------------------------------------------------------------ */
struct NeRowAcc { double fmin, fmax; int nmin, nmax, jmin, jmax; };
struct NeRowAccOp {
    NE_M NeRowAcc operator()(const NeRowAcc &a, const NeRowAcc &b) const
    {
        NeRowAcc r;
        r.fmin = a.fmin + b.fmin; r.fmax = a.fmax + b.fmax;
        r.nmin = a.nmin + b.nmin; r.nmax = a.nmax + b.nmax;
        r.jmin = a.jmin < b.jmin ? a.jmin : b.jmin;
        r.jmax = a.jmax < b.jmax ? a.jmax : b.jmax;
        return r;
    }
};

/* returns 1 if the node is proven infeasible.  l/u (column bounds, +-DBL_MAX
   when absent) live in S.cbar / S.trow for the duration of the call. */
NE_D int ne_preprocess(const NeT &t, const NeProb &P, NeS &S, int max_pass, int have_inc, double mip_obj)
{
    const int m = P.m, n = P.n;
    double *l = S.cbar, *u = S.trow, *L = S.Lr, *U = S.Ur;
    for (int i = t.tid; i <= m; i += t.nt) {
        if (i == 0) {
            if (have_inc) {
                if (P.dir == NE_MIN) { L[0] = -DBL_MAX; U[0] = mip_obj - P.c0; }
                else { L[0] = mip_obj - P.c0; U[0] = +DBL_MAX; }
            } else { L[0] = -DBL_MAX; U[0] = +DBL_MAX; }
        } else { L[i] = ne_get_lb(S, i - 1); U[i] = ne_get_ub(S, i - 1); }
        S.mark[i] = 1; S.pass[i] = 0;
        S.list[i] = i;
    }
    for (int j = t.tid; j < n; j += t.nt) { l[j] = ne_get_lb(S, m + j); u[j] = ne_get_ub(S, m + j); }
    NE_SYNC();
    int size = m + 1;
    int infeasible = 0;
    while (size > 0) {
        const int i = S.list[size - 1];
        size--;
        if (t.tid == 0) { S.mark[i] = 0; S.pass[i]++; }       /* only thread 0 reads mark / pass */
        const double Li0 = L[i], Ui0 = U[i];
        if (Li0 == -DBL_MAX && Ui0 == +DBL_MAX) continue;
        for (int r = t.tid; r <= m; r += t.nt) S.flag[r] = 0;   /* visible after the barrier of the reduction below */
        /* prepare_row_info: sums over the finite terms, number and first
           index of the infinite ones */
        NeRowAcc acc = {0.0, 0.0, 0, 0, INT_MAX, INT_MAX};
        for (int j = t.tid; j < n; j += t.nt) {
            double a = (i == 0) ? P.ucoef[j] : S.A[(size_t)(i - 1) * P.lda + j];
            if (a == 0.0) continue;
            if (i != 0 && !P.unit_scale) a /= (P.rii[i - 1] * P.sjj[j]);
            double bmin = a > 0.0 ? l[j] : u[j], bmax = a > 0.0 ? u[j] : l[j];
            if (bmin == (a > 0.0 ? -DBL_MAX : +DBL_MAX)) { acc.nmin++; if (j < acc.jmin) acc.jmin = j; }
            else acc.fmin += a * bmin;
            if (bmax == (a > 0.0 ? +DBL_MAX : -DBL_MAX)) { acc.nmax++; if (j < acc.jmax) acc.jmax = j; }
            else acc.fmax += a * bmax;
        }
        acc = ne_block_reduce(t, acc, NeRowAccOp(), (NeRowAcc *)S.red);
        double f_min = acc.fmin, f_max = acc.fmax;
        int j_min = -1, j_max = -1;
        if (acc.nmin == 1) j_min = acc.jmin; else if (acc.nmin > 1) f_min = -DBL_MAX;
        if (acc.nmax == 1) j_max = acc.jmax; else if (acc.nmax > 1) f_max = +DBL_MAX;
        /* check_row_bounds */
        double Li = Li0, Ui = Ui0;
        {
            double LL = (j_min < 0 ? f_min : -DBL_MAX), UU = (j_max < 0 ? f_max : +DBL_MAX);
            if (Li != -DBL_MAX && UU < Li - 1e-3 * (1.0 + fabs(Li))) { infeasible = 1; break; }
            if (Ui != +DBL_MAX && LL > Ui + 1e-3 * (1.0 + fabs(Ui))) { infeasible = 1; break; }
            if (Li != -DBL_MAX && LL > Li - 1e-12 * (1.0 + fabs(Li))) Li = -DBL_MAX;
            if (Ui != +DBL_MAX && UU < Ui + 1e-12 * (1.0 + fabs(Ui))) Ui = +DBL_MAX;
        }
        if (t.tid == 0) { L[i] = Li; U[i] = Ui; }
        if (Li == -DBL_MAX && Ui == +DBL_MAX) { NE_SYNC(); continue; }
        /* columns of the row: check_col_bounds + check_efficiency */
        int bad = 0;
        for (int j = t.tid; j < n; j += t.nt) {
            double a = (i == 0) ? P.ucoef[j] : S.A[(size_t)(i - 1) * P.lda + j];
            if (a == 0.0) continue;
            if (i != 0 && !P.unit_scale) a /= (P.rii[i - 1] * P.sjj[j]);
            const double lj0 = l[j], uj0 = u[j];
            double ilb, iub, ll, uu;
            if (Li == -DBL_MAX || f_max == +DBL_MAX) ilb = -DBL_MAX;
            else if (j_max < 0) ilb = Li - (f_max - a * (a > 0.0 ? uj0 : lj0));
            else if (j_max == j) ilb = Li - f_max;
            else ilb = -DBL_MAX;
            if (Ui == +DBL_MAX || f_min == -DBL_MAX) iub = +DBL_MAX;
            else if (j_min < 0) iub = Ui - (f_min - a * (a > 0.0 ? lj0 : uj0));
            else if (j_min == j) iub = Ui - f_min;
            else iub = +DBL_MAX;
            if (fabs(a) < 1e-6) { ll = -DBL_MAX; uu = +DBL_MAX; }
            else if (a > 0.0) { ll = (ilb == -DBL_MAX ? -DBL_MAX : ilb / a); uu = (iub == +DBL_MAX ? +DBL_MAX : iub / a); }
            else { ll = (iub == +DBL_MAX ? -DBL_MAX : iub / a); uu = (ilb == -DBL_MAX ? +DBL_MAX : ilb / a); }
            const int isint = P.kind[j] != 0;
            if (isint) {
                if (ll != -DBL_MAX) ll = (ll - floor(ll) < 1e-3 ? floor(ll) : ceil(ll));
                if (uu != +DBL_MAX) uu = (ceil(uu) - uu < 1e-3 ? ceil(uu) : floor(uu));
            }
            double lj = lj0, uj = uj0;
            if (lj != -DBL_MAX && uu < lj - 1e-3 * (1.0 + fabs(lj))) { bad = 1; continue; }
            if (uj != +DBL_MAX && ll > uj + 1e-3 * (1.0 + fabs(uj))) { bad = 1; continue; }
            if (ll != -DBL_MAX && lj < ll - 1e-3 * (1.0 + fabs(ll))) lj = ll;
            if (uu != +DBL_MAX && uj > uu + 1e-3 * (1.0 + fabs(uu))) uj = uu;
            if (!(lj == -DBL_MAX || uj == +DBL_MAX)) {
                double t1 = fabs(lj), t2 = fabs(uj);
                double eps = 1e-10 * (1.0 + (t1 <= t2 ? t1 : t2));
                if (lj > uj - eps) {
                    if (lj == lj0) uj = lj;
                    else if (uj == uj0) lj = uj;
                    else if (t1 <= t2) uj = lj;
                    else lj = uj;
                }
            }
            int eff = 0;
            if (lj0 < lj) {
                if (isint || lj0 == -DBL_MAX) eff++;
                else if (lj - lj0 >= 0.25 * ((uj0 == +DBL_MAX) ? 1.0 + fabs(lj0) : 1.0 + (uj0 - lj0))) eff++;
            }
            if (uj0 > uj) {
                if (isint || uj0 == +DBL_MAX) eff++;
                else if (uj0 - uj >= 0.25 * ((lj0 == -DBL_MAX) ? 1.0 + fabs(uj0) : 1.0 + (uj0 - lj0))) eff++;
            }
            l[j] = lj; u[j] = uj;
            if (eff > 0)
                for (int r = 0; r < m; r++)
                    if (S.A[(size_t)r * P.lda + j] != 0.0) S.flag[r + 1] = 1;   /* benign race: all write 1 */
        }
        bad = ne_block_or(t, bad);
        if (bad) { infeasible = 1; break; }
        /* rows touched by an efficient change go back on the list (ascending) */
        if (t.tid == 0) {
            int sz = size;
            for (int r = 1; r <= m; r++) {
                if (!S.flag[r]) continue;
                if (S.pass[r] >= max_pass) continue;
                if (L[r] == -DBL_MAX && U[r] == +DBL_MAX) continue;
                if (!S.mark[r]) { S.list[sz++] = r; S.mark[r] = 1; }
            }
            S.si[0] = sz;
        }
        NE_SYNC();
        size = S.si[0];          /* rewritten only after the barriers of the next row */
    }
    NE_SYNC();
    if (infeasible) return 1;
    /* relaxed row bounds (basic rows only), tightened column bounds */
    for (int i = t.tid; i < m; i += t.nt)
        if (S.kstat[i] == NE_BS) {
            double Li = L[1 + i], Ui = U[1 + i];
            if (Li == -DBL_MAX && Ui == +DBL_MAX) ne_set_bnds(S, i, NE_FR, 0.0, 0.0);
            else if (Ui == +DBL_MAX) ne_set_bnds(S, i, NE_LO, Li, 0.0);
            else if (Li == -DBL_MAX) ne_set_bnds(S, i, NE_UP, 0.0, Ui);
        }
    for (int j = t.tid; j < n; j += t.nt) {
        int type;
        if (l[j] == -DBL_MAX && u[j] == +DBL_MAX) type = NE_FR;
        else if (u[j] == +DBL_MAX) type = NE_LO;
        else if (l[j] == -DBL_MAX) type = NE_UP;
        else if (l[j] != u[j]) type = NE_DB;
        else type = NE_FX;
        ne_set_bnds(S, m + j, type, l[j], u[j]);
    }
    NE_SYNC();
    return 0;
}

/* ------------------------------------------------------------------ */
/* dual simplex on the node LP                                        */
/* ------------------------------------------------------------------ */

/* basis header from the statuses: basic variables in ascending k
   (glp_factorize, lib/glpapi12.js:44-67), non-basic rows then columns
   (init_csa, lib/glpspx02.js:160-175).  Returns 0, or 1 if the number of
   basic variables is not m (GLP_EBADB). */
NE_D int ne_build_head(const NeT &t, const NeProb &P, NeS &S, int keep_basic_order)
{
    const int m = P.m, n = P.n, mn = m + n;
    /* keep_basic_order: S.head[0..m) already lists the basic variables in the order the
       inherited inverse was built for (the statuses must agree with it) */
    if (keep_basic_order) {
        int bad = 0;
        for (int i = t.tid; i < m; i += t.nt) { int k = S.head[i]; if (k < 0 || k >= mn || S.kstat[k] != NE_BS) bad = 1; else S.bind[k] = i; }
        if (ne_block_or(t, bad)) return 1;
    }
#ifdef NE_EMUL
    int nb = 0, nn = 0;
    for (int k = 0; k < mn; k++) {
        if (S.kstat[k] == NE_BS) { if (!keep_basic_order && nb < m) { S.head[nb] = k; S.bind[k] = nb; } nb++; }
        else { if (nn < n) { S.head[m + nn] = k; S.bind[k] = m + nn; S.nstat[nn] = S.kstat[k]; } nn++; }
    }
    return (nb == m) ? 0 : 1;
#else
    if (t.warp == 0) {           /* ordered compaction of both lists by ballots */
        int nb = 0, nn = 0;
        for (int base = 0; base < mn; base += 32) {
            const int k = base + t.lane;
            const int st = (k < mn) ? S.kstat[k] : 0;
            const unsigned mb = __ballot_sync(0xffffffffu, st == NE_BS);
            const unsigned mnb = __ballot_sync(0xffffffffu, st != NE_BS && k < mn);
            const unsigned lt = (1u << t.lane) - 1u;
            if (k < mn) {
                if (st == NE_BS) { int o = nb + __popc(mb & lt); if (o < m && !keep_basic_order) { S.head[o] = k; S.bind[k] = o; } }
                else { int o = nn + __popc(mnb & lt); if (o < n) { S.head[m + o] = k; S.bind[k] = m + o; S.nstat[o] = (signed char)st; } }
            }
            nb += __popc(mb); nn += __popc(mnb);
        }
        if (t.lane == 0) S.si[0] = (nb == m) ? 0 : 1;
    }
    __syncthreads();
    return S.si[0];              /* si[0] is next written after at least one more barrier */
#endif
}

/* scaled working bounds from the node bounds (lib/glpspx02.js:112-131) */
NE_D void ne_orig_bounds(const NeT &t, const NeProb &P, NeS &S)
{
    const int m = P.m, mn = P.m + P.n;
    for (int k = t.tid; k < mn; k += t.nt) {
        S.wtype[k] = S.type[k];
        if (k < m) { S.sl[k] = S.lb[k] * P.rii[k]; S.su[k] = S.ub[k] * P.rii[k]; }
        else { S.sl[k] = S.lb[k] / P.sjj[k - m]; S.su[k] = S.ub[k] / P.sjj[k - m]; }
    }
}

struct NePiv { double v; int idx, pad; };
struct NePivOp {
    NE_M NePiv operator()(const NePiv &a, const NePiv &b) const
    {
        if (b.v > a.v || (b.v == a.v && b.idx < a.idx)) return b;
        return a;
    }
};

/* explicit inverse of B = columns head[0..m) of (I | -A) by Gauss-Jordan with
   partial pivoting (replaces luf_factorize for an m x m basis that lives in
   shared memory).  Returns 1 if singular (lib/glpapi12.js:86-95 GLP_ESING). */
NE_D int ne_invert(const NeT &t, const NeProb &P, NeS &S)
{
    const int m = P.m, ldb = P.ldb;
    double *X = S.Bi;
    /* X = B (row r, column = basis position c) */
    for (int e = t.tid; e < m * m; e += t.nt) {
        int r = e / m, c = e - r * m;
        int k = S.head[c];
        X[r * ldb + c] = (k < m) ? (k == r ? 1.0 : 0.0) : -S.A[(size_t)r * P.lda + (k - m)];
    }
    int *perm = S.list;              /* row swapped with c at step c (m <= NE_MAXM, list has m+2) */
    NE_SYNC();
    for (int c = 0; c < m; c++) {
        /* pivot: largest |X[r][c]|, r >= c, lowest r on ties */
        NePiv best = {-1.0, INT_MAX, 0};
        for (int r = c + t.tid; r < m; r += t.nt) {
            double v = fabs(X[r * ldb + c]);
            if (v > best.v) { best.v = v; best.idx = r; }
        }
        best = ne_block_reduce(t, best, NePivOp(), (NePiv *)S.red);
        if (!(best.v > 1e-13)) return 1;
        const int pr = best.idx;
        if (t.tid == 0) perm[c] = pr;
        if (pr != c) {
            for (int j = t.tid; j < m; j += t.nt) {
                double a = X[c * ldb + j];
                X[c * ldb + j] = X[pr * ldb + j];
                X[pr * ldb + j] = a;
            }
            NE_SYNC();
        }
        const double piv = X[c * ldb + c];
        NE_SYNC();
        /* row c := row c / piv with X[c][c] := 1/piv; column c of the other rows kept as factor */
        for (int j = t.tid; j < m; j += t.nt) X[c * ldb + j] = (j == c) ? 1.0 / piv : X[c * ldb + j] / piv;
        NE_SYNC();
        for (int e = t.tid; e < m * m; e += t.nt) {
            int r = e / m, j = e - r * m;
            if (r == c) continue;
            double f = X[r * ldb + c];
            if (f == 0.0) continue;
            if (j != c) X[r * ldb + j] -= f * X[c * ldb + j];
        }
        NE_SYNC();
        for (int r = t.tid; r < m; r += t.nt)
            if (r != c) X[r * ldb + c] = -X[r * ldb + c] * X[c * ldb + c];
        NE_SYNC();
    }
    /* undo the row interchanges: inv(B) = inv(P B) P -> swap COLUMNS in reverse order */
    for (int c = m - 1; c >= 0; c--) {
        const int pr = perm[c];
        if (pr != c) {
            for (int r = t.tid; r < m; r += t.nt) {
                double a = X[r * ldb + c];
                X[r * ldb + c] = X[r * ldb + pr];
                X[r * ldb + pr] = a;
            }
            NE_SYNC();
        }
    }
    /* X now maps constraint rows -> basis positions: X[pos][row] */
    NE_SYNC();
    return 0;
}

/* get_xN: lib/glpspx01.js:442-471 (scaled) */
NE_D double ne_xN(const NeS &S, int m, int j)
{
    int k = S.head[m + j];
    switch (S.nstat[j]) {
    case NE_NL: return S.sl[k];
    case NE_NU: return S.su[k];
    case NE_NF: return 0.0;
    default: return S.sl[k];
    }
}

/* x := inv(B) h  (bfd_ftran): x[pos] = sum_r Bi[pos][r] h[r] */
NE_D void ne_ftran(const NeT &t, const NeProb &P, const NeS &S, const double *h, double *x)
{
    const int m = P.m, ldb = P.ldb;
    for (int pos = t.tid; pos < m; pos += t.nt) {
        double s = 0.0;
        for (int r = 0; r < m; r++) s += S.Bi[pos * ldb + r] * h[r];
        x[pos] = s;
    }
}

/* eval_bbar: beta = inv(B)(-N xN), lib/glpspx01.js:473-512.  Rows are summed
   by warps (lanes stride the non-basic columns). */
NE_D void ne_eval_bbar(const NeT &t, const NeProb &P, NeS &S)
{
    const int m = P.m, n = P.n;
    double *xn = S.trow, *h = S.w;
    for (int j = t.tid; j < n; j += t.nt) xn[j] = ne_xN(S, m, j);
    NE_SYNC();
    for (int i = t.warp; i < m; i += t.nwarps) {
        double s = 0.0;
        for (int j = t.lane; j < n; j += t.wsize) {
            int k = S.head[m + j];
            double x = xn[j];
            if (x == 0.0) continue;
            if (k < m) { if (k == i) s -= x; }
            else s += S.A[(size_t)i * P.lda + (k - m)] * x;
        }
        s = ne_warp_reduce(t, s, NeSum());
        if (t.lane == 0) h[i] = s;
    }
    NE_SYNC();
    ne_ftran(t, P, S, h, S.bbar);
    NE_SYNC();
}

/* eval_cbar: pi = inv(B)' cB, d_j = c_k - N_j' pi, lib/glpspx01.js:514-584 */
NE_D void ne_eval_cbar(const NeT &t, const NeProb &P, NeS &S)
{
    const int m = P.m, n = P.n, ldb = P.ldb;
    double *pi = S.w;
    for (int r = t.tid; r < m; r += t.nt) {
        double s = 0.0;
        for (int pos = 0; pos < m; pos++) {
            int k = S.head[pos];
            if (k >= m) s += S.Bi[pos * ldb + r] * P.cw[k - m];
        }
        pi[r] = s;
    }
    NE_SYNC();
    for (int j = t.tid; j < n; j += t.nt) {
        int k = S.head[m + j];
        double d;
        if (k < m) d = -pi[k];
        else {
            d = P.cw[k - m] + ne_coldot(S.A, P.lda, m, k - m, pi);
        }
        S.cbar[j] = d;
    }
    NE_SYNC();
}

/* eval_obj: lib/glpspx02.js:1424-1450 */
NE_D double ne_eval_obj(const NeT &t, const NeProb &P, NeS &S)
{
    const int m = P.m, n = P.n;
    double s = 0.0;
    for (int i = t.tid; i < m; i += t.nt) {
        int k = S.head[i];
        if (k >= m) s += P.obj[k - m] * S.bbar[i];
    }
    for (int j = t.tid; j < n; j += t.nt) {
        int k = S.head[m + j];
        if (k >= m) s += P.obj[k - m] * ne_xN(S, m, j);
    }
    s = ne_block_reduce(t, s, NeSum(), (double *)S.red);
    return P.c0 + s;
}

/* check_feas: lib/glpspx02.js:1296-1315 (types of the node, not the working ones) */
NE_D int ne_check_feas(const NeT &t, const NeProb &P, NeS &S, double tol_dj)
{
    int bad = 0;
    for (int j = t.tid; j < P.n; j += t.nt) {
        int ty = S.type[S.head[P.m + j]];
        if (S.cbar[j] < -tol_dj && (ty == NE_LO || ty == NE_FR)) bad = 1;
        if (S.cbar[j] > +tol_dj && (ty == NE_UP || ty == NE_FR)) bad = 1;
    }
    return ne_block_or(t, bad);
}

/* check_stab: lib/glpspx02.js:1410-1422 */
NE_D int ne_check_stab(const NeT &t, const NeProb &P, NeS &S, double tol_dj)
{
    int bad = 0;
    for (int j = t.tid; j < P.n; j += t.nt) {
        int st = S.nstat[j];
        if (S.cbar[j] < -tol_dj && (st == NE_NL || st == NE_NF)) bad = 1;
        if (S.cbar[j] > +tol_dj && (st == NE_NU || st == NE_NF)) bad = 1;
    }
    return ne_block_or(t, bad);
}

/* set_aux_bnds: lib/glpspx02.js:1317-1359 */
NE_D void ne_set_aux_bnds(const NeT &t, const NeProb &P, NeS &S)
{
    const int m = P.m, n = P.n, mn = m + n;
    for (int k = t.tid; k < mn; k += t.nt) {
        switch (S.type[k]) {
        case NE_FR: S.wtype[k] = NE_DB; S.sl[k] = -1e3; S.su[k] = +1e3; break;
        case NE_LO: S.wtype[k] = NE_DB; S.sl[k] = 0.0; S.su[k] = +1.0; break;
        case NE_UP: S.wtype[k] = NE_DB; S.sl[k] = -1.0; S.su[k] = 0.0; break;
        default: S.wtype[k] = NE_FX; S.sl[k] = S.su[k] = 0.0; break;
        }
    }
    NE_SYNC();
    for (int j = t.tid; j < n; j += t.nt) {
        int k = S.head[m + j];
        if (S.wtype[k] == NE_FX) S.nstat[j] = NE_NS;
        else if (S.cbar[j] >= 0.0) S.nstat[j] = NE_NL;
        else S.nstat[j] = NE_NU;
    }
    NE_SYNC();
}

/* set_orig_bnds: lib/glpspx02.js:1361-1408 */
NE_D void ne_set_orig_bnds(const NeT &t, const NeProb &P, NeS &S)
{
    const int m = P.m, n = P.n;
    ne_orig_bounds(t, P, S);
    NE_SYNC();
    for (int j = t.tid; j < n; j += t.nt) {
        int k = S.head[m + j];
        switch (S.wtype[k]) {
        case NE_FR: S.nstat[j] = NE_NF; break;
        case NE_LO: S.nstat[j] = NE_NL; break;
        case NE_UP: S.nstat[j] = NE_NU; break;
        case NE_DB:
            if (S.cbar[j] >= +DBL_EPSILON) S.nstat[j] = NE_NL;
            else if (S.cbar[j] <= -DBL_EPSILON) S.nstat[j] = NE_NU;
            else if (fabs(S.sl[k]) <= fabs(S.su[k])) S.nstat[j] = NE_NL;
            else S.nstat[j] = NE_NU;
            break;
        default: S.nstat[j] = NE_NS; break;
        }
    }
    NE_SYNC();
}

/* store_sol: lib/glpspx02.js:1499-1590 */
NE_D void ne_store_sol(const NeT &t, const NeProb &P, NeS &S)
{
    const int m = P.m, n = P.n;
    for (int i = t.tid; i < m; i += t.nt) {
        int k = S.head[i];
        S.kstat[k] = NE_BS;
        S.prim[k] = (k < m) ? S.bbar[i] / P.rii[k] : S.bbar[i] * P.sjj[k - m];
        S.dual[k] = 0.0;
    }
    for (int j = t.tid; j < n; j += t.nt) {
        int k = S.head[m + j];
        int st = S.nstat[j];
        S.kstat[k] = (signed char)st;
        S.prim[k] = (st == NE_NU) ? S.ub[k] : (st == NE_NF ? 0.0 : S.lb[k]);
        S.dual[k] = (k < m) ? (S.cbar[j] * P.rii[k]) / P.zeta : (S.cbar[j] / P.sjj[k - m]) / P.zeta;
    }
    NE_SYNC();
}

struct NeChuzrOp {     /* larger a wins, lowest idx on ties */
    NE_M NeKey operator()(const NeKey &x, const NeKey &y) const
    {
        if (y.a > x.a || (y.a == x.a && y.idx < x.idx)) return y;
        return x;
    }
};
struct NeRatio1Op {    /* smaller a, then larger b, then lowest idx */
    NE_M NeKey operator()(const NeKey &x, const NeKey &y) const
    {
        if (y.idx == INT_MAX) return x;
        if (x.idx == INT_MAX) return y;
        if (y.a < x.a || (y.a == x.a && (y.b > x.b || (y.b == x.b && y.idx < x.idx)))) return y;
        return x;
    }
};
struct NeRatio2Op {    /* larger b, lowest idx on ties */
    NE_M NeKey operator()(const NeKey &x, const NeKey &y) const
    {
        if (y.idx == INT_MAX) return x;
        if (x.idx == INT_MAX) return y;
        if (y.b > x.b || (y.b == x.b && y.idx < x.idx)) return y;
        return x;
    }
};

struct NeLpOut { int ret, pbs, dbs, iters, refacs; double obj; };

/* ios_solve_node -> glp_simplex(meth = GLP_DUALP, obj_ll/obj_ul = incumbent):
   lib/glpios01.js:866-910, lib/glpapi06.js:3-39, lib/glpspx02.js:1592-1966.
   The primal fall-back of solve_lp is not built here: a failing dual solve
   ends the task with NE_R_FAIL. */
NE_D void ne_solve_lp(const NeT &t, const NeProb &P, NeS &S, double obj_ll, double obj_ul, NeLpOut &out,
                      int &bi_upd, long long *cyc)
{
    const int m = P.m, n = P.n, ldb = P.ldb;
    /* bi_upd >= 0: S.Bi already holds the inverse of this basis (the parent's final one, or
       the one the previous solve of this node ended with) after bi_upd updates: the solve
       starts like the reference's with a valid factorisation (binv_st = 2) */
    int binv_st = (bi_upd >= 0) ? 2 : 0, bbar_st = 0, cbar_st = 0, rigorous = 0, phase = 0, refct = 0;
    int upd = (bi_upd >= 0) ? bi_upd : 0;
    int it = 0, refacs = 0;
    double objt = 0.0;
    long long c0 = ne_clock();
    out.ret = 0; out.pbs = out.dbs = NE_UNDEF; out.obj = 0.0;
    bi_upd = -1;
    if (ne_build_head(t, P, S, binv_st == 2)) { out.ret = NE_EFAIL; out.iters = 0; out.refacs = 0; return; }
    ne_orig_bounds(t, P, S);
    for (int i = t.tid; i < m; i += t.nt) S.gamma[i] = 1.0;
    NE_SYNC();
    for (;;) {
        if (binv_st == 0) {
            long long c1 = ne_clock();
            if (ne_invert(t, P, S)) { NE_TRACE("  lp: singular basis\n"); out.ret = NE_EFAIL; break; }
            refacs++;
            binv_st = 1; bbar_st = cbar_st = 0; upd = 0;
            if (cyc) { long long c2 = ne_clock(); cyc[1] += c2 - c1; c0 += c2 - c1; }
        }
        if (cbar_st == 0) {
            ne_eval_cbar(t, P, S);
            cbar_st = 1;
            if (phase == 0) {
                if (ne_check_feas(t, P, S, 0.90 * P.tol_dj)) { phase = 1; ne_set_aux_bnds(t, P, S); }
                else { phase = 2; ne_set_orig_bnds(t, P, S); }
                refct = 0;
                bbar_st = 0;
            }
            if (ne_check_stab(t, P, S, P.tol_dj)) { NE_TRACE("  lp: check_stab failed (phase %d it %d)\n", phase, it); out.ret = NE_EFAIL; break; }   /* meth == GLP_DUALP */
        }
        if (phase == 1 && !ne_check_feas(t, P, S, P.tol_dj)) {
            phase = 2;
            if (cbar_st != 1) { ne_eval_cbar(t, P, S); cbar_st = 1; }
            ne_set_orig_bnds(t, P, S);
            refct = 0;
            bbar_st = 0;
        }
        if (bbar_st == 0) {
            ne_eval_bbar(t, P, S);
            if (phase == 2) objt = ne_eval_obj(t, P, S);
            bbar_st = 1;
        }
        if (refct == 0) {   /* reset_refsp: lib/glpspx02.js:497-512 */
            refct = 1000;
            for (int k = t.tid; k < m + n; k += t.nt) S.refsp[k] = 0;
            NE_SYNC();
            for (int i = t.tid; i < m; i += t.nt) { S.refsp[S.head[i]] = 1; S.gamma[i] = 1.0; }
            NE_SYNC();
        }
        /* objective cut-off: lib/glpspx02.js:1729-1760 */
        if (phase == 2 && ((P.zeta < 0.0 && obj_ll > -DBL_MAX && objt <= obj_ll) ||
                           (P.zeta > 0.0 && obj_ul < +DBL_MAX && objt >= obj_ul))) {
            if (bbar_st != 1 || cbar_st != 1) {
                if (bbar_st != 1) bbar_st = 0;
                if (cbar_st != 1) cbar_st = 0;
                continue;
            }
            out.pbs = NE_INFEAS; out.dbs = NE_FEAS;
            out.ret = (P.zeta < 0.0) ? NE_EOBJLL : NE_EOBJUL;
            break;
        }
        if (it >= P.it_max) { out.ret = NE_EITLIM; break; }
        if (cyc && it == 0 && bbar_st == 1 && cbar_st == 1) { long long c2 = ne_clock(); cyc[2] += c2 - c0; c0 = c2; }
        NE_TRACE("  lp: it %d phase %d obj %.10g st %d%d%d\n", it, phase, objt, binv_st, bbar_st, cbar_st);
        /* chuzr: lib/glpspx02.js:572-625 */
        NeKey kp = {0.0, 0.0, INT_MAX, 0};
        for (int i = t.tid; i < m; i += t.nt) {
            int k = S.head[i];
            int ty = S.wtype[k];
            double ri = 0.0, b = S.bbar[i];
            if (ty == NE_LO || ty == NE_DB || ty == NE_FX) {
                double eps = P.tol_bnd * (1.0 + 0.10 * fabs(S.sl[k]));
                if (b < S.sl[k] - eps) ri = S.sl[k] - b;
            }
            if (ty == NE_UP || ty == NE_DB || ty == NE_FX) {
                double eps = P.tol_bnd * (1.0 + 0.10 * fabs(S.su[k]));
                if (b > S.su[k] + eps) ri = S.su[k] - b;
            }
            if (ri == 0.0) continue;
            double g = S.gamma[i];
            if (g < DBL_EPSILON) g = DBL_EPSILON;
            double temp = (ri * ri) / g;
            if (temp > kp.a) { kp.a = temp; kp.b = ri; kp.idx = i; }
        }
        kp = ne_block_reduce(t, kp, NeChuzrOp(), S.red);
        if (kp.idx == INT_MAX || !(kp.a > 0.0)) {
            if (bbar_st != 1 || cbar_st != 1) {
                if (bbar_st != 1) bbar_st = 0;
                if (cbar_st != 1) cbar_st = 0;
                continue;
            }
            if (phase == 1) {
                ne_set_orig_bnds(t, P, S);
                ne_eval_bbar(t, P, S);
                out.pbs = NE_INFEAS; out.dbs = NE_NOFEAS;
            } else out.pbs = out.dbs = NE_FEAS;
            out.ret = 0;
            break;
        }
        const int p = kp.idx;
        const double delta = kp.b;
        /* eval_rho = row p of the inverse; eval_trow1 (column dots), lib/glpspx02.js:627-693 */
        for (int r = t.tid; r < m; r += t.nt) S.rho[r] = S.Bi[p * ldb + r];
        NE_SYNC();
        double big = 0.0;
        for (int j = t.tid; j < n; j += t.nt) {
            double v = 0.0;
            if (S.nstat[j] != NE_NS) {
                int k = S.head[m + j];
                if (k < m) v = -S.rho[k];
                else v = ne_coldot(S.A, P.lda, m, k - m, S.rho);
            }
            S.trow[j] = v;
            if (fabs(v) > big) big = fabs(v);
        }
        big = ne_block_reduce(t, big, NeMax(), (double *)S.red);
        /* sort_trow(tol_bnd): lib/glpspx02.js:754-791, 1851 */
        const double eps_t = P.tol_bnd * (1.0 + 0.01 * big);
        /* chuzc (Harris): lib/glpspx02.js:793-935 */
        const double s = (delta > 0.0 ? +1.0 : -1.0);
        const double rtol = 0.30 * P.tol_dj;
        NeKey k1 = {DBL_MAX, 0.0, INT_MAX, 0};
        for (int j = t.tid; j < n; j += t.nt) {
            double tv = S.trow[j];
            if (!(fabs(tv) >= eps_t) || tv == 0.0) continue;
            double alfa = s * tv, tt;
            int st = S.nstat[j];
            if (alfa > 0.0) { if (st == NE_NL || st == NE_NF) tt = (S.cbar[j] + rtol) / alfa; else continue; }
            else { if (st == NE_NU || st == NE_NF) tt = (S.cbar[j] - rtol) / alfa; else continue; }
            if (tt < 0.0) tt = 0.0;
            if (k1.idx == INT_MAX || tt < k1.a || (tt == k1.a && fabs(alfa) > k1.b)) { k1.a = tt; k1.b = fabs(alfa); k1.idx = j; }
        }
        k1 = ne_block_reduce(t, k1, NeRatio1Op(), S.red);
        int q = (k1.idx == INT_MAX) ? -1 : k1.idx;
        double teta = k1.a;
        if (q >= 0 && teta != 0.0) {
            const double tmax = teta;
            NeKey k2 = {0.0, 0.0, INT_MAX, 0};
            for (int j = t.tid; j < n; j += t.nt) {
                double tv = S.trow[j];
                if (!(fabs(tv) >= eps_t) || tv == 0.0) continue;
                double alfa = s * tv, tt;
                int st = S.nstat[j];
                if (alfa > 0.0) { if (st == NE_NL || st == NE_NF) tt = S.cbar[j] / alfa; else continue; }
                else { if (st == NE_NU || st == NE_NF) tt = S.cbar[j] / alfa; else continue; }
                if (tt < 0.0) tt = 0.0;
                if (tt <= tmax && (k2.idx == INT_MAX || fabs(alfa) > k2.b)) { k2.a = tt; k2.b = fabs(alfa); k2.idx = j; }
            }
            k2 = ne_block_reduce(t, k2, NeRatio2Op(), S.red);
            q = k2.idx; teta = k2.a;      /* pass 2 always finds the pass-1 winner at least */
        }
        if (q < 0) {
            if (bbar_st != 1 || cbar_st != 1 || !rigorous) {
                if (bbar_st != 1) bbar_st = 0;
                if (cbar_st != 1) cbar_st = 0;
                if (binv_st != 1) binv_st = 0;      /* rigorous: here a fresh inverse replaces refine_rho */
                rigorous = 1;
                continue;
            }
            if (phase == 1) { NE_TRACE("  lp: no q in phase 1\n"); out.ret = NE_EFAIL; break; }
            out.pbs = NE_NOFEAS; out.dbs = NE_FEAS; out.ret = 0;
            break;
        }
        const double new_dq = s * teta;
        NE_TRACE("    p %d (k %d) delta %.6g bbar %.6g [%.6g,%.6g] q %d (k %d) trow %.6g cbar %.6g new_dq %.6g\n", p, S.head[p], delta, S.bbar[p], S.sl[S.head[p]], S.su[S.head[p]], q, S.head[m + q], S.trow[q], S.cbar[q], new_dq);
        {
            double piv = S.trow[q];
            double eps = 1e-5 * (1.0 + 0.01 * big);
            if (fabs(piv) < eps && !rigorous) {
                rigorous = 5;
                if (binv_st != 1) binv_st = 0;
                continue;
            }
        }
        /* eval_tcol: lib/glpspx02.js:937-977 */
        {
            int k = S.head[m + q];
            for (int pos = t.tid; pos < m; pos += t.nt) {
                double v = 0.0;
                if (k < m) v = -S.Bi[pos * ldb + k];
                else for (int r = 0; r < m; r++) v += S.Bi[pos * ldb + r] * S.A[(size_t)r * P.lda + (k - m)];
                S.tcol[pos] = v;
            }
            NE_SYNC();
        }
        {
            double piv1 = S.tcol[p], piv2 = S.trow[q];
            if (piv1 == 0.0 || fabs(piv1 - piv2) > 1e-8 * (1.0 + fabs(piv1)) ||
                !((piv1 > 0.0 && piv2 > 0.0) || (piv1 < 0.0 && piv2 < 0.0))) {
                if (binv_st != 1 || !rigorous) {
                    if (binv_st != 1) binv_st = 0;
                    rigorous = 5;
                    NE_SYNC();
                    continue;
                }
                NE_SYNC();
                if (t.tid == 0) S.tcol[p] = piv2;
                NE_SYNC();
            }
        }
        const double pivot = S.tcol[p];
        /* update_bbar: lib/glpspx02.js:1042-1073; objective: :1936-1938 */
        {
            const double tb = delta / pivot;
            const double xq = ne_xN(S, m, q);
            const double dq = S.cbar[q];
            NE_SYNC();
            for (int i = t.tid; i < m; i += t.nt) {
                if (i == p) S.bbar[i] = xq + tb;
                else if (tb != 0.0) S.bbar[i] += S.tcol[i] * tb;
            }
            if (phase == 2) objt += (dq / P.zeta) * tb;
            bbar_st = 2;
        }
        /* update_cbar: lib/glpspx02.js:1020-1040 */
        for (int j = t.tid; j < n; j += t.nt) {
            if (j == q) S.cbar[j] = new_dq;
            else if (new_dq != 0.0) S.cbar[j] -= S.trow[j] * new_dq;
        }
        cbar_st = 2;
        /* update_gamma: lib/glpspx02.js:1075-1188 */
        {
            const int kq = S.head[m + q], kp_ = S.head[p];
            refct--;
            double gsum = 0.0;
            for (int j = t.tid; j < n; j += t.nt) {
                double tv = S.trow[j];
                if (tv != 0.0 && S.refsp[S.head[m + j]]) gsum += tv * tv;
            }
            gsum = ne_block_reduce(t, gsum, NeSum(), (double *)S.red);
            const double eta_p = S.refsp[kp_] ? 1.0 : 0.0;
            const double gamma_p = eta_p + gsum;
            for (int i = t.warp; i < m; i += t.nwarps) {
                double sacc = 0.0;
                for (int j = t.lane; j < n; j += t.wsize) {
                    double tv = S.trow[j];
                    if (tv == 0.0) continue;
                    int k = S.head[m + j];
                    if (!S.refsp[k]) continue;
                    if (k < m) { if (k == i) sacc += tv; }
                    else sacc -= tv * S.A[(size_t)i * P.lda + (k - m)];
                }
                sacc = ne_warp_reduce(t, sacc, NeSum());
                if (t.lane == 0) S.w[i] = sacc;
            }
            NE_SYNC();
            ne_ftran(t, P, S, S.w, S.u);
            NE_SYNC();
            const int q_free = (S.wtype[kq] == NE_FR);
            const int p_fixed_ref = (S.wtype[kp_] == NE_FX && S.refsp[kp_]);
            for (int i = t.tid; i < m; i += t.nt) {
                double g;
                if (i == p) {
                    if (q_free) g = 1.0;
                    else { g = gamma_p / (pivot * pivot); if (g < DBL_EPSILON) g = DBL_EPSILON; }
                    if (p_fixed_ref && !q_free) { double tt = 1.0 / pivot; g -= tt * tt; if (g < DBL_EPSILON) g = DBL_EPSILON; }
                } else {
                    g = S.gamma[i];
                    double tc = S.tcol[i];
                    if (tc != 0.0 && S.wtype[S.head[i]] != NE_FR) {
                        double tt = tc / pivot;
                        double t1 = g + tt * tt * gamma_p + 2.0 * tt * S.u[i];
                        double t2 = (S.refsp[S.head[i]] ? 1.0 : 0.0) + eta_p * tt * tt;
                        g = (t1 >= t2 ? t1 : t2);
                        if (g < DBL_EPSILON) g = DBL_EPSILON;
                        if (p_fixed_ref) { g -= tt * tt; if (g < DBL_EPSILON) g = DBL_EPSILON; }
                    }
                }
                S.gamma[i] = g;
            }
            NE_SYNC();
            if (p_fixed_ref && t.tid == 0) S.refsp[kp_] = 0;
        }
        /* basis change on the explicit inverse (replaces fhv_update_it): row p of
           the old inverse is rho, the new column is w = inv(B) N_q = -tcol, so
           Bi'[p] = rho / w_p = -rho / pivot and Bi'[i] = Bi[i] - (w_i / w_p) rho */
        for (int e = t.tid; e < m * m; e += t.nt) {
            int i = e / m, r = e - i * m;
            double rp = S.rho[r] / pivot;
            if (i == p) S.Bi[i * ldb + r] = -rp;
            else { double tc = S.tcol[i]; if (tc != 0.0) S.Bi[i * ldb + r] -= tc * rp; }
        }
        /* change_basis: lib/glpspx02.js:1259-1294 */
        if (t.tid == 0) {
            int k = S.head[p];
            S.head[p] = S.head[m + q];
            S.head[m + q] = k;
            S.bind[S.head[p]] = p;
            S.bind[S.head[m + q]] = m + q;
            if (S.wtype[k] == NE_FX) S.nstat[q] = NE_NS;
            else if (delta > 0.0) S.nstat[q] = NE_NL;
            else S.nstat[q] = NE_NU;
        }
        NE_SYNC();
        it++;
        upd++;
        binv_st = 2;
        if (upd >= P.refac_period) binv_st = 0;
        if (rigorous > 0) rigorous--;
    }
    if (out.ret != NE_EFAIL && out.ret != NE_EITLIM) {
        out.obj = ne_eval_obj(t, P, S);
        ne_store_sol(t, P, S);
        if (binv_st != 0) bi_upd = upd;       /* S.Bi is the inverse of the final basis */
    }
    out.iters = it; out.refacs = refacs;
    if (cyc) cyc[3] += ne_clock() - c0;
}

/* ------------------------------------------------------------------ */
/* after the LP: bounds, integrality, branching                       */
/* ------------------------------------------------------------------ */
struct NeRound { int ok; long long d; double s; };
struct NeRoundAcc { int bad, big, any; long long g; double s; };
NE_D long long ne_gcd(long long a, long long b) { while (b > 0) { long long r = a % b; a = b; b = r; } return a; }
struct NeRoundOp {
    NE_M NeRoundAcc operator()(const NeRoundAcc &a, const NeRoundAcc &b) const
    {
        NeRoundAcc r;
        r.bad = a.bad | b.bad; r.big = a.big | b.big; r.any = a.any | b.any;
        r.g = ne_gcd(a.g, b.g); r.s = a.s + b.s;
        return r;
    }
};

/* the column scan of ios_round_bound (lib/glpios01.js:730-787) */
NE_D NeRound ne_round_prepare(const NeT &t, const NeProb &P, NeS &S)
{
    const int m = P.m;
    NeRoundAcc a = {0, 0, 0, 0, 0.0};
    for (int j = t.tid; j < P.n; j += t.nt) {
        double cf = P.ucoef[j];
        if (cf == 0.0) continue;
        if (S.type[m + j] == NE_FX) a.s += cf * S.prim[m + j];
        else {
            if (!P.kind[j] || cf != floor(cf)) { a.bad = 1; continue; }
            if (fabs(cf) <= (double)INT_MAX) { a.any = 1; a.g = ne_gcd(a.g, (long long)fabs(cf)); }
            else a.big = 1;
        }
    }
    a = ne_block_reduce(t, a, NeRoundOp(), (NeRoundAcc *)S.red);
    NeRound r;
    r.s = P.c0 + a.s;
    if (a.bad) { r.ok = 0; r.d = 1; return r; }
    if (a.big) { r.ok = 1; r.d = 1; return r; }
    if (!a.any) { r.ok = 0; r.d = 1; return r; }
    r.ok = 1; r.d = a.g;
    return r;
}

NE_D double ne_round_bound(const NeProb &P, const NeRound &r, double bound)
{
    if (!r.ok) return bound;
    const double d = (double)r.d;
    if (P.dir == NE_MIN) {
        if (bound != +DBL_MAX) {
            double h = (bound - r.s) / d;
            if (h >= floor(h) + 0.001) bound = d * ceil(h) + r.s;
        }
    } else {
        if (bound != -DBL_MAX) {
            double h = (bound - r.s) / d;
            if (h <= ceil(h) - 0.001) bound = d * floor(h) + r.s;
        }
    }
    return bound;
}

/* ios_is_hopeful: lib/glpios01.js:789-819 */
NE_D int ne_is_hopeful(const NeProb &P, int have_inc, double mip_obj, double bound)
{
    if (have_inc) {
        double eps = P.tol_obj * (1.0 + fabs(mip_obj));
        if (P.dir == NE_MIN) { if (bound >= mip_obj - eps) return 0; }
        else { if (bound <= mip_obj + eps) return 0; }
    } else {
        if (P.dir == NE_MIN) { if (bound == +DBL_MAX) return 0; }
        else { if (bound == -DBL_MAX) return 0; }
    }
    return 1;
}

struct NeIntAcc { int cnt, pad; double sum; };
struct NeIntOp {
    NE_M NeIntAcc operator()(const NeIntAcc &a, const NeIntAcc &b) const
    {
        NeIntAcc r; r.cnt = a.cnt + b.cnt; r.pad = 0; r.sum = a.sum + b.sum; return r;
    }
};

struct NeTabKey { double teta, big, alfa; int kk, pad; };
struct NeTabOp {
    NE_M NeTabKey operator()(const NeTabKey &x, const NeTabKey &y) const
    {
        if (y.kk == INT_MAX) return x;
        if (x.kk == INT_MAX) return y;
        if (y.teta < x.teta || (y.teta == x.teta && (y.big > x.big || (y.big == x.big && y.kk < x.kk)))) return y;
        return x;
    }
};

/* reduced cost of kk with the sign forced (lib/glpios09.js:190-212) */
NE_D double ne_fixed_sign_dual(const NeProb &P, const NeS &S, int kk)
{
    int st = S.kstat[kk];
    double g = S.dual[kk];
    if (P.dir == NE_MIN) { if ((st == NE_NL && g < 0.0) || (st == NE_NU && g > 0.0) || st == NE_NF) g = 0.0; }
    else { if ((st == NE_NL && g > 0.0) || (st == NE_NU && g < 0.0) || st == NE_NF) g = 0.0; }
    return g;
}

/* Driebeck-Tomlin estimates of every fractional column: one warp per column
   computes its simplex-table row (glp_eval_tab_row, lib/glpapi12.js:401-453)
   on the fly and runs both dual ratio tests (glp_dual_rtest,
   lib/glpapi12.js:687-762, eps 1e-9) over it.  S.dz[0..3][f] = dz_dn, dz_up of
   branch_drtom (lib/glpios09.js:84-270) and the down / up objective estimates
   of ios_eval_degrad (lib/glpios01.js:615-728). */
NE_D void ne_tab_estimates(const NeT &t, const NeProb &P, NeS &S, int nfrac, double obj_val)
{
    const int m = P.m, n = P.n, ldb = P.ldb;
    const double osign = (P.dir == NE_MIN ? +1.0 : -1.0);
    const double inf = (P.dir == NE_MIN ? +DBL_MAX : -DBL_MAX);
    for (int f = t.warp; f < nfrac; f += t.nwarps) {
        const int j0 = S.frac[f];
        const int kb = m + j0;
        const int pos = S.bind[kb];
        const double sb = ne_scale(P, kb);
        const double *rho = S.Bi + pos * ldb;
        NeTabKey dn = {DBL_MAX, 0.0, 0.0, INT_MAX, 0}, up = {DBL_MAX, 0.0, 0.0, INT_MAX, 0};
        for (int jj = t.lane; jj < n; jj += t.wsize) {
            const int kk = S.head[m + jj];
            const int st = S.kstat[kk];
            if (!(st == NE_NL || st == NE_NU || st == NE_NF)) continue;
            double v = 0.0;
            if (kk < m) v = -rho[kk];
            else v = ne_coldot(S.A, P.lda, m, kk - m, rho);
            if (!P.unit_scale) v = v * sb / ne_scale(P, kk);
            if (v == 0.0) continue;
            const double cost = S.dual[kk];
            for (int kase = 0; kase < 2; kase++) {
                const double alfa = kase ? +v : -v;
                double temp;
                if (st == NE_NL) { if (alfa < +1e-9) continue; temp = (osign * cost) / alfa; }
                else if (st == NE_NU) { if (alfa > -1e-9) continue; temp = (osign * cost) / alfa; }
                else { if (-1e-9 < alfa && alfa < +1e-9) continue; temp = 0.0; }
                if (temp < 0.0) temp = 0.0;
                NeTabKey &b = kase ? up : dn;
                if (b.kk == INT_MAX || temp < b.teta || (temp == b.teta && (fabs(alfa) > b.big || (fabs(alfa) == b.big && kk < b.kk)))) {
                    b.teta = temp; b.big = fabs(alfa); b.alfa = v; b.kk = kk;
                }
            }
        }
        dn = ne_warp_reduce(t, dn, NeTabOp());
        up = ne_warp_reduce(t, up, NeTabOp());
        if (t.lane == 0) {
            const double x = S.prim[kb];
            for (int kase = 0; kase < 2; kase++) {
                const NeTabKey &b = kase ? up : dn;
                double dz, est;
                if (b.kk == INT_MAX) { dz = inf; est = inf; }
                else {
                    double delta_k = ((kase ? ceil(x) : floor(x)) - x) / b.alfa;
                    double dual_k = ne_fixed_sign_dual(P, S, b.kk);
                    est = obj_val + dual_k * delta_k;
                    if (b.kk >= m && P.kind[b.kk - m])
                        if (fabs(delta_k - floor(delta_k + 0.5)) > 1e-3)
                            delta_k = (delta_k > 0.0) ? ceil(delta_k) : floor(delta_k);
                    dz = dual_k * delta_k;
                }
                S.dz[kase * m + f] = dz;
                S.dz[(2 + kase) * m + f] = est;
            }
        }
    }
    NE_SYNC();
}

/* ordered list of the set flags (ascending index); returns the count */
NE_D int ne_compact(const NeT &t, const signed char *flag, int n, int *out, int cap, int *si)
{
#ifdef NE_EMUL
    (void)t; (void)si;
    int c = 0;
    for (int j = 0; j < n; j++) if (flag[j]) { if (c < cap) out[c] = j; c++; }
    return c;
#else
    if (t.warp == 0) {
        int c = 0;
        for (int base = 0; base < n; base += 32) {
            int j = base + t.lane;
            int f = (j < n) ? (flag[j] != 0) : 0;
            unsigned mask = __ballot_sync(0xffffffffu, f);
            if (f) { int o = c + __popc(mask & ((1u << t.lane) - 1u)); if (o < cap) out[o] = j; }
            c += __popc(mask);
        }
        if (t.lane == 0) si[0] = c;
    }
    __syncthreads();
    int c = si[0];
    __syncthreads();
    return c;
#endif
}

/* improve the incumbent shared by the CTAs of this launch */
NE_D void ne_publish_incumbent(const NeProb &P, double obj)
{
#ifdef NE_EMUL
    if (!*P.inc_have || (P.dir == NE_MIN ? obj < *P.inc : obj > *P.inc)) { *P.inc = obj; *P.inc_have = 1; }
#else
    unsigned long long *addr = (unsigned long long *)P.inc;
    if (atomicCAS(P.inc_have, 0, 2) == 0) {          /* first incumbent: publish value, then flag */
        atomicExch(addr, (unsigned long long)__double_as_longlong(obj));
        __threadfence();
        atomicExch(P.inc_have, 1);
        return;
    }
    while (atomicAdd(P.inc_have, 0) == 2) { }        /* another CTA is publishing the first one */
    unsigned long long old = atomicAdd(addr, 0ull);
    for (;;) {
        double cur = __longlong_as_double((long long)old);
        if (!(P.dir == NE_MIN ? obj < cur : obj > cur)) break;
        unsigned long long prev = atomicCAS(addr, old, (unsigned long long)__double_as_longlong(obj));
        if (prev == old) break;
        old = prev;
    }
#endif
}

NE_D void ne_read_incumbent(const NeT &t, const NeProb &P, NeS &S, int &have, double &obj)
{
    if (t.tid == 0) {
#ifdef NE_EMUL
        S.si[1] = *P.inc_have; S.sc[1] = *P.inc;
#else
        int h = atomicAdd(P.inc_have, 0);
        S.si[1] = (h == 1);
        S.sc[1] = __longlong_as_double((long long)atomicAdd((unsigned long long *)P.inc, 0ull));
#endif
    }
    NE_SYNC();
    have = S.si[1]; obj = S.sc[1];
    NE_SYNC();
}

NE_D void ne_store_state(const NeT &t, const NeProb &P, const NeS &S, int slot)
{
    const int mn = P.m + P.n;
    double *lb = P.slab_lb + (size_t)slot * mn, *ub = P.slab_ub + (size_t)slot * mn;
    signed char *ty = P.slab_type + (size_t)slot * mn, *st = P.slab_stat + (size_t)slot * mn;
    for (int k = t.tid; k < mn; k += t.nt) { lb[k] = S.lb[k]; ub[k] = S.ub[k]; ty[k] = S.type[k]; st[k] = S.kstat[k]; }
}

/* the whole `more:` state of ios_driver for one node (lib/glpios03.js:615-905) */
NE_D void ne_process_node(const NeT &t, const NeProb &P, NeS &S, const NeTask &task, NeResult &res, double *xout)
{
    const int m = P.m, n = P.n, mn = m + n;
    const long long ck0 = ne_clock();
    {
        const double *lb = P.slab_lb + (size_t)task.node * mn, *ub = P.slab_ub + (size_t)task.node * mn;
        const signed char *ty = P.slab_type + (size_t)task.node * mn, *st = P.slab_stat + (size_t)task.node * mn;
        for (int k = t.tid; k < mn; k += t.nt) { S.lb[k] = lb[k]; S.ub[k] = ub[k]; S.type[k] = ty[k]; S.kstat[k] = st[k]; }
    }
    int bi_upd = P.slab_upd[task.node];
    if (bi_upd >= 0) {
        const double *bi = P.slab_bi + (size_t)task.node * m * P.ldb;
        for (int e = t.tid; e < m * P.ldb; e += t.nt) S.Bi[e] = bi[e];
        const int *hd = P.slab_head + (size_t)task.node * m;
        for (int i = t.tid; i < m; i += t.nt) S.head[i] = hd[i];
    }
    NE_SYNC();
    long long cyc[6] = {0, 0, 0, 0, 0, 0};
    long long ck = ne_clock();
    cyc[5] += ck - ck0;
    double bound = task.bound, lp_obj = task.lp_obj;
    int code = 0, jv = -1, next = NE_NO_BRNCH, ii_cnt = 0, iters = 0, solves = 0, refacs = 0, lpret = 0;
    double ii_sum = 0.0, dn_lp = 0.0, dn_bnd = 0.0, up_lp = 0.0, up_bnd = 0.0, x_jv = 0.0, obj_val = 0.0;
    auto improve = [&](double &b, double v) { if (P.dir == NE_MIN) { if (b < v) b = v; } else { if (b > v) b = v; } };
    for (;;) {
        int have_inc; double mip_obj;
        ne_read_incumbent(t, P, S, have_inc, mip_obj);
        ck = ne_clock();
        if (P.pp_tech == NE_PP_ALL || (P.pp_tech == NE_PP_ROOT && task.level == 0)) {
            int inf_ = ne_preprocess(t, P, S, task.level == 0 ? 100 : 10, have_inc, mip_obj);
            cyc[0] += ne_clock() - ck;
            if (inf_) { code = NE_R_FATHOM; break; }
        }
        if (!ne_is_hopeful(P, have_inc, mip_obj, bound)) { code = NE_R_FATHOM; break; }
        /* ios_solve_node */
        NeLpOut lp;
        double obj_ll = -DBL_MAX, obj_ul = +DBL_MAX;
        if (have_inc) { if (P.dir == NE_MIN) obj_ul = mip_obj; else obj_ll = mip_obj; }
        ne_solve_lp(t, P, S, obj_ll, obj_ul, lp, bi_upd, cyc);
        ck = ne_clock();
        NE_TRACE("node %d lvl %d: lp ret %d pbs %d dbs %d obj %.10g it %d\n", task.node, task.level, lp.ret, lp.pbs, lp.dbs, lp.obj, lp.iters);
        solves++; iters += lp.iters; refacs += lp.refacs; lpret = lp.ret;
        if (!(lp.ret == 0 || lp.ret == NE_EOBJLL || lp.ret == NE_EOBJUL)) { code = NE_R_FAIL; break; }
        if (lp.pbs == NE_FEAS && lp.dbs == NE_FEAS) { }
        else if (lp.dbs == NE_NOFEAS) { code = NE_R_FAIL; break; }
        else if (lp.pbs == NE_INFEAS && lp.dbs == NE_FEAS) { code = NE_R_FATHOM; break; }
        else if (lp.pbs == NE_NOFEAS) { code = NE_R_FATHOM; break; }
        else { code = NE_R_FAIL; break; }
        obj_val = lp.obj;
        lp_obj = obj_val;
        NeRound rb = ne_round_prepare(t, P, S);
        improve(bound, ne_round_bound(P, rb, obj_val));
        if (!ne_is_hopeful(P, have_inc, mip_obj, bound)) { code = NE_R_FATHOM; break; }
        /* check_integrality: lib/glpios03.js:56-116 */
        auto integrality = [&]() {
            NeIntAcc a = {0, 0, 0.0};
            for (int j = t.tid; j < n; j += t.nt) {
                int k = m + j;
                S.flag[j] = 0;
                if (!P.kind[j] || S.kstat[k] != NE_BS) continue;
                int ty = S.type[k];
                double l = S.lb[k], u = S.ub[k], x = S.prim[k];
                if (ty == NE_LO || ty == NE_DB || ty == NE_FX) {
                    if (l - P.tol_int <= x && x <= l + P.tol_int) continue;
                    if (x < l) continue;
                }
                if (ty == NE_UP || ty == NE_DB || ty == NE_FX) {
                    if (u - P.tol_int <= x && x <= u + P.tol_int) continue;
                    if (x > u) continue;
                }
                double r = floor(x + 0.5);
                if (r - P.tol_int <= x && x <= r + P.tol_int) continue;
                S.flag[j] = 1;
                a.cnt++;
                double t1 = x - floor(x), t2 = ceil(x) - x;
                a.sum += (t1 <= t2 ? t1 : t2);
            }
            return ne_block_reduce(t, a, NeIntOp(), (NeIntAcc *)S.red);
        };
        NeIntAcc ia = integrality();
        if (ia.cnt == 0 && lp.refacs == 0) {
            /* an incumbent candidate whose values come from an inherited inverse: recompute them
               from a fresh one, as the reference's per-node glp_factorize would (lib/glpapi06.js:6-7) */
            NE_SYNC();
            if (!ne_invert(t, P, S)) {
                refacs++;
                ne_eval_cbar(t, P, S);
                ne_eval_bbar(t, P, S);
                obj_val = ne_eval_obj(t, P, S);
                ne_store_sol(t, P, S);
                lp_obj = obj_val;
                bi_upd = 0;
                ia = integrality();
            } else bi_upd = -1;
        }
        ii_cnt = ia.cnt; ii_sum = ia.sum;
        if (ii_cnt == 0) {
            /* record_solution: lib/glpios03.js:118-139 */
            for (int k = t.tid; k < mn; k += t.nt)
                xout[k] = (k >= m && P.kind[k - m]) ? floor(S.prim[k] + 0.5) : S.prim[k];
            if (t.tid == 0) ne_publish_incumbent(P, obj_val);
            code = NE_R_INTEGRAL;
            break;
        }
        /* fix_by_red_cost: lib/glpios03.js:307-377 */
        if (have_inc) {
            for (int j = t.tid; j < n; j += t.nt) {
                int k = m + j;
                if (!P.kind[j]) continue;
                double l = S.lb[k], u = S.ub[k], dj = S.dual[k];
                int st = S.kstat[k];
                if (P.dir == NE_MIN) {
                    if (st == NE_NL) { if (dj < 0.0) dj = 0.0; if (obj_val + dj >= mip_obj) ne_set_bnds(S, k, NE_FX, l, l); }
                    else if (st == NE_NU) { if (dj > 0.0) dj = 0.0; if (obj_val - dj >= mip_obj) ne_set_bnds(S, k, NE_FX, u, u); }
                } else {
                    if (st == NE_NL) { if (dj > 0.0) dj = 0.0; if (obj_val + dj <= mip_obj) ne_set_bnds(S, k, NE_FX, l, l); }
                    else if (st == NE_NU) { if (dj < 0.0) dj = 0.0; if (obj_val - dj <= mip_obj) ne_set_bnds(S, k, NE_FX, u, u); }
                }
            }
            NE_SYNC();
        }
        /* ios_choose_var: lib/glpios09.js:1-26 */
        const int nfrac = ne_compact(t, S.flag, n, S.frac, m, S.si);
        ne_tab_estimates(t, P, S, nfrac < m ? nfrac : m, obj_val);
        if (t.tid == 0) {
            int jj = -1, nx = NE_NO_BRNCH, fsel = -1;
            const int nf = nfrac < m ? nfrac : m;
            auto mostf = [&]() {
                double most = DBL_MAX;
                for (int f = 0; f < nf; f++) {
                    double beta = S.prim[m + S.frac[f]], temp = floor(beta) + 0.5;
                    if (most > fabs(beta - temp)) { jj = S.frac[f]; fsel = f; most = fabs(beta - temp); nx = (beta < temp) ? NE_DN_BRNCH : NE_UP_BRNCH; }
                }
            };
            if (P.br_tech == NE_BR_FFV || P.br_tech == NE_BR_LFV) {
                fsel = (P.br_tech == NE_BR_FFV) ? 0 : nf - 1;
                jj = S.frac[fsel];
                double beta = S.prim[m + jj];
                nx = (beta - floor(beta) < ceil(beta) - beta) ? NE_DN_BRNCH : NE_UP_BRNCH;
            } else if (P.br_tech == NE_BR_MFV) mostf();
            else {
                double degrad = -1.0;
                for (int f = 0; f < nf; f++) {
                    double dzd = S.dz[f], dzu = S.dz[m + f];
                    if (degrad < fabs(dzd) || degrad < fabs(dzu)) {
                        jj = S.frac[f]; fsel = f;
                        if (fabs(dzd) < fabs(dzu)) { nx = NE_DN_BRNCH; degrad = fabs(dzu); }
                        else { nx = NE_UP_BRNCH; degrad = fabs(dzd); }
                        if (degrad == DBL_MAX) break;
                    }
                }
                if (degrad < 1e-6 * (1.0 + 0.001 * fabs(obj_val))) mostf();
            }
            S.si[2] = jj; S.si[3] = nx; S.si[4] = fsel;
        }
        NE_SYNC();
        jv = S.si[2]; next = S.si[3];
        const int fsel = S.si[4];
        NE_SYNC();
        /* branch_on: lib/glpios03.js:141-305 */
        const int k = m + jv;
        const int type = S.type[k];
        const double l = S.lb[k], u = S.ub[k], beta = S.prim[k];
        const double new_ub = floor(beta), new_lb = ceil(beta);
        int dn_type, up_type;
        switch (type) {
        case NE_FR: dn_type = NE_UP; up_type = NE_LO; break;
        case NE_LO: dn_type = (l == new_ub ? NE_FX : NE_DB); up_type = NE_LO; break;
        case NE_UP: dn_type = NE_UP; up_type = (new_lb == u ? NE_FX : NE_DB); break;
        default: dn_type = (l == new_ub ? NE_FX : NE_DB); up_type = (new_lb == u ? NE_FX : NE_DB); break;
        }
        x_jv = beta;
        dn_lp = S.dz[2 * m + fsel]; up_lp = S.dz[3 * m + fsel];
        NeRound rb2 = have_inc ? ne_round_prepare(t, P, S) : rb;    /* fix_by_red_cost may have fixed columns */
        dn_bnd = ne_round_bound(P, rb2, dn_lp); up_bnd = ne_round_bound(P, rb2, up_lp);
        const int dn_bad = !ne_is_hopeful(P, have_inc, mip_obj, dn_bnd);
        const int up_bad = !ne_is_hopeful(P, have_inc, mip_obj, up_bnd);
        if (dn_bad && up_bad) { code = NE_R_FATHOM; break; }
        if (up_bad || dn_bad) {
            NE_SYNC();
            if (t.tid == 0) {
                if (up_bad) ne_set_bnds(S, k, dn_type, l, new_ub);
                else ne_set_bnds(S, k, up_type, new_lb, u);
            }
            if (up_bad) { lp_obj = dn_lp; improve(bound, dn_bnd); }
            else { lp_obj = up_lp; improve(bound, up_bnd); }
            NE_SYNC();
            continue;                       /* `more` again with the tightened bound */
        }
        /* two children: the frozen node state plus one bound each, and the inverse of
           the basis both start from */
        NE_SYNC();
        cyc[4] += ne_clock() - ck;
        ck = ne_clock();
        if (t.tid == 0) ne_set_bnds(S, k, dn_type, l, new_ub);
        NE_SYNC();
        ne_store_state(t, P, S, task.node);
        NE_SYNC();
        if (t.tid == 0) ne_set_bnds(S, k, up_type, new_lb, u);
        NE_SYNC();
        ne_store_state(t, P, S, task.child);
        if (bi_upd >= 0) {
            double *b1 = P.slab_bi + (size_t)task.node * m * P.ldb, *b2 = P.slab_bi + (size_t)task.child * m * P.ldb;
            for (int e = t.tid; e < m * P.ldb; e += t.nt) { double v = S.Bi[e]; b1[e] = v; b2[e] = v; }
            int *h1 = P.slab_head + (size_t)task.node * m, *h2 = P.slab_head + (size_t)task.child * m;
            for (int i = t.tid; i < m; i += t.nt) { h1[i] = S.head[i]; h2[i] = S.head[i]; }
        }
        if (t.tid == 0) { P.slab_upd[task.node] = bi_upd; P.slab_upd[task.child] = bi_upd; }
        cyc[5] += ne_clock() - ck;
        ck = ne_clock();
        code = NE_R_BRANCH;
        break;
    }
    if (code != NE_R_BRANCH) cyc[4] += ne_clock() - ck;
    if (t.tid == 0) {
        for (int c = 0; c < 6; c++) res.cyc[c] = cyc[c];
        res.code = code; res.jv = jv; res.next = next; res.ii_cnt = ii_cnt;
        res.iters = iters; res.solves = solves; res.refacs = refacs; res.ret = lpret;
        res.obj = (code == NE_R_INTEGRAL) ? obj_val : lp_obj;
        res.bound = bound; res.ii_sum = ii_sum;
        res.dn_lp = dn_lp; res.dn_bnd = dn_bnd; res.up_lp = up_lp; res.up_bnd = up_bnd; res.x_jv = x_jv;
    }
}

#endif /* GLPB_NODEENGINE_CUH */
