/* kernels.cuh -- hand-written sm_100a kernels of the simplex hot path.
 *
 * Every kernel that belongs to the iteration starts with
 *     if (ctrl->status != ST_OK) return;
 * so that the host can enqueue a whole iteration (or several) and look at the
 * control block once: the first kernel that meets one of the reference's
 * exceptional branches records it and the rest of the queue drains as no-ops.
 *
 * Selection kernels (pricing, ratio tests) use __dmul_rn/__dadd_rn/__ddiv_rn so
 * that no FMA contraction can make them differ from the reference's IEEE
 * double arithmetic: fed the same arrays they return the same index.
 * Reductions are two-stage and combine block partials in index order, so the
 * result does not depend on block scheduling.
 */
#ifndef GLPB_KERNELS_CUH
#define GLPB_KERNELS_CUH
#include "glpb_internal.cuh"

#define FULLMASK 0xffffffffu

/* ------------------------------------------------------------------ */
/* reduction machinery                                                */
/* ------------------------------------------------------------------ */

__device__ __forceinline__ Key key_shfl_down(const Key &v, int off)
{
    Key o;
    o.a = __shfl_down_sync(FULLMASK, v.a, off);
    o.b = __shfl_down_sync(FULLMASK, v.b, off);
    o.c = __shfl_down_sync(FULLMASK, v.c, off);
    o.pos = __shfl_down_sync(FULLMASK, v.pos, off);
    o.aux = __shfl_down_sync(FULLMASK, v.aux, off);
    return o;
}

/* reduction over the lanes of a warp, result in lane 0.  `none` is the identity of comb
   (comb(none, x) = x), which most lanes hold in the selection passes -- a ratio test has a
   handful of candidates among thousands of entries: a warp without a live value skips the
   shuffle tree (40 SHFL per Key), a warp with ONE live value fetches it with a single round. */
template <class Comb>
__device__ __forceinline__ Key warp_reduce(Key v, const Key &none, Comb comb)
{
    const bool live = !(v.pos == none.pos && v.aux == none.aux && v.a == none.a && v.b == none.b && v.c == none.c);
    const unsigned int mask = __ballot_sync(FULLMASK, live);
    if (mask == 0u) return v;
    if ((mask & (mask - 1u)) == 0u) {
        const int src = __ffs(mask) - 1;
        Key o;
        o.a = __shfl_sync(FULLMASK, v.a, src); o.b = __shfl_sync(FULLMASK, v.b, src); o.c = __shfl_sync(FULLMASK, v.c, src);
        o.pos = __shfl_sync(FULLMASK, v.pos, src); o.aux = __shfl_sync(FULLMASK, v.aux, src);
        return o;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Key o = key_shfl_down(v, off);
        comb(v, o);
    }
    return v;
}

/* result valid in thread 0 of the block; all threads must call */
template <class Comb>
__device__ Key block_reduce(Key v, const Key &none, Comb comb)
{
    __shared__ Key sm[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_reduce(v, none, comb);
    __syncthreads(); /* protect sm against a previous use */
    if (lane == 0) sm[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    if (w == 0) {
        v = (lane < nw) ? sm[lane] : none;
        v = warp_reduce(v, none, comb);
    }
    return v;
}

/* two-stage grid reduction; fin(result) runs in one thread of the last block */
template <class Comb, class Fin>
__device__ void grid_reduce(Key v, const Key &none, Key *scratch, unsigned int *ticket,
                            Comb comb, Fin fin)
{
    __shared__ int is_last;
    v = block_reduce(v, none, comb);
    if (threadIdx.x == 0) {
        scratch[blockIdx.x] = v;
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        Key r = none;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
            Key o;
            o.a = __ldcg(&scratch[i].a); o.b = __ldcg(&scratch[i].b); o.c = __ldcg(&scratch[i].c);
            o.pos = __ldcg(&scratch[i].pos); o.aux = __ldcg(&scratch[i].aux);
            comb(r, o);
        }
        /* NOTE: per-thread accumulation above visits partials in a fixed
           strided order, the tree below is fixed too: deterministic */
        r = block_reduce(r, none, comb);
        if (threadIdx.x == 0) {
            *ticket = 0u;
            fin(r);
        }
    }
}

struct CombArgMax { /* larger a wins, then smaller pos */
    __device__ void operator()(Key &v, const Key &o) const
    {
        if (o.a > v.a || (o.a == v.a && o.pos < v.pos)) v = o;
    }
};
/* Ratio tests: candidates that agree in every compared field are ordered by `pos`.
   With a materialised list (k_sort_list) pos is the place in the reference's own
   sort_tcol / sort_trow order and the winner is the reference's; in dense scans pos
   is the index, and the winner of an exact tie carries RATIO_TIE in aux so that the
   caller can have the iteration decided in list order instead.  The bound-flip
   candidate (pos -1) is examined first by the reference as well: no tie. */
#define RATIO_TIE 512
struct CombRatio1 { /* smaller a (ratio), then larger b (|alfa|), then smaller pos */
    __device__ void operator()(Key &v, const Key &o) const
    {
        if (o.a < v.a || (o.a == v.a && o.b > v.b)) v = o;
        else if (o.a == v.a && o.b == v.b && o.pos != v.pos && o.pos != INT_MAX && v.pos != INT_MAX) {
            const int tie = ((o.pos >= 0 && v.pos >= 0) ? RATIO_TIE : 0) | ((o.aux | v.aux) & RATIO_TIE);
            if (o.pos < v.pos) v = o;
            v.aux |= tie;
        }
    }
};
struct CombRatio2 { /* larger b (|alfa|), then smaller pos */
    __device__ void operator()(Key &v, const Key &o) const
    {
        if (o.b > v.b) v = o;
        else if (o.b == v.b && o.pos != v.pos && o.pos != INT_MAX && v.pos != INT_MAX) {
            const int tie = ((o.pos >= 0 && v.pos >= 0) ? RATIO_TIE : 0) | ((o.aux | v.aux) & RATIO_TIE);
            if (o.pos < v.pos) v = o;
            v.aux |= tie;
        }
    }
};
struct CombArgMaxCnt { /* CombArgMax on (a, pos); aux is a count that is summed */
    __device__ void operator()(Key &v, const Key &o) const
    {
        const int cnt = v.aux + o.aux;
        if (o.a > v.a || (o.a == v.a && o.pos < v.pos)) v = o;
        v.aux = cnt;
    }
};
struct CombSum2 { /* a and b are sums, c is a maximum, aux a count, pos unused */
    __device__ void operator()(Key &v, const Key &o) const
    {
        v.a += o.a; v.b += o.b; v.c = fmax(v.c, o.c); v.aux += o.aux;
    }
};

__device__ __forceinline__ void atomic_max_abs(double *addr, double x)
{
    atomicMax((unsigned long long *)addr, (unsigned long long)__double_as_longlong(fabs(x)));
}

__device__ __forceinline__ double relax(double rtol, double bnd)
{
    /* rtol * (1.0 + kappa * |bnd|) without contraction */
    return __dmul_rn(rtol, __dadd_rn(1.0, __dmul_rn(GLPB_KAPPA, fabs(bnd))));
}

__device__ __forceinline__ double get_xN(const signed char *stat, const int *head, const double *lb,
                                         const double *ub, int m, int j)
{
    int k = head[m + j];
    switch (stat[j]) {
    case GLP_NL: return lb[k];
    case GLP_NU: return ub[k];
    case GLP_NF: return 0.0;
    default: return lb[k]; /* GLP_NS */
    }
}

__device__ __forceinline__ bool gamma_on(const Ctrl *ctrl) { return ctrl->pse && ctrl->refct > 0; }

/* end-of-iteration bookkeeping (it_cnt++, rigorous--, state flags, refct,
   update count) and the conditions under which the host must step in before
   the next iteration; runs in one thread of the last kernel of an iteration */
__device__ void iter_end(Ctrl *ctrl, bool basis_changed, int dual)
{
    if (ctrl->piv_log && ctrl->it_cnt >= 0 && ctrl->it_cnt < ctrl->piv_cap) {
        ctrl->piv_log[2 * ctrl->it_cnt] = ctrl->q;
        ctrl->piv_log[2 * ctrl->it_cnt + 1] = ctrl->p;
    }
    ctrl->it_cnt++;
    ctrl->n_done++;
    if (ctrl->rigorous > 0) ctrl->rigorous--;
    ctrl->bbar_fresh = 0;
    if (basis_changed) {
        ctrl->cbar_fresh = 0;
        ctrl->binv_fresh = 0;
        ctrl->upd_cnt++;
        if (gamma_on(ctrl)) ctrl->refct--;
    }
    ctrl->big = 0.0;
    ctrl->skip2 = 0;
    if (basis_changed && ctrl->upd_cnt >= ctrl->period) ctrl->status = ST_REFAC;
    else if (ctrl->it_cnt >= ctrl->it_max) ctrl->status = ST_LIMIT;
    else if (ctrl->pse && ctrl->refct == 0) ctrl->status = ST_REFSP;
    else if (dual && ctrl->phase == 2 &&
             ((ctrl->zeta < 0.0 && ctrl->obj_ll > -DBL_MAX && ctrl->obj <= ctrl->obj_ll) ||
              (ctrl->zeta > 0.0 && ctrl->obj_ul < +DBL_MAX && ctrl->obj >= ctrl->obj_ul)))
        ctrl->status = ST_OBJLIM;
}

/* ------------------------------------------------------------------ */
/* pricing                                                            */
/* ------------------------------------------------------------------ */

/* chuzc (primal pricing), lib/glpspx01.js:646-688: the scan over columns
   start, start+stride, ...  Shared by the stand-alone kernel and the persistent
   iteration engine; (jx, stx) overrides the status of one column whose header
   update is still in flight (engine), jx = -1 otherwise. */
__device__ __forceinline__ void price_primal(Key &v, int j, int st, double dj, double gam, double tol_dj)
{
    bool ok;
    switch (st) {
    case GLP_NL: ok = !(dj >= -tol_dj); break;
    case GLP_NU: ok = !(dj <= +tol_dj); break;
    case GLP_NF: ok = !(-tol_dj <= dj && dj <= +tol_dj); break;
    default: ok = false;
    }
    if (ok) {
        double temp = __ddiv_rn(__dmul_rn(dj, dj), gam);
        if (v.a < temp || (v.a == temp && temp > 0.0 && j < v.pos)) { v.a = temp; v.pos = j; }
    }
}

__device__ __forceinline__ void scan_chuzc_primal(Key &v, int start, int stride, int n,
                                                  const signed char *stat, const double *cbar,
                                                  const double *gamma, double tol_dj, int jx, int stx)
{
    for (int j = start; j < n; j += stride)
        price_primal(v, j, j == jx ? stx : (int)stat[j], cbar[j], gamma[j], tol_dj);
}

/* The same scan with 16-byte loads: a thread takes 4 consecutive columns per step (stat as
   one 4-byte word, cbar and gamma as two double2 each: 68 bytes in flight per thread and step),
   steps of one thread are independent so several are in flight.  The arrays must be 16-byte
   (stat: 4-byte) aligned -- the handle's own arrays are; otherwise the scalar scan is used.
   Selection is index-aware, so the visiting order does not matter. */
__device__ __forceinline__ void scan_chuzc_primal_v4(Key &v, int tid, int nthreads, int n,
                                                     const signed char *stat, const double *cbar,
                                                     const double *gamma, double tol_dj)
{
    const int n4 = n >> 2;
    const uchar4 *s4 = reinterpret_cast<const uchar4 *>(stat);
    const double2 *c2 = reinterpret_cast<const double2 *>(cbar);
    const double2 *g2 = reinterpret_cast<const double2 *>(gamma);
#pragma unroll 2
    for (int q = tid; q < n4; q += nthreads) {
        const uchar4 st = __ldg(s4 + q);
        const double2 ca = __ldg(c2 + 2 * q), cb = __ldg(c2 + 2 * q + 1);
        const double2 ga = __ldg(g2 + 2 * q), gb = __ldg(g2 + 2 * q + 1);
        const int j = q << 2;
        price_primal(v, j, (signed char)st.x, ca.x, ga.x, tol_dj);
        price_primal(v, j + 1, (signed char)st.y, ca.y, ga.y, tol_dj);
        price_primal(v, j + 2, (signed char)st.z, cb.x, gb.x, tol_dj);
        price_primal(v, j + 3, (signed char)st.w, cb.y, gb.y, tol_dj);
    }
    for (int j = (n4 << 2) + tid; j < n; j += nthreads) price_primal(v, j, stat[j], cbar[j], gamma[j], tol_dj);
}

/* Algorithmic bytes: 17 per column (stat 1 + cbar 8 + gamma 8). */
__global__ void k_chuzc_primal(Ctrl *ctrl, int n, const signed char *__restrict__ stat,
                               const double *__restrict__ cbar, const double *__restrict__ gamma,
                               double tol_dj, int set_status, Key *scratch)
{
    if (ctrl->status != ST_OK) return;
    Key none = {0.0, 0.0, 0.0, INT_MAX, 0};
    Key v = none;
    const bool aligned = (((uintptr_t)stat & 3) | ((uintptr_t)cbar & 15) | ((uintptr_t)gamma & 15)) == 0;
    if (aligned)
        scan_chuzc_primal_v4(v, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, n, stat, cbar, gamma, tol_dj);
    else
    scan_chuzc_primal(v, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, n, stat, cbar, gamma,
                      tol_dj, -1, 0);
    grid_reduce(v, none, scratch, &ctrl->ticket[0], CombArgMax(), [=](const Key &r) {
        ctrl->q = (r.a > 0.0 && r.pos != INT_MAX) ? r.pos : P_NONE;
        ctrl->big = 0.0;
        if (ctrl->q == P_NONE && set_status) ctrl->status = ST_NONE1;
    });
}

/* chuzr (dual pricing), lib/glpspx02.js:572-625: the scan over basic positions.
   (ix, kx) overrides head[ix] (engine: header update in flight), ix = -1 otherwise. */
__device__ __forceinline__ void price_dual(Key &v, int i, int t, double l, double u, double bi, double g,
                                           double tol_bnd)
{
    double ri = 0.0;
    if (t == GLP_LO || t == GLP_DB || t == GLP_FX) {
        double eps = relax(tol_bnd, l);
        if (bi < __dsub_rn(l, eps)) ri = __dsub_rn(l, bi);
    }
    if (t == GLP_UP || t == GLP_DB || t == GLP_FX) {
        double eps = relax(tol_bnd, u);
        if (bi > __dadd_rn(u, eps)) ri = __dsub_rn(u, bi);
    }
    if (ri != 0.0) {
        if (g < DBL_EPSILON) g = DBL_EPSILON;
        double temp = __ddiv_rn(__dmul_rn(ri, ri), g);
        if (v.a < temp || (v.a == temp && temp > 0.0 && i < v.pos)) { v.a = temp; v.b = ri; v.pos = i; }
    }
}

__device__ __forceinline__ void scan_chuzr_dual(Key &v, int start, int stride, int m,
                                                const signed char *type, const double *lb,
                                                const double *ub, const int *head, const double *bbar,
                                                const double *gamma, double tol_bnd, int ix, int kx)
{
    for (int i = start; i < m; i += stride) {
        int k = (i == ix) ? kx : head[i];
        price_dual(v, i, type[k], lb[k], ub[k], bbar[i], gamma[i], tol_bnd);
    }
}

/* Algorithmic bytes: 37 per row (head 4, type 1, lb 8, ub 8, bbar 8, gamma 8). */
__global__ void k_chuzr_dual(Ctrl *ctrl, int m, const signed char *__restrict__ type,
                             const double *__restrict__ lb, const double *__restrict__ ub,
                             const int *__restrict__ head, const double *__restrict__ bbar,
                             const double *__restrict__ gamma, double tol_bnd, int set_status,
                             Key *scratch)
{
    if (ctrl->status != ST_OK) return;
    Key none = {0.0, 0.0, 0.0, INT_MAX, 0};
    Key v = none;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if ((((uintptr_t)head & 15) | ((uintptr_t)bbar & 15) | ((uintptr_t)gamma & 15)) == 0) {
        /* 4 basic positions per step: head as one int4, bbar / gamma as two double2 each; the 12
           gathers of type / lb / ub through head are issued together before any is used */
        const int m4 = m >> 2;
        const int4 *h4 = reinterpret_cast<const int4 *>(head);
        const double2 *b2 = reinterpret_cast<const double2 *>(bbar), *g2 = reinterpret_cast<const double2 *>(gamma);
        for (int q = tid; q < m4; q += nth) {
            const int4 k = __ldg(h4 + q);
            const double2 ba = __ldg(b2 + 2 * q), bb = __ldg(b2 + 2 * q + 1), ga = __ldg(g2 + 2 * q), gb = __ldg(g2 + 2 * q + 1);
            const int t0 = type[k.x], t1 = type[k.y], t2 = type[k.z], t3 = type[k.w];
            const double l0 = lb[k.x], l1 = lb[k.y], l2 = lb[k.z], l3 = lb[k.w];
            const double u0 = ub[k.x], u1 = ub[k.y], u2 = ub[k.z], u3 = ub[k.w];
            const int i = q << 2;
            price_dual(v, i, t0, l0, u0, ba.x, ga.x, tol_bnd);
            price_dual(v, i + 1, t1, l1, u1, ba.y, ga.y, tol_bnd);
            price_dual(v, i + 2, t2, l2, u2, bb.x, gb.x, tol_bnd);
            price_dual(v, i + 3, t3, l3, u3, bb.y, gb.y, tol_bnd);
        }
        scan_chuzr_dual(v, (m4 << 2) + tid, nth, m, type, lb, ub, head, bbar, gamma, tol_bnd, -1, 0);
    } else
    scan_chuzr_dual(v, tid, nth, m, type, lb, ub, head, bbar, gamma, tol_bnd, -1, 0);
    grid_reduce(v, none, scratch, &ctrl->ticket[0], CombArgMax(), [=](const Key &r) {
        bool found = (r.a > 0.0 && r.pos != INT_MAX);
        ctrl->p = found ? r.pos : P_NONE;
        ctrl->delta = found ? r.b : 0.0;
        ctrl->big = 0.0;
        if (!found && set_status) ctrl->status = ST_NONE1;
    });
}

/* ------------------------------------------------------------------ */
/* ratio tests                                                        */
/* ------------------------------------------------------------------ */

/* chuzr (primal ratio test), lib/glpspx01.js:808-1028: scan of one pass over
   list positions start, start+stride, ...  ind == NULL: dense mode over
   positions 0..num-1 with the significance mask |tcol| >= eps; ind != NULL:
   the caller's sorted list (kernel parity).  Ties are broken by list position,
   as the sequential loop does.  Shared by the stand-alone kernel and the engine.
   Algorithmic bytes: 49 per examined entry and pass. */
/* one candidate of the primal ratio test: both the ratio against the relaxed
   bound (pass 1) and against the exact bound (pass 2), the pivot |alfa|, the
   status the leaving variable would take, the signed tcol entry and whether the
   variable is fixed.  Returns false if position i gives no ratio. */
struct RatioCand {
    double t1, t2, aabs, tc;
    int i_stat, fx;
};
__device__ __forceinline__ bool ratio_primal_elem(RatioCand &C, int i, int phase, double s, double eps, double rtol,
                                                  const signed char *type, const double *lb, const double *ub,
                                                  const double *coef, const int *head, const double *bbar,
                                                  const double *tcol)
{
    const double tc = tcol[i];
    if (tc == 0.0 || fabs(tc) < eps) return false;
    const int k = head[i];
    const double alfa = __dmul_rn(s, tc), bi = bbar[i];
    const int tk = type[k];
    const double ck = (phase == 1 ? coef[k] : 0.0);
    double b, r1;
    int i_stat;
    if (alfa > 0.0) {
        if (phase == 1 && ck < 0.0) { b = lb[k]; r1 = __dadd_rn(b, relax(rtol, b)); i_stat = GLP_NL; }
        else if (phase == 1 && ck > 0.0) return false;
        else if (tk == GLP_UP || tk == GLP_DB || tk == GLP_FX) { b = ub[k]; r1 = __dadd_rn(b, relax(rtol, b)); i_stat = GLP_NU; }
        else return false;
    } else {
        if (phase == 1 && ck > 0.0) { b = ub[k]; r1 = __dsub_rn(b, relax(rtol, b)); i_stat = GLP_NU; }
        else if (phase == 1 && ck < 0.0) return false;
        else if (tk == GLP_LO || tk == GLP_DB || tk == GLP_FX) { b = lb[k]; r1 = __dsub_rn(b, relax(rtol, b)); i_stat = GLP_NL; }
        else return false;
    }
    double t1 = __ddiv_rn(__dsub_rn(r1, bi), alfa), t2 = __ddiv_rn(__dsub_rn(b, bi), alfa);
    if (t1 < 0.0) t1 = 0.0;
    if (t2 < 0.0) t2 = 0.0;
    C.t1 = t1; C.t2 = t2; C.aabs = fabs(alfa); C.tc = tc; C.i_stat = i_stat; C.fx = (tk == GLP_FX);
    return true;
}

/* key of a candidate: a = ratio, b = |alfa|, c = signed tcol entry, aux = status (+256 if the variable is fixed) */
__device__ __forceinline__ Key ratio_primal_key(const RatioCand &C, int pass, int pos)
{
    Key c = {pass == 1 ? C.t1 : C.t2, C.aabs, C.tc, pos, C.i_stat | (C.fx ? 256 : 0)};
    return c;
}

__device__ __forceinline__ void scan_ratio_primal(Key &v, int start, int stride, int pass, int phase,
                                                  double s, double eps, double tmax, double rtol,
                                                  const signed char *type, const double *lb,
                                                  const double *ub, const double *coef, const int *head,
                                                  const double *bbar, const double *tcol,
                                                  const int *ind, int num)
{
    for (int pos = start; pos < num; pos += stride) {
        const int i = ind ? ind[pos] : pos;
        RatioCand C;
        if (!ratio_primal_elem(C, i, phase, s, eps, rtol, type, lb, ub, coef, head, bbar, tcol)) continue;
        const Key c = ratio_primal_key(C, pass, pos);
        if (pass == 1) CombRatio1()(v, c);
        else if (c.a <= tmax) CombRatio2()(v, c);
    }
}

/* what the reference does with the winner of a pass (chuzr, :1009-1028, and
   the pivot-size test of the main loop, :1960-1975) */
__device__ __forceinline__ void fin_ratio_primal(Ctrl *ctrl, const Key &r, int pass, double s, double rtol,
                                                 const int *ind, bool tie_stop = false)
{
    int p, p_stat = r.aux & 255;
    double teta = r.a;
    if (r.pos == INT_MAX) p = P_NONE;
    else if (r.pos == -1) p = P_FLIP;
    else p = ind ? ind[r.pos] : r.pos;
    if (pass == 1) {
        ctrl->skip2 = (rtol == 0.0 || p < 0 || teta == 0.0);
        ctrl->tmax = teta;
        if (!ctrl->skip2) { ctrl->p = p; return; }
    }
    /* the decisive pass ended in an exact tie that only the reference's list order settles */
    if (tie_stop && ind == nullptr && (r.aux & RATIO_TIE)) { ctrl->p = p; ctrl->status = ST_TIE; return; }
    if (p >= 0 && (r.aux & 256)) p_stat = GLP_NS;        /* a fixed variable leaves: lib/glpspx01.js:1019 */
    ctrl->p = p;
    ctrl->p_stat = p_stat;
    ctrl->teta = __dmul_rn(s, teta);
    if (p == P_NONE) { ctrl->status = ST_NONE2; return; }
    if (p >= 0) {
        double piv = r.c;                                  /* tcol[p], carried by the key */
        double e5 = 1e-5 * (1.0 + 0.01 * ctrl->tcol_max);
        ctrl->piv1 = piv;
        if (fabs(piv) < e5 && !ctrl->rigorous) ctrl->status = ST_PIVSMALL;
    }
}

__device__ __forceinline__ void ratio_primal_body(Ctrl *ctrl, int pass, int m, const signed char *__restrict__ type,
                               const double *__restrict__ lb, const double *__restrict__ ub,
                               const double *__restrict__ coef, const int *__restrict__ head,
                               const double *__restrict__ bbar, const double *__restrict__ tcol,
                               const int *__restrict__ ind, int num, double rtol,
                               Key *scratch)
{
    if (ctrl->status != ST_OK) return;
    if (pass == 2 && ctrl->skip2) return;
    const int q = ctrl->q, phase = ctrl->phase;
    if (num < 0) num = ctrl->list_num;                /* the list k_sort_list left */
    const double s = (ctrl->d1 > 0.0 ? -1.0 : +1.0); /* d1 holds the (corrected) cbar[q] */
    const double eps = (ind == nullptr ? ctrl->eps : 0.0);
    const double tmax = ctrl->tmax;
    const int kq = head[m + q];
    Key none = {DBL_MAX, 0.0, 0.0, INT_MAX, 0};
    Key v = none;
    if (pass == 1 && blockIdx.x == 0 && threadIdx.x == 0 && type[kq] == GLP_DB) {
        v.a = __dsub_rn(ub[kq], lb[kq]); v.b = 1.0; v.pos = -1; v.aux = 0;
    }
    scan_ratio_primal(v, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, pass, phase, s, eps, tmax,
                      rtol, type, lb, ub, coef, head, bbar, tcol, ind, num);
    auto fin = [=](const Key &r) { fin_ratio_primal(ctrl, r, pass, s, rtol, ind); };
    if (pass == 1) grid_reduce(v, none, scratch, &ctrl->ticket[1], CombRatio1(), fin);
    else grid_reduce(v, none, scratch, &ctrl->ticket[1], CombRatio2(), fin);
}

/* pass = 1 or 2: one pass per launch (any grid); pass = 0: both passes in one
   launch of a single block (vectors up to a few 10^4 entries) */
__global__ void k_ratio_primal(Ctrl *ctrl, int pass, int m, const signed char *__restrict__ type,
                               const double *__restrict__ lb, const double *__restrict__ ub,
                               const double *__restrict__ coef, const int *__restrict__ head,
                               const double *__restrict__ bbar, const double *__restrict__ tcol,
                               const int *__restrict__ ind, int num, double rtol, Key *scratch)
{
    if (pass != 0) { ratio_primal_body(ctrl, pass, m, type, lb, ub, coef, head, bbar, tcol, ind, num, rtol, scratch); return; }
    ratio_primal_body(ctrl, 1, m, type, lb, ub, coef, head, bbar, tcol, ind, num, rtol, scratch);
    __syncthreads();
    ratio_primal_body(ctrl, 2, m, type, lb, ub, coef, head, bbar, tcol, ind, num, rtol, scratch);
}

/* chuzc (dual ratio test), lib/glpspx02.js:793-935: scan of one pass.
   Algorithmic bytes: 21 per examined entry and pass (idx 4, trow 8, stat 1, cbar 8). */
struct RatioCandD {
    double t1, t2, aabs, tr;
};
__device__ __forceinline__ bool ratio_dual_elem(RatioCandD &C, int j, double s, double eps, double rtol,
                                                const signed char *stat, const double *cbar, const double *trow)
{
    const double tr = trow[j];
    if (tr == 0.0 || fabs(tr) < eps) return false;
    const double alfa = __dmul_rn(s, tr);
    const int st = stat[j];
    double d = cbar[j], r1;
    if (alfa > 0.0) {
        if (st == GLP_NL || st == GLP_NF) r1 = __dadd_rn(d, rtol);
        else return false;
    } else {
        if (st == GLP_NU || st == GLP_NF) r1 = __dsub_rn(d, rtol);
        else return false;
    }
    double t1 = __ddiv_rn(r1, alfa), t2 = __ddiv_rn(d, alfa);
    if (t1 < 0.0) t1 = 0.0;
    if (t2 < 0.0) t2 = 0.0;
    C.t1 = t1; C.t2 = t2; C.aabs = fabs(alfa); C.tr = tr;
    return true;
}

__device__ __forceinline__ void scan_ratio_dual(Key &v, int start, int stride, int pass, double s, double eps,
                                                double tmax, double rtol, const signed char *stat,
                                                const double *cbar, const double *trow, const int *ind,
                                                int num)
{
    for (int pos = start; pos < num; pos += stride) {
        const int j = ind ? ind[pos] : pos;
        RatioCandD C;
        if (!ratio_dual_elem(C, j, s, eps, rtol, stat, cbar, trow)) continue;
        const Key c = {pass == 1 ? C.t1 : C.t2, C.aabs, C.tr, pos, 0};       /* c = signed trow entry */
        if (pass == 1) CombRatio1()(v, c);
        else if (c.a <= tmax) CombRatio2()(v, c);
    }
}

__device__ __forceinline__ void fin_ratio_dual(Ctrl *ctrl, const Key &r, int pass, double s, double rtol,
                                               const int *ind, bool tie_stop = false)
{
    int q = (r.pos == INT_MAX) ? P_NONE : (ind ? ind[r.pos] : r.pos);
    double teta = r.a;
    if (pass == 1) {
        ctrl->skip2 = (rtol == 0.0 || q < 0 || teta == 0.0);
        ctrl->tmax = teta;
        if (!ctrl->skip2) { ctrl->q = q; return; }
    }
    if (tie_stop && ind == nullptr && (r.aux & RATIO_TIE)) { ctrl->q = q; ctrl->status = ST_TIE; return; }
    ctrl->q = q;
    ctrl->new_dq = __dmul_rn(s, teta);
    if (q == P_NONE) { ctrl->status = ST_NONE2; return; }
    double piv = r.c;                                      /* trow[q], carried by the key */
    double e5 = 1e-5 * (1.0 + 0.01 * ctrl->trow_max);
    ctrl->piv2 = piv;
    if (fabs(piv) < e5 && !ctrl->rigorous) ctrl->status = ST_PIVSMALL;
}

__device__ __forceinline__ void ratio_dual_body(Ctrl *ctrl, int pass, const signed char *__restrict__ stat,
                             const double *__restrict__ cbar, const double *__restrict__ trow,
                             const int *__restrict__ ind, int num, double rtol,
                             Key *scratch)
{
    if (ctrl->status != ST_OK) return;
    if (pass == 2 && ctrl->skip2) return;
    const double s = (ctrl->delta > 0.0 ? +1.0 : -1.0);
    if (num < 0) num = ctrl->list_num;                /* the list k_sort_list left */
    const double eps = (ind == nullptr ? ctrl->eps : 0.0);
    const double tmax = ctrl->tmax;
    Key none = {DBL_MAX, 0.0, 0.0, INT_MAX, 0};
    Key v = none;
    scan_ratio_dual(v, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, pass, s, eps, tmax, rtol,
                    stat, cbar, trow, ind, num);
    auto fin = [=](const Key &r) { fin_ratio_dual(ctrl, r, pass, s, rtol, ind); };
    if (pass == 1) grid_reduce(v, none, scratch, &ctrl->ticket[1], CombRatio1(), fin);
    else grid_reduce(v, none, scratch, &ctrl->ticket[1], CombRatio2(), fin);
}

__global__ void k_ratio_dual(Ctrl *ctrl, int pass, const signed char *__restrict__ stat,
                             const double *__restrict__ cbar, const double *__restrict__ trow,
                             const int *__restrict__ ind, int num, double rtol, Key *scratch)
{
    if (pass != 0) { ratio_dual_body(ctrl, pass, stat, cbar, trow, ind, num, rtol, scratch); return; }
    ratio_dual_body(ctrl, 1, stat, cbar, trow, ind, num, rtol, scratch);
    __syncthreads();
    ratio_dual_body(ctrl, 2, stat, cbar, trow, ind, num, rtol, scratch);
}

/* sort_tcol / sort_trow (lib/glpspx01.js:773-806, lib/glpspx02.js:754-791) in closed form.
   The reference lists the non-zeros of the vector in ascending index order (eval_tcol,
   eval_trow1/2) and then moves the significant ones (|v| >= eps) to the front with a loop that
   always examines the LAST entry of the shrinking list: a significant entry is swapped with the
   entry at the growing front, an insignificant one is dropped.  Entries are therefore drawn
   from the back after an insignificant one and from the front after a significant one, which
   gives, with Ns significant entries and list indices j = 1, 2, ...:
     - a significant entry with j <= Ns - 1 ends up at place j + 1;
     - the other significant entries take the remaining places in DESCENDING index order,
       the remaining places being 1 and every j + 1 whose entry j <= Ns - 1 is insignificant.
   One CTA; two chunked scans of the vector.  list[0..num) = indices in the reference's order,
   ctrl->list_num = num = Ns.  tmp: 2 N + 2 ints. */
__global__ void k_sort_list(Ctrl *ctrl, const double *__restrict__ vec, int N, int *__restrict__ list, int *__restrict__ tmp)
{
    if (ctrl->status != ST_OK) return;
    __shared__ int wsum[32];
    __shared__ int s_tot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x, nw = nt >> 5;
    const double eps = ctrl->eps;
    /* block-wide inclusive scan of a packed pair of counters (non-zero: low 16 bits, significant: high) */
    auto scan = [&](int x, int &total) {
        int v = x;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int o = __shfl_up_sync(FULLMASK, v, off);
            if (lane >= off) v += o;
        }
        __syncthreads();
        if (lane == 31) wsum[warp] = v;
        __syncthreads();
        if (warp == 0) {
            int w = (lane < nw) ? wsum[lane] : 0, ws = w;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int o = __shfl_up_sync(FULLMASK, ws, off);
                if (lane >= off) ws += o;
            }
            wsum[lane] = ws - w;
            if (lane == 31) s_tot = ws;
        }
        __syncthreads();
        total = s_tot;
        return v + wsum[warp];
    };
    /* pass 1: Ns */
    int cnt = 0;
    for (int i = tid; i < N; i += nt) {
        const double v = vec[i];
        cnt += (v != 0.0 && !(fabs(v) < eps)) ? 1 : 0;
    }
    int Ns;
    {
        /* counts per thread can exceed 16 bits: plain sum by the same scan in two halves */
        int t0, t1;
        scan(cnt & 0xffff, t0);
        scan(cnt >> 16, t1);
        Ns = t0 + (t1 << 16);
    }
    const int F = Ns - 1;
    int *freerank = tmp, *backelem = tmp + N + 1;          /* 1-based, [r] */
    /* pass 2 */
    int carry_nz = 0, carry_sg = 0, my_back = 0;
    for (int base = 0; base < N; base += nt) {
        const int i = base + tid;
        const double v = (i < N) ? vec[i] : 0.0;
        const int nz = (v != 0.0) ? 1 : 0, sg = (nz && !(fabs(v) < eps)) ? 1 : 0;
        int tot;
        const int inc = scan(nz | (sg << 16), tot);
        const int j = carry_nz + (inc & 0xffff), sgi = carry_sg + (inc >> 16);
        if (nz) {
            if (sg) {
                if (j <= F) list[j] = i;                   /* place j + 1 */
                else { backelem[Ns - (sgi - 1)] = i; my_back++; }   /* r = significant entries with list index >= j */
            } else if (j <= F)
                freerank[(j - sgi) + 1] = j + 1;           /* the (t+1)-th free place, t = its ordinal among the insignificant */
        }
        carry_nz += tot & 0xffff; carry_sg += tot >> 16;
    }
    /* the entries drawn from the back take the free places in turn: r = 1 -> place 1 */
    int nback, nback_hi;
    scan(my_back & 0xffff, nback);
    scan(my_back >> 16, nback_hi);
    nback += nback_hi << 16;
    __syncthreads();
    for (int r = 1 + tid; r <= nback; r += nt) {
        const int place = (r == 1) ? 1 : freerank[r];
        list[place - 1] = backelem[r];
    }
    if (tid == 0) ctrl->list_num = Ns;
}

/* ------------------------------------------------------------------ */
/* sparse passes over A (all gathers: no atomics, deterministic)       */
/* ------------------------------------------------------------------ */

template <int G> __device__ __forceinline__ double group_sum(double v)
{
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) v += __shfl_down_sync(FULLMASK, v, off, G);
    return v;
}

/* Pivot row, lib/glpspx01.js:1058-1098 and lib/glpspx02.js:655-693:
       trow[j] = -rho' N_j   for every non-basic, non-fixed j
   as dot products over the CSC columns, G lanes per column.  With u != NULL
   it also returns the PSE inner products s_j = N_j' u of update_gamma
   (lib/glpspx01.js:1231-1241) from the same pass over A.
   Algorithmic bytes: 12 per non-zero of a non-basic column + 13 n + 8 m. */
template <int G>
__global__ void k_trow(Ctrl *ctrl, int m, int n, const int *__restrict__ a_ptr,
                       const int *__restrict__ a_ind, const double *__restrict__ a_val,
                       const int *__restrict__ head, const signed char *__restrict__ stat,
                       const double *__restrict__ rho, const double *u,
                       double *__restrict__ trow, double *__restrict__ svec, int want_max)
{
    if (ctrl->status != ST_OK) return;
    if (u != nullptr && !gamma_on(ctrl)) u = nullptr;   /* PSE inner products only while weights are live */
    /* every group of G lanes takes TROW_C columns per step and walks them in lock step, so the
       dependent chain stat -> head -> a_ptr -> (a_ind, a_val) -> rho of one column overlaps with
       the chains of the others (the pass is latency-, not bandwidth-limited otherwise) */
    constexpr int C = 4;
    /* launched with m*8 bytes of dynamic shared memory (want_max & 2): rho is staged there once per
       CTA, so that the 16 gathers per column hit shared memory instead of L1/L2 */
    extern __shared__ double trow_sh_rho[];
    if (want_max & 2) {
        for (int r = threadIdx.x; r < m; r += blockDim.x) trow_sh_rho[r] = rho[r];
        __syncthreads();
        rho = trow_sh_rho;
    }
    const int ngroups = (gridDim.x * blockDim.x) / G;
    const int g0 = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x % G;
    double amax = 0.0;
    for (int base = g0 * C;; base += ngroups * C) {
        if (__all_sync(FULLMASK, base >= n)) break;         /* the groups of a warp leave together: shuffles below */
        int beg[C], end[C], kk[C];
        double t[C], s[C];
#pragma unroll
        for (int c = 0; c < C; c++) {
            const int j = base + c;
            t[c] = s[c] = 0.0; beg[c] = end[c] = 0; kk[c] = -1;
            if (j < n && stat[j] != GLP_NS) kk[c] = head[m + j];
        }
#pragma unroll
        for (int c = 0; c < C; c++)
            if (kk[c] >= m) { beg[c] = a_ptr[kk[c] - m]; end[c] = a_ptr[kk[c] - m + 1]; }
            else if (kk[c] >= 0 && lane == 0) { t[c] = -rho[kk[c]]; if (u) s[c] = u[kk[c]]; }
        int len = 0;
#pragma unroll
        for (int c = 0; c < C; c++) len = max(len, end[c] - beg[c]);
        for (int o = lane; o < len; o += G) {
            int r[C]; double a[C];
#pragma unroll
            for (int c = 0; c < C; c++) {
                const bool in = beg[c] + o < end[c];
                r[c] = in ? a_ind[beg[c] + o] : -1;
                a[c] = in ? a_val[beg[c] + o] : 0.0;
            }
#pragma unroll
            for (int c = 0; c < C; c++)
                if (r[c] >= 0) { t[c] += rho[r[c]] * a[c]; if (u) s[c] -= a[c] * u[r[c]]; }
        }
#pragma unroll
        for (int c = 0; c < C; c++) {
            t[c] = group_sum<G>(t[c]);
            if (u) s[c] = group_sum<G>(s[c]);
            const int j = base + c;
            if (j < n && lane == 0) {
                trow[j] = t[c];
                if (u) svec[j] = s[c];
                amax = fmax(amax, fabs(t[c]));
            }
        }
    }
    if (want_max & 1) {
        /* |trow|_inf: warp max, then one atomic per block (bit pattern of a
           non-negative double orders like an unsigned integer) */
        __shared__ double wmax[32];
        double a = amax;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) a = fmax(a, __shfl_down_sync(FULLMASK, a, off));
        if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = a;
        __syncthreads();
        if (threadIdx.x < 32) {
            a = (threadIdx.x < (blockDim.x >> 5)) ? wmax[threadIdx.x] : 0.0;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) a = fmax(a, __shfl_down_sync(FULLMASK, a, off));
            if (threadIdx.x == 0 && a > 0.0) atomic_max_abs(&ctrl->big, a);
        }
    }
}

/* tail of FTRAN: x[i] for every basic position from y = T * h_N
   (see glpb_internal.cuh).  One group of G lanes per position, gathering the
   row of A that belongs to a basic auxiliary variable. */
template <int G>
__global__ void k_ftran_tail(Ctrl *ctrl, int m, const int *__restrict__ at_ptr,
                             const int *__restrict__ at_ind, const double *__restrict__ at_val,
                             const int *__restrict__ head, const int *__restrict__ bind,
                             const int *__restrict__ rslot, const double *__restrict__ h,
                             const double *__restrict__ y, double *__restrict__ x, int cond)
{
    if (ctrl->status != ST_OK) return;
    if (cond && !gamma_on(ctrl)) return;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x % G;
    double acc = 0.0;
    int kk = (i < m) ? head[i] : m;
    if (i < m && kk < m) {
        int beg = at_ptr[kk], end = at_ptr[kk + 1];
        for (int ptr = beg + lane; ptr < end; ptr += G) {
            int pb = bind[m + at_ind[ptr]];
            if (pb < m) acc += at_val[ptr] * y[rslot[pb]];
        }
    }
    acc = group_sum<G>(acc);
    if (i < m && lane == 0) x[i] = (kk < m) ? h[kk] + acc : y[rslot[i]];
}

/* head of BTRAN: w[b] = c[pos_b] + sum_{r in R_B} A[r, j_b] c[bind[r]] */
template <int G>
__global__ void k_btran_head(Ctrl *ctrl, int m, const int *__restrict__ a_ptr,
                             const int *__restrict__ a_ind, const double *__restrict__ a_val,
                             const int *__restrict__ head, const int *__restrict__ bind,
                             const int *__restrict__ slot_pos, const double *__restrict__ c,
                             double *__restrict__ w, int cond)
{
    if (ctrl->status != ST_OK) return;
    if (cond && !gamma_on(ctrl)) return;
    const int k = ctrl->k;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x % G;
    double acc = 0.0;
    int i = 0;
    if (b < k) {
        i = slot_pos[b];
        int j = head[i] - m;
        int beg = a_ptr[j], end = a_ptr[j + 1];
        for (int ptr = beg + lane; ptr < end; ptr += G) {
            int pr = bind[a_ind[ptr]];
            if (pr < m) acc += a_val[ptr] * c[pr];
        }
    }
    acc = group_sum<G>(acc);
    if (b < k && lane == 0) w[b] = c[i] + acc;
}

__global__ void k_btran_tail(Ctrl *ctrl, int m, const int *__restrict__ cslot,
                             const int *__restrict__ bind, const double *__restrict__ c,
                             const double *__restrict__ zn, double *__restrict__ z, int cond)
{
    if (ctrl->status != ST_OK) return;
    if (cond && !gamma_on(ctrl)) return;
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    int cs = cslot[r];
    z[r] = (cs >= 0) ? zn[cs] : c[bind[r]];
}

/* residual of B x = h (error_ftran, lib/glpspx01.js:219-249), by rows */
template <int G>
__global__ void k_resid_ftran(Ctrl *ctrl, int m, const int *__restrict__ at_ptr,
                              const int *__restrict__ at_ind, const double *__restrict__ at_val,
                              const int *__restrict__ bind, const double *__restrict__ h,
                              const double *__restrict__ x, double *__restrict__ r)
{
    if (ctrl->status != ST_OK) return;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x % G;
    double acc = 0.0;
    if (row < m) {
        for (int ptr = at_ptr[row] + lane; ptr < at_ptr[row + 1]; ptr += G) {
            int pb = bind[m + at_ind[ptr]];
            if (pb < m) acc += at_val[ptr] * x[pb];
        }
    }
    acc = group_sum<G>(acc);
    if (row < m && lane == 0) {
        int pr = bind[row];
        r[row] = h[row] - (pr < m ? x[pr] : 0.0) + acc;
    }
}

/* residual of B' x = h (error_btran, lib/glpspx01.js:265-293), by columns */
template <int G>
__global__ void k_resid_btran(Ctrl *ctrl, int m, const int *__restrict__ a_ptr,
                              const int *__restrict__ a_ind, const double *__restrict__ a_val,
                              const int *__restrict__ head, const double *__restrict__ h,
                              const double *__restrict__ x, double *__restrict__ r)
{
    if (ctrl->status != ST_OK) return;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x % G;
    double acc = 0.0;
    int k = (i < m) ? head[i] : 0;
    if (i < m && k >= m) {
        for (int ptr = a_ptr[k - m] + lane; ptr < a_ptr[k - m + 1]; ptr += G)
            acc += a_val[ptr] * x[a_ind[ptr]];
    }
    acc = group_sum<G>(acc);
    if (i < m && lane == 0) r[i] = (k < m) ? h[i] - x[k] : h[i] + acc;
}

/* right-hand side of eval_beta: h = -N xN (lib/glpspx01.js:483-505), by rows */
template <int G>
__global__ void k_beta_rhs(Ctrl *ctrl, int m, int n, const int *__restrict__ at_ptr,
                           const int *__restrict__ at_ind, const double *__restrict__ at_val,
                           const int *__restrict__ head, const int *__restrict__ bind,
                           const signed char *__restrict__ stat, const double *__restrict__ lb,
                           const double *__restrict__ ub, double *__restrict__ h)
{
    if (ctrl->status != ST_OK) return;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x % G;
    double acc = 0.0;
    if (row < m) {
        for (int ptr = at_ptr[row] + lane; ptr < at_ptr[row + 1]; ptr += G) {
            int pb = bind[m + at_ind[ptr]];
            if (pb >= m) acc += at_val[ptr] * get_xN(stat, head, lb, ub, m, pb - m);
        }
    }
    acc = group_sum<G>(acc);
    if (row < m && lane == 0) {
        int pr = bind[row];
        h[row] = acc - (pr >= m ? get_xN(stat, head, lb, ub, m, pr - m) : 0.0);
    }
}

/* eval_cbar: d_j = c_k - N_j' pi (lib/glpspx01.js:531-584), by columns */
template <int G>
__global__ void k_cbar(Ctrl *ctrl, int m, int n, const int *__restrict__ a_ptr,
                       const int *__restrict__ a_ind, const double *__restrict__ a_val,
                       const int *__restrict__ head, const double *__restrict__ coef,
                       const double *__restrict__ pi, double *__restrict__ cbar)
{
    if (ctrl->status != ST_OK) return;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x % G;
    double acc = 0.0;
    int k = (j < n) ? head[m + j] : 0;
    if (j < n && k >= m) {
        for (int ptr = a_ptr[k - m] + lane; ptr < a_ptr[k - m + 1]; ptr += G)
            acc += a_val[ptr] * pi[a_ind[ptr]];
    }
    acc = group_sum<G>(acc);
    if (j < n && lane == 0) cbar[j] = (k < m) ? coef[k] - pi[k] : coef[k] + acc;
}

/* dual update_gamma, first half (lib/glpspx02.js:1103-1132): by rows,
       v[r] = sum_{j in C, non-basic} N_j[r] * trow_j */
template <int G>
__global__ void k_dual_gamma_rhs(Ctrl *ctrl, int m, const int *__restrict__ at_ptr,
                                 const int *__restrict__ at_ind, const double *__restrict__ at_val,
                                 const int *__restrict__ bind, const signed char *__restrict__ refsp,
                                 const double *__restrict__ trow, double *__restrict__ v, int cond)
{
    if (ctrl->status != ST_OK) return;
    if (cond && !gamma_on(ctrl)) return;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x % G;
    double acc = 0.0;
    if (row < m) {
        for (int ptr = at_ptr[row] + lane; ptr < at_ptr[row + 1]; ptr += G) {
            int jj = at_ind[ptr];
            int pb = bind[m + jj];
            if (pb >= m && refsp[m + jj]) acc -= trow[pb - m] * at_val[ptr];
        }
    }
    acc = group_sum<G>(acc);
    if (row < m && lane == 0) {
        int pr = bind[row];
        v[row] = acc + ((pr >= m && refsp[row]) ? trow[pr - m] : 0.0);
    }
}

/* ------------------------------------------------------------------ */
/* dense work on T                                                    */
/* ------------------------------------------------------------------ */

#define GEMV_TILE 128

/* y_part[s][b] = sum over a tile of columns  T[b,cs] * h[slot_row[cs]].
   Columns whose right-hand side is exactly zero are skipped, which makes the
   same kernel the sparse-rhs FTRAN of eval_tcol.  8 k^2 bytes when h is dense. */
__global__ void k_gemvN_part(Ctrl *ctrl, const double *__restrict__ T, int ldt,
                             const double *__restrict__ h, const int *__restrict__ slot_row,
                             double *__restrict__ part, int single, int cond)
{
    if (ctrl->status != ST_OK) return;
    if (cond && !gamma_on(ctrl)) return;
    const int k = ctrl->k;
    if ((int)(blockIdx.x * GEMV_TILE) >= k) return;
    __shared__ double hs[GEMV_TILE];
    const int b = blockIdx.x * GEMV_TILE + threadIdx.x;
    /* single != 0: one block walks all column tiles and writes y itself
       (small kernels: saves the second launch); else one tile per block */
    const int t0 = single ? 0 : blockIdx.y, t1 = single ? (k + GEMV_TILE - 1) / GEMV_TILE : blockIdx.y + 1;
    double acc = 0.0;
    for (int tile = t0; tile < t1; tile++) {
        const int c0 = tile * GEMV_TILE;
        if (c0 >= k) break;
        const int cn = min(GEMV_TILE, k - c0);
        __syncthreads();
        if ((int)threadIdx.x < cn) hs[threadIdx.x] = h[slot_row[c0 + threadIdx.x]];
        __syncthreads();
        if (b < k) {
            const double *col = T + (size_t)c0 * ldt + b;
#pragma unroll 4
            for (int c = 0; c < cn; c++) {
                double hv = hs[c];
                if (hv != 0.0) acc += col[(size_t)c * ldt] * hv;
            }
        }
    }
    if (b < k) part[(size_t)(single ? 0 : blockIdx.y) * ldt + b] = acc;
}

__global__ void k_gemvN_fin(Ctrl *ctrl, int ldt, const double *__restrict__ part,
                            double *__restrict__ y, int cond)
{
    if (ctrl->status != ST_OK) return;
    if (cond && !gamma_on(ctrl)) return;
    const int k = ctrl->k;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= k) return;
    const int S = (k + GEMV_TILE - 1) / GEMV_TILE;
    double acc = 0.0;
    for (int s = 0; s < S; s++) acc += part[(size_t)s * ldt + b];
    y[b] = acc;
}

/* zn[cs] = sum_b T[b,cs] * w[b]; one warp per column, coalesced */
__global__ void k_gemvT(Ctrl *ctrl, const double *__restrict__ T, int ldt,
                        const double *__restrict__ w, double *__restrict__ zn, int cond)
{
    if (ctrl->status != ST_OK) return;
    if (cond && !gamma_on(ctrl)) return;
    const int k = ctrl->k;
    const int cs = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (cs >= k) return;
    const double *col = T + (size_t)cs * ldt;
    double acc = 0.0;
    for (int b = lane; b < k; b += 32) acc += col[b] * w[b];
    acc = group_sum<32>(acc);
    if (lane == 0) zn[cs] = acc;
}

/* rho = row p of inv(B) (eval_rho, lib/glpspx01.js:1030-1042) read straight
   out of T instead of a BTRAN of e_p */
/* Fast form: row p of inv(B) restricted to R_N is T' w with
       w = e_b                       (a structural variable leaves)
       w[b] = A[kp, j_b]             (the auxiliary variable of row kp leaves)
   Every block rebuilds the sparse w in shared memory (zero + scatter of one
   CSR row), then each warp owns one column of T: coalesced dot products
   instead of a per-thread walk of the CSR row.  Rows whose auxiliary variable
   is basic are unit entries.  Dynamic shared memory: k doubles.
   Algorithmic bytes: 8 k^2 + 12 rowlen + 12 m. */
__global__ void k_rho_fast(Ctrl *ctrl, int m, const double *__restrict__ T, int ldt,
                           const int *__restrict__ at_ptr, const int *__restrict__ at_ind,
                           const double *__restrict__ at_val, const int *__restrict__ head,
                           const int *__restrict__ bind, const int *__restrict__ rslot,
                           const int *__restrict__ cslot, const int *__restrict__ slot_row,
                           double *__restrict__ rho)
{
    extern __shared__ double w[];
    if (ctrl->status != ST_OK) return;
    const int p = ctrl->p;
    if (p < 0) return;
    const int k = ctrl->k;
    const int kp = head[p];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int gid = blockIdx.x * nt + tid;
    if (gid < m && cslot[gid] < 0) rho[gid] = (bind[gid] == p) ? 1.0 : 0.0;
    if ((int)((blockIdx.x * nt) >> 5) >= k) return;      /* no column for this block */
    for (int b = tid; b < k; b += nt) w[b] = 0.0;
    __syncthreads();
    if (kp >= m) { if (tid == 0) w[rslot[p]] = 1.0; }
    else
        for (int ptr = at_ptr[kp] + tid; ptr < at_ptr[kp + 1]; ptr += nt) {
            int pb = bind[m + at_ind[ptr]];
            if (pb < m) w[rslot[pb]] = at_val[ptr];
        }
    __syncthreads();
    const int cs = gid >> 5, lane = tid & 31;
    if (cs >= k) return;
    const double *col = T + (size_t)cs * ldt;
    double acc = 0.0;
    for (int b = lane; b < k; b += 32) {
        double wb = w[b];
        if (wb != 0.0) acc += col[b] * wb;
    }
    acc = group_sum<32>(acc);
    if (lane == 0) rho[slot_row[cs]] = acc;
}

/* reference form (one thread per row, sequential walk); kept for kernels
   larger than the shared-memory budget of k_rho_fast */
__global__ void k_rho(Ctrl *ctrl, int m, const double *__restrict__ T, int ldt,
                      const int *__restrict__ at_ptr, const int *__restrict__ at_ind,
                      const double *__restrict__ at_val, const int *__restrict__ head,
                      const int *__restrict__ bind, const int *__restrict__ rslot,
                      const int *__restrict__ cslot, double *__restrict__ rho)
{
    if (ctrl->status != ST_OK) return;
    const int p = ctrl->p;
    if (p < 0) return;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const int cs = cslot[r];
    if (cs < 0) { rho[r] = (bind[r] == p) ? 1.0 : 0.0; return; }
    const int kp = head[p];
    const double *col = T + (size_t)cs * ldt;
    if (kp >= m) { rho[r] = col[rslot[p]]; return; }
    double acc = 0.0;
    for (int ptr = at_ptr[kp]; ptr < at_ptr[kp + 1]; ptr++) {
        int pb = bind[m + at_ind[ptr]];
        if (pb < m) acc += at_val[ptr] * col[rslot[pb]];
    }
    rho[r] = acc;
}

/* rank-1 part of a basis change:  inv(B)' = E inv(B)  restricted to T.
   For b, cs < k:  T[b,cs] -= (tcol[pos_b]/tcol_p) rho[row_cs]   (pos_b != p)
                   T[b,cs]  = -rho[row_cs]/tcol_p                (pos_b == p)
   Algorithmic bytes: 16 k^2 (read + write of T). */
#define UPD_TB 256
#define UPD_TC 16
__global__ void k_update_rank1(Ctrl *ctrl, double *__restrict__ T, int ldt,
                               const double *__restrict__ tcol, const double *__restrict__ rho,
                               const int *__restrict__ slot_pos, const int *__restrict__ slot_row)
{
    if (ctrl->status != ST_OK) return;
    const int p = ctrl->p;
    if (p < 0) return;
    const int k = ctrl->k;
    const int b = blockIdx.x * UPD_TB + threadIdx.x;
    const int c0 = blockIdx.y * UPD_TC;
    if (c0 >= k) return;
    __shared__ double rs[UPD_TC];
    const int cn = min(UPD_TC, k - c0);
    if ((int)threadIdx.x < cn) rs[threadIdx.x] = rho[slot_row[c0 + threadIdx.x]];
    __syncthreads();
    if (b >= k) return;
    const int i = slot_pos[b];
    const double tp = tcol[p];
    double *col = T + (size_t)c0 * ldt + b;
    if (i == p) {
        const double f = -1.0 / tp;
#pragma unroll 4
        for (int c = 0; c < cn; c++) col[(size_t)c * ldt] = rs[c] * f;
    } else {
        const double f = tcol[i] / tp;
        if (f != 0.0) {
#pragma unroll 4
            for (int c = 0; c < cn; c++) col[(size_t)c * ldt] -= f * rs[c];
        }
    }
}

/* the O(k) remainder of a basis change: append/remove a row and/or column of
   T, maintain the slot maps, swap head/bind and set the new non-basic status
   (change_basis, lib/glpspx01.js:1310-1371 / lib/glpspx02.js:1259-1294).
   One block. */
__global__ void k_update_fix(Ctrl *ctrl, int m, double *__restrict__ T, int ldt,
                             const double *__restrict__ tcol, const double *__restrict__ rho,
                             int *__restrict__ rslot, int *__restrict__ slot_pos,
                             int *__restrict__ cslot, int *__restrict__ slot_row,
                             int *__restrict__ head, int *__restrict__ bind,
                             signed char *__restrict__ stat, const signed char *__restrict__ type,
                             signed char *__restrict__ refsp, int dual)
{
    if (ctrl->status != ST_OK) return;
    const int p = ctrl->p, q = ctrl->q;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (p == P_FLIP) {
        if (tid == 0) {
            stat[q] = (stat[q] == GLP_NL) ? GLP_NU : GLP_NL;
            iter_end(ctrl, false, dual);
        }
        return;
    }
    const int k = ctrl->k;
    const int kp = head[p], kq = head[m + q];
    const bool LS = (kp < m), ES = (kq < m);
    const double tp = tcol[p];
    const int csq = ES ? cslot[kq] : -1;
    const int bp = LS ? -1 : rslot[p];
    /* 1. new column for row kp (it joins R_N): inv(B)'[i,kp] = -tcol_i/tcol_p,
          i != p; 1/d_p = -1/tcol_p at i = p */
    const int ctgt = LS ? (ES ? csq : k) : -1;
    const int bnew = (LS && !ES) ? k : -1;
    if (LS) {
        double *col = T + (size_t)ctgt * ldt;
        for (int b = tid; b < k; b += nt) col[b] = -tcol[slot_pos[b]] / tp;
        if (bnew >= 0 && tid == 0) col[bnew] = -1.0 / tp;
    }
    /* 2. new row for position p (it joins P_S) */
    if (bnew >= 0) {
        for (int cs = tid; cs < k; cs += nt)
            T[(size_t)cs * ldt + bnew] = -rho[slot_row[cs]] / tp;
    }
    __syncthreads();
    /* 3. removals: the last column/row moves into the hole */
    int knew = k;
    if (LS && !ES) knew = k + 1;
    if (!LS && ES) {
        knew = k - 1;
        /* column csq <- column k-1 ; row bp <- row k-1 */
        if (csq != k - 1) {
            double *dst = T + (size_t)csq * ldt;
            const double *src = T + (size_t)(k - 1) * ldt;
            for (int b = tid; b < k; b += nt) dst[b] = src[b];
        }
        __syncthreads();
        if (bp != k - 1) {
            for (int cs = tid; cs < k; cs += nt)
                T[(size_t)cs * ldt + bp] = T[(size_t)cs * ldt + (k - 1)];
        }
    }
    __syncthreads();
    if (tid == 0) {
        /* slot maps */
        if (LS) { cslot[kp] = ctgt; slot_row[ctgt] = kp; }
        if (ES) {
            cslot[kq] = -1;
            if (!LS) {
                if (csq != k - 1) { int rl = slot_row[k - 1]; slot_row[csq] = rl; cslot[rl] = csq; }
            }
        }
        if (bnew >= 0) { rslot[p] = bnew; slot_pos[bnew] = p; }
        if (!LS && ES) {
            rslot[p] = -1;
            if (bp != k - 1) { int pl = slot_pos[k - 1]; slot_pos[bp] = pl; rslot[pl] = bp; }
        }
        ctrl->k = knew;
        /* basis header */
        head[p] = kq; head[m + q] = kp;
        bind[kq] = p; bind[kp] = m + q;
        if (dual) {
            if (type[kp] == GLP_FX) stat[q] = GLP_NS;
            else stat[q] = (ctrl->delta > 0.0) ? GLP_NL : GLP_NU;
        } else
            stat[q] = (signed char)ctrl->p_stat;
        /* dual update_gamma: a fixed leaving variable drops out of the
           reference space (lib/glpspx02.js:1171-1173) */
        if (dual && gamma_on(ctrl) && type[kp] == GLP_FX && refsp[kp]) refsp[kp] = 0;
        iter_end(ctrl, true, dual);
    }
}

/* ------------------------------------------------------------------ */
/* refactorisation: slot maps (the inversion is in refactor.cuh)       */
/* ------------------------------------------------------------------ */

/* slot maps from head/bind: structural basic positions and non-basic rows in
   ascending order.  One block of 1024 threads, chunked scan. */
__global__ void k_build_slots(Ctrl *ctrl, int m, const int *__restrict__ head,
                              const int *__restrict__ bind, int *__restrict__ rslot,
                              int *__restrict__ slot_pos, int *__restrict__ cslot,
                              int *__restrict__ slot_row)
{
    __shared__ int cnt_r[1024], cnt_c[1024];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int chunk = (m + nt - 1) / nt;
    const int beg = min(m, tid * chunk), end = min(m, beg + chunk);
    int nr = 0, nc = 0;
    for (int i = beg; i < end; i++) {
        if (head[i] >= m) nr++;
        if (bind[i] >= m) nc++;
    }
    cnt_r[tid] = nr; cnt_c[tid] = nc;
    __syncthreads();
    if (tid == 0) {
        int sr = 0, sc = 0;
        for (int t = 0; t < nt; t++) {
            int a = cnt_r[t], b = cnt_c[t];
            cnt_r[t] = sr; cnt_c[t] = sc;
            sr += a; sc += b;
        }
        ctrl->k = sr;
        ctrl->sing = (sr != sc);
        ctrl->max_a = 0.0;
    }
    __syncthreads();
    int br = cnt_r[tid], bc = cnt_c[tid];
    for (int i = beg; i < end; i++) {
        if (head[i] >= m) { rslot[i] = br; slot_pos[br] = i; br++; } else rslot[i] = -1;
        if (bind[i] >= m) { cslot[i] = bc; slot_row[bc] = i; bc++; } else cslot[i] = -1;
    }
}

/* (the inversion itself is k_refactor, refactor.cuh) */

/* ------------------------------------------------------------------ */
/* per-iteration scalar logic and vector updates                      */
/* ------------------------------------------------------------------ */

/* scatter h = -N[q] (lib/glpspx01.js:702-719) into a zeroed dense vector */
__global__ void k_col_rhs(Ctrl *ctrl, int m, const int *__restrict__ a_ptr,
                          const int *__restrict__ a_ind, const double *__restrict__ a_val,
                          const int *__restrict__ head, double *__restrict__ h)
{
    if (ctrl->status != ST_OK) return;
    const int q = ctrl->q;
    if (q < 0) return;
    const int k = head[m + q];
    if (k < m) { if (blockIdx.x == 0 && threadIdx.x == 0) h[k] = -1.0; return; }
    for (int ptr = a_ptr[k - m] + blockIdx.x * blockDim.x + threadIdx.x; ptr < a_ptr[k - m + 1];
         ptr += gridDim.x * blockDim.x)
        h[a_ind[ptr]] = a_val[ptr];
}

/* primal, after eval_tcol: |tcol|_inf, the recomputed reduced cost d2
   (reeval_cost, lib/glpspx01.js:1133-1152), the d1/d2 test (:1901-1919), the
   PSE scalars gamma_q, delta_q and the masked vector v that update_gamma
   solves with (:1208-1218) */
__global__ void k_primal_prep(Ctrl *ctrl, int m, const int *__restrict__ head,
                              const double *__restrict__ coef, const double *__restrict__ tcol,
                              const signed char *__restrict__ refsp, double *__restrict__ cbar,
                              double *__restrict__ v, double tol_piv, Key *scratch)
{
    if (ctrl->status != ST_OK) return;
    const int q = ctrl->q;
    const int cbar_fresh = ctrl->cbar_fresh, rigorous = ctrl->rigorous, pse = gamma_on(ctrl);
    Key none = {0.0, 0.0, 0.0, 0, 0};
    Key acc = none;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        double t = tcol[i];
        int k = head[i];
        acc.a += coef[k] * t;
        acc.c = fmax(acc.c, fabs(t));
        if (pse) {
            double vv = refsp[k] ? t : 0.0;
            v[i] = vv;
            acc.b += vv * vv;
        }
    }
    grid_reduce(acc, none, scratch, &ctrl->ticket[2], CombSum2(), [=](const Key &r) {
        const int kq = head[m + q];
        double big = r.c;
        ctrl->tcol_max = big;
        ctrl->eps = tol_piv * (1.0 + 0.01 * big);
        double d1 = cbar[q], d2 = coef[kq] + r.a;
        ctrl->d2 = d2;
        double dq = (refsp[kq] ? 1.0 : 0.0);
        ctrl->delta_q = dq;
        ctrl->gamma_q = dq + r.b;
        if (fabs(d1 - d2) > 1e-5 * (1.0 + fabs(d2)) ||
            !((d1 < 0.0 && d2 < 0.0) || (d1 > 0.0 && d2 > 0.0))) {
            if (!cbar_fresh || !rigorous) { ctrl->status = ST_D1D2; return; }
        }
        if (d1 > 0.0) d1 = (d2 > 0.0 ? d2 : +DBL_EPSILON);
        else d1 = (d2 < 0.0 ? d2 : -DBL_EPSILON);
        cbar[q] = d1;
        ctrl->d1 = d1;
    });
}

/* primal, after eval_trow: pivot agreement test (lib/glpspx01.js:1985-2006)
   and the scalars of update_cbar (:1154-1176).  One thread. */
__global__ void k_primal_piv(Ctrl *ctrl, int m, const int *__restrict__ head,
                             double *__restrict__ trow, double *__restrict__ cbar,
                             double *__restrict__ coef, const double *__restrict__ tcol)
{
    if (ctrl->status != ST_OK) return;
    const int p = ctrl->p, q = ctrl->q;
    const int binv_fresh = ctrl->binv_fresh, rigorous = ctrl->rigorous;
    if (p < 0) return;
    double piv1 = tcol[p], piv2 = trow[q];
    ctrl->piv1 = piv1; ctrl->piv2 = piv2;
    if (fabs(piv1 - piv2) > 1e-8 * (1.0 + fabs(piv1)) ||
        !((piv1 > 0.0 && piv2 > 0.0) || (piv1 < 0.0 && piv2 < 0.0))) {
        if (!binv_fresh || !rigorous) { ctrl->status = ST_PIV12; return; }
        trow[q] = piv1;
        piv2 = piv1;
    }
    ctrl->new_dq = cbar[q] / piv2;
}

/* primal: update_bbar (:1100-1131), update_cbar (:1154-1176), the phase-1
   cost fix (:2016-2020) and update_gamma's vector part (:1223-1254) */
__global__ void k_primal_update(Ctrl *ctrl, int m, int n, const int *__restrict__ head,
                                const signed char *__restrict__ stat,
                                const signed char *__restrict__ type, const double *__restrict__ lb,
                                const double *__restrict__ ub, double *__restrict__ coef,
                                double *__restrict__ bbar, double *__restrict__ cbar,
                                double *__restrict__ gamma, const signed char *__restrict__ refsp,
                                const double *__restrict__ tcol, const double *__restrict__ trow,
                                const double *__restrict__ svec, double *__restrict__ hz)
{
    if (ctrl->status != ST_OK) return;
    const int p = ctrl->p, q = ctrl->q;
    const int do_gamma = gamma_on(ctrl);
    const double teta = ctrl->teta;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < m) {
        hz[t] = 0.0;     /* keep the dense right-hand side of eval_tcol all-zero */
        if (t == p) bbar[t] = get_xN(stat, head, lb, ub, m, q) + teta;
        else if (teta != 0.0) {
            double tc = tcol[t];
            if (tc != 0.0) bbar[t] += tc * teta;
        }
    }
    if (p < 0 || t >= n) return;
    const double new_dq = ctrl->new_dq;
    const double pivot = trow[q];
    const int kp = head[p];
    if (t == q) {
        double c = new_dq;
        if (ctrl->phase == 1) { c -= coef[kp]; coef[kp] = 0.0; }
        cbar[q] = c;
        if (do_gamma) {
            double g = 1.0;
            if (type[kp] != GLP_FX) {
                g = ctrl->gamma_q / (pivot * pivot);
                if (g < DBL_EPSILON) g = DBL_EPSILON;
            }
            gamma[q] = g;
        }
    } else {
        double tr = trow[t];
        if (tr != 0.0) {
            cbar[t] -= tr * new_dq;
            if (do_gamma) {
                double tt = tr / pivot;
                int k = head[m + t];
                double t1 = gamma[t] + tt * tt * ctrl->gamma_q + 2.0 * tt * svec[t];
                double t2 = (refsp[k] ? 1.0 : 0.0) + ctrl->delta_q * tt * tt;
                double g = (t1 >= t2 ? t1 : t2);
                if (g < DBL_EPSILON) g = DBL_EPSILON;
                gamma[t] = g;
            }
        }
    }
}

/* dual, after eval_trow: |trow|_inf -> eps of sort_trow (lib/glpspx02.js:754-791) */
__global__ void k_dual_rowmax(Ctrl *ctrl, double tol)
{
    if (ctrl->status != ST_OK) return;
    double big = *((volatile double *)&ctrl->big);
    ctrl->trow_max = big;
    ctrl->eps = tol * (1.0 + 0.01 * big);
    ctrl->big = 0.0;
}

/* dual, after eval_tcol: pivot agreement test (lib/glpspx02.js:1913-1933),
   objective tracking (:1936-1938), PSE scalars gamma_p, eta_p (:1103-1115) */
__global__ void k_dual_prep(Ctrl *ctrl, int m, int n, const int *__restrict__ head,
                            const signed char *__restrict__ refsp, const double *__restrict__ trow,
                            double *__restrict__ tcol, const double *__restrict__ cbar,
                            const signed char *__restrict__ stat, double zeta, Key *scratch)
{
    if (ctrl->status != ST_OK) return;
    const int p = ctrl->p, q = ctrl->q;
    const int binv_fresh = ctrl->binv_fresh, rigorous = ctrl->rigorous, pse = gamma_on(ctrl);
    Key none = {0.0, 0.0, 0.0, 0, 0};
    Key acc = none;
    if (pse) {
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
            double t = trow[j];
            if (t != 0.0 && refsp[head[m + j]]) acc.a += t * t;
        }
    }
    grid_reduce(acc, none, scratch, &ctrl->ticket[2], CombSum2(), [=](const Key &r) {
        double piv1 = tcol[p], piv2 = trow[q];
        ctrl->piv1 = piv1; ctrl->piv2 = piv2;
        if (fabs(piv1 - piv2) > 1e-8 * (1.0 + fabs(piv1)) ||
            !((piv1 > 0.0 && piv2 > 0.0) || (piv1 < 0.0 && piv2 < 0.0))) {
            if (!binv_fresh || !rigorous) { ctrl->status = ST_PIV12; return; }
            tcol[p] = piv2;
            piv1 = piv2;
        }
        double eta = (refsp[head[p]] ? 1.0 : 0.0);
        ctrl->delta_q = eta;              /* eta_p   */
        ctrl->gamma_q = eta + r.a;        /* gamma_p */
        ctrl->teta = ctrl->delta / piv1;
        if (ctrl->phase == 2) ctrl->obj += (cbar[q] / zeta) * (ctrl->delta / piv1);
    });
}

/* dual: update_bbar (:1042-1073), update_cbar (:1020-1040), update_gamma's
   vector part (:1137-1187) */
__global__ void k_dual_update(Ctrl *ctrl, int m, int n, const int *__restrict__ head,
                              const signed char *__restrict__ stat,
                              const signed char *__restrict__ type, const double *__restrict__ lb,
                              const double *__restrict__ ub, double *__restrict__ bbar,
                              double *__restrict__ cbar, double *__restrict__ gamma,
                              signed char *__restrict__ refsp, const double *__restrict__ tcol,
                              const double *__restrict__ trow, const double *__restrict__ u,
                              double *__restrict__ hz)
{
    if (ctrl->status != ST_OK) return;
    const int p = ctrl->p, q = ctrl->q;
    const int do_gamma = gamma_on(ctrl);
    const double teta = ctrl->teta, new_dq = ctrl->new_dq;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int kp = head[p], kq = head[m + q];
    const double pivot = tcol[p];
    if (t < n) {
        if (t == q) cbar[q] = new_dq;
        else if (new_dq != 0.0) {
            double tr = trow[t];
            if (tr != 0.0) cbar[t] -= tr * new_dq;
        }
    }
    if (t < m) {
        hz[t] = 0.0;
        const bool drop = (type[kp] == GLP_FX && refsp[kp]);
        if (t == p) {
            bbar[p] = get_xN(stat, head, lb, ub, m, q) + teta;
            if (do_gamma) {
                double g = 1.0;
                if (type[kq] != GLP_FR) {
                    g = ctrl->gamma_q / (pivot * pivot);
                    if (g < DBL_EPSILON) g = DBL_EPSILON;
                    if (drop) {
                        double tt = 1.0 / pivot;
                        g -= tt * tt;
                        if (g < DBL_EPSILON) g = DBL_EPSILON;
                    }
                }
                gamma[p] = g;
            }
        } else {
            double tc = tcol[t];
            if (tc != 0.0) {
                if (teta != 0.0) bbar[t] += tc * teta;
                int k = head[t];
                if (do_gamma && type[k] != GLP_FR) {
                    double tt = tc / pivot;
                    double t1 = gamma[t] + tt * tt * ctrl->gamma_q + 2.0 * tt * u[t];
                    double t2 = (refsp[k] ? 1.0 : 0.0) + ctrl->delta_q * tt * tt;
                    double g = (t1 >= t2 ? t1 : t2);
                    if (g < DBL_EPSILON) g = DBL_EPSILON;
                    if (drop) {
                        g -= tt * tt;
                        if (g < DBL_EPSILON) g = DBL_EPSILON;
                    }
                    gamma[t] = g;
                }
            }
        }
    }
}

/* second half of the fixed-variable rule of dual update_gamma (:1171-1173):
   runs after k_dual_update so that no thread still reads refsp[kp] */
__global__ void k_dual_drop_refsp(Ctrl *ctrl, const int *__restrict__ head,
                                  const signed char *__restrict__ type, signed char *__restrict__ refsp)
{
    if (ctrl->status != ST_OK) return;
    int kp = head[ctrl->p];
    if (type[kp] == GLP_FX && refsp[kp]) refsp[kp] = 0;
}

/* reset_refsp: lib/glpspx01.js:586-601 (primal, dual = 0) / lib/glpspx02.js:497-512 */
__global__ void k_reset_refsp(int m, int n, const int *__restrict__ head,
                              signed char *__restrict__ refsp, double *__restrict__ gamma, int dual)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < m + n) {
        int k = head[t];
        bool in = dual ? (t < m) : (t >= m);
        refsp[k] = in ? 1 : 0;
    }
    if (dual ? (t < m) : (t < n)) gamma[t] = 1.0;
}

/* set_aux_obj (lib/glpspx01.js:1373-1414): phase-1 costs; ctrl->cnt = violations */
__global__ void k_set_aux_obj(Ctrl *ctrl, int m, int n, const int *__restrict__ head,
                              const signed char *__restrict__ type, const double *__restrict__ lb,
                              const double *__restrict__ ub, const double *__restrict__ bbar,
                              double *__restrict__ coef, double tol_bnd)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m + n) return;
    const int k = head[t];
    double c = 0.0;
    if (t < m) {
        int tk = type[k];
        double b = bbar[t];
        if (tk == GLP_LO || tk == GLP_DB || tk == GLP_FX)
            if (b < lb[k] - relax(tol_bnd, lb[k])) c = -1.0;
        if (tk == GLP_UP || tk == GLP_DB || tk == GLP_FX)
            if (b > ub[k] + relax(tol_bnd, ub[k])) c = +1.0;
        if (c != 0.0) atomicAdd(&ctrl->cnt, 1);
    }
    coef[k] = c;
}

/* set_orig_obj (lib/glpspx01.js:1416-1427): coef = zeta * obj on structurals */
__global__ void k_set_orig_obj(int m, int n, const double *__restrict__ obj,
                               double *__restrict__ coef, double zeta)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m + n) return;
    coef[k] = (k < m) ? 0.0 : zeta * obj[1 + (k - m)];
}

/* primal check_stab (:1429-1481) and check_feas (:1483-1522): ctrl->flag |= 1 */
__global__ void k_primal_check(Ctrl *ctrl, int m, int mode, int phase, const int *__restrict__ head,
                               const signed char *__restrict__ type, const double *__restrict__ lb,
                               const double *__restrict__ ub, const double *__restrict__ coef,
                               const double *__restrict__ bbar, double tol_bnd)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int k = head[i];
    const double b = bbar[i], c = coef[k];
    const int tk = type[k];
    bool bad = false;
    if (mode == 0) { /* check_stab */
        if (phase == 1 && c < 0.0) bad = b > lb[k] + relax(tol_bnd, lb[k]);
        else if (phase == 1 && c > 0.0) bad = b < ub[k] - relax(tol_bnd, ub[k]);
        else {
            if (tk == GLP_LO || tk == GLP_DB || tk == GLP_FX) bad = b < lb[k] - relax(tol_bnd, lb[k]);
            if (!bad && (tk == GLP_UP || tk == GLP_DB || tk == GLP_FX))
                bad = b > ub[k] + relax(tol_bnd, ub[k]);
        }
    } else { /* check_feas (phase 1) */
        if (c < 0.0) bad = b < lb[k] - relax(tol_bnd, lb[k]);
        else if (c > 0.0) bad = b > ub[k] + relax(tol_bnd, ub[k]);
    }
    if (bad) ctrl->flag = 1;
}

/* dual check_feas (lib/glpspx02.js:1296-1315, mode 1) and check_stab
   (:1410-1422, mode 0) */
__global__ void k_dual_check(Ctrl *ctrl, int m, int n, int mode, const int *__restrict__ head,
                             const signed char *__restrict__ orig_type,
                             const signed char *__restrict__ stat, const double *__restrict__ cbar,
                             double tol_dj, int to_cnt = 0)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double d = cbar[j];
    bool bad = false;
    if (mode == 1) {
        int t = orig_type[head[m + j]];
        if (d < -tol_dj) bad = (t == GLP_LO || t == GLP_FR);
        if (d > +tol_dj) bad = bad || (t == GLP_UP || t == GLP_FR);
    } else {
        int s = stat[j];
        if (d < -tol_dj) bad = (s == GLP_NL || s == GLP_NF);
        if (d > +tol_dj) bad = bad || (s == GLP_NU || s == GLP_NF);
    }
    if (bad) { if (to_cnt) ctrl->cnt = 1; else ctrl->flag = 1; }
}

/* dual set_aux_bnds (:1317-1359, aux = 1) / set_orig_bnds (:1361-1408, aux = 0) */
__global__ void k_dual_set_bnds(const Ctrl *ctrl, int m, int n, int aux, const int *__restrict__ head,
                                const int *__restrict__ bind,
                                const signed char *__restrict__ orig_type,
                                const double *__restrict__ orig_lb, const double *__restrict__ orig_ub,
                                signed char *__restrict__ type, double *__restrict__ lb,
                                double *__restrict__ ub, signed char *__restrict__ stat,
                                const double *__restrict__ cbar)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m + n) return;
    if (aux < 0) aux = ctrl->flag;      /* phase decided on the device by the preceding k_dual_check(mode 1) */
    int t;
    double l, u;
    if (aux) {
        switch (orig_type[k]) {
        case GLP_FR: t = GLP_DB; l = -1e3; u = +1e3; break;
        case GLP_LO: t = GLP_DB; l = 0.0; u = +1.0; break;
        case GLP_UP: t = GLP_DB; l = -1.0; u = 0.0; break;
        default: t = GLP_FX; l = u = 0.0; break;
        }
    } else { t = orig_type[k]; l = orig_lb[k]; u = orig_ub[k]; }
    type[k] = (signed char)t; lb[k] = l; ub[k] = u;
    const int pos = bind[k];
    if (pos < m) return;
    const int j = pos - m;
    const double d = cbar[j];
    int s;
    if (aux) s = (t == GLP_FX) ? GLP_NS : (d >= 0.0 ? GLP_NL : GLP_NU);
    else {
        switch (t) {
        case GLP_FR: s = GLP_NF; break;
        case GLP_LO: s = GLP_NL; break;
        case GLP_UP: s = GLP_NU; break;
        case GLP_DB:
            if (d >= +DBL_EPSILON) s = GLP_NL;
            else if (d <= -DBL_EPSILON) s = GLP_NU;
            else s = (fabs(l) <= fabs(u)) ? GLP_NL : GLP_NU;
            break;
        default: s = GLP_NS;
        }
    }
    stat[j] = (signed char)s;
}

/* eval_obj (lib/glpspx01.js:1524-1548): two-sum reduction over positions */
__global__ void k_eval_obj(Ctrl *ctrl, int m, int n, const int *__restrict__ head,
                           const double *__restrict__ obj, const double *__restrict__ bbar,
                           const signed char *__restrict__ stat, const double *__restrict__ lb,
                           const double *__restrict__ ub, Key *scratch)
{
    Key none = {0.0, 0.0, 0.0, 0, 0};
    Key acc = none;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < m + n; t += gridDim.x * blockDim.x) {
        int k = head[t];
        if (k >= m) {
            double x = (t < m) ? bbar[t] : get_xN(stat, head, lb, ub, m, t - m);
            acc.a += obj[1 + (k - m)] * x;
        }
    }
    grid_reduce(acc, none, scratch, &ctrl->ticket[3], CombSum2(),
                [=](const Key &r) { ctrl->obj = obj[0] + r.a; });
}

__global__ void k_gather_cB(int m, const int *__restrict__ head, const double *__restrict__ coef,
                            double *__restrict__ cB)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) cB[i] = coef[head[i]];
}

__global__ void k_axpy1(int m, double *__restrict__ x, const double *__restrict__ d)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) x[i] += d[i];
}

__global__ void k_unit(Ctrl *ctrl, int m, double *__restrict__ e)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) e[i] = (i == ctrl->p) ? 1.0 : 0.0;
}

/* Batched glp_eval_tab_row (lib/glpapi12.js:401-453) for branch-and-bound: CTA f
   evaluates the simplex-table row of the basic variable at position pos[f]:
   rho_f = row pos[f] of inv(B) (the reference form of k_rho), then
   trow_f[j] = -rho_f' N_j for every non-basic non-fixed column (k_trow).
   One launch and one read-back for all fractional variables of a node instead
   of one round trip each. */
__global__ void k_tab_rows(const Ctrl *ctrl, int m, int n, const int *__restrict__ pos,
                           const double *__restrict__ T, int ldt,
                           const int *__restrict__ a_ptr, const int *__restrict__ a_ind,
                           const double *__restrict__ a_val, const int *__restrict__ at_ptr,
                           const int *__restrict__ at_ind, const double *__restrict__ at_val,
                           const int *__restrict__ head, const int *__restrict__ bind,
                           const signed char *__restrict__ stat, const int *__restrict__ rslot,
                           const int *__restrict__ cslot, double *__restrict__ rho_b,
                           double *__restrict__ trow_b)
{
    const int f = blockIdx.x;
    const int p = pos[f];
    const int kp = head[p];
    double *rho = rho_b + (size_t)f * m;
    double *trow = trow_b + (size_t)f * n;
    for (int r = threadIdx.x; r < m; r += blockDim.x) {
        const int cs = cslot[r];
        double v;
        if (cs < 0) v = (bind[r] == p) ? 1.0 : 0.0;
        else {
            const double *col = T + (size_t)cs * ldt;
            if (kp >= m) v = col[rslot[p]];
            else {
                v = 0.0;
                for (int ptr = at_ptr[kp]; ptr < at_ptr[kp + 1]; ptr++) {
                    const int pb = bind[m + at_ind[ptr]];
                    if (pb < m) v += at_val[ptr] * col[rslot[pb]];
                }
            }
        }
        rho[r] = v;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        double t = 0.0;
        if (stat[j] != GLP_NS) {
            const int k = head[m + j];
            if (k < m) t = -rho[k];
            else
                for (int ptr = a_ptr[k - m]; ptr < a_ptr[k - m + 1]; ptr++) t += rho[a_ind[ptr]] * a_val[ptr];
        }
        trow[j] = t;
    }
}

/* seeds the device-resident loop state before a batch of iterations */
/* The reference looks at the basic solution at the top of EVERY iteration of phase 1 and switches to phase 2 as
   soon as nothing violates its bound any more (primal check_feas, lib/glpspx01.js:1768-1775; dual check_feas,
   lib/glpspx02.js:1681-1696).  Inside a batch of enqueued iterations this kernel makes that test after each
   iteration and ends the batch (ST_PHASE); the host loop then repeats the test where the reference makes it and
   changes the phase.  One CTA. */
__global__ void k_phase1_stop(Ctrl *ctrl, int dual, int m, int n, const int *__restrict__ head,
                              const signed char *__restrict__ orig_type, const double *__restrict__ lb,
                              const double *__restrict__ ub, const double *__restrict__ coef,
                              const double *__restrict__ bbar, const double *__restrict__ cbar, double tol)
{
    if (ctrl->status != ST_OK || ctrl->phase != 1) return;
    int bad = 0;
    if (!dual) {
        for (int i = threadIdx.x; i < m && !bad; i += blockDim.x) {
            const int k = head[i];
            const double b = bbar[i], c = coef[k];
            if (c < 0.0) bad = b < lb[k] - relax(tol, lb[k]);
            else if (c > 0.0) bad = b > ub[k] + relax(tol, ub[k]);
        }
    } else {
        for (int j = threadIdx.x; j < n && !bad; j += blockDim.x) {
            const double d = cbar[j];
            const int t = orig_type[head[m + j]];
            if (d < -tol) bad = (t == GLP_LO || t == GLP_FR);
            if (d > +tol) bad = bad || (t == GLP_UP || t == GLP_FR);
        }
    }
    if (!__syncthreads_or(bad) && threadIdx.x == 0) ctrl->status = ST_PHASE;
}

__global__ void k_batch_begin(Ctrl *ctrl, int phase, int it_cnt, int it_max, int refct, int upd_cnt,
                              int period, int rigorous, int bbar_fresh, int cbar_fresh, int binv_fresh,
                              int pse, double obj_ll, double obj_ul, double zeta)
{
    ctrl->status = ST_OK; ctrl->flag = 0; ctrl->cnt = 0; ctrl->big = 0.0; ctrl->skip2 = 0;
    ctrl->phase = phase; ctrl->it_cnt = it_cnt; ctrl->it_max = it_max; ctrl->refct = refct;
    ctrl->upd_cnt = upd_cnt; ctrl->period = period; ctrl->n_done = 0; ctrl->rigorous = rigorous;
    ctrl->bbar_fresh = bbar_fresh; ctrl->cbar_fresh = cbar_fresh; ctrl->binv_fresh = binv_fresh;
    ctrl->pse = pse; ctrl->obj_ll = obj_ll; ctrl->obj_ul = obj_ul; ctrl->zeta = zeta;
}

__global__ void k_clear_ctrl(Ctrl *ctrl, int phase)
{
    ctrl->status = ST_OK; ctrl->flag = 0; ctrl->cnt = 0; ctrl->big = 0.0; ctrl->skip2 = 0;
    ctrl->phase = phase;
}

#endif /* GLPB_KERNELS_CUH */
