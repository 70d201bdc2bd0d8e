/* engine.cuh -- persistent iteration engine of libglpb200 (sm_100a).
 *
 * One cooperative launch runs MANY simplex iterations: the grid (one CTA of
 * 1024 threads per SM; a single CTA for problems that fit one SM) walks the
 * phases of spx_primal's / spx_dual's main loop (lib/glpspx01.js:1868-2056,
 * lib/glpspx02.js:1780-1966) separated by grid-wide barriers instead of kernel
 * boundaries.  8 barriers per iteration replace ~18 kernel launches.
 *
 *   - The scalar loop state (Ctrl) is REPLICATED: every CTA keeps a copy in
 *     shared memory and applies the same deterministic updates after each
 *     barrier from the same per-CTA partial results, so no broadcast round is
 *     needed; CTA 0 writes it back when the engine stops.
 *   - Every arg-reduction is two-stage (CTA partial -> scratch slot -> barrier
 *     -> every CTA combines the G partials in index order): deterministic,
 *     ties to the lowest index exactly as the sequential loops of the reference.
 *   - The engine runs while the reference's loop would simply `continue` to
 *     the next iteration.  The first exceptional branch (nothing to price, no
 *     ratio, small pivot, d1/d2 or piv1/piv2 disagreement, refactorisation due,
 *     limits, reference-space reset, objective cut-off) is recorded in
 *     Ctrl.status and the engine stops; the host loop controller handles it
 *     exactly where the reference's loop branches and starts the engine again.
 *   - Iterations in "rigorous" mode (iterative refinement of tcol/rho) are rare
 *     and run through the per-kernel path of solver.cu.
 *
 * Data that other CTAs modify during the launch is never read through
 * __restrict__/__ldg (non-coherent) paths; the constraint matrix, which is
 * immutable, is.
 */
#ifndef GLPB_ENGINE_CUH
#define GLPB_ENGINE_CUH
#include "kernels.cuh"

#define ENG_NT 1024          /* threads per CTA                                  */
#define ENG_MAXG 160         /* scratch slot width (>= number of SMs)            */
#define ENG_LCAP 2048        /* staged entries of one sparse column / row        */
#define ENG_NSLOT 10

struct EngArgs {
    Ctrl *ctrl;
    int m, n, ldt, max_iters;
    int gc, gr;               /* lanes per column / row of A                     */
    int dcap;                 /* doubles of dynamic shared memory for staging    */
    double tol_bnd, tol_dj, tol_piv, rtol;
    const int *a_ptr, *a_ind; const double *a_val;
    const int *at_ptr, *at_ind; const double *at_val;
    signed char *type, *stat, *refsp;
    double *lb, *ub, *coef;
    int *head, *bind;
    double *bbar, *cbar, *gamma, *tcol, *trow, *rho, *svec;
    double *hz;               /* dense rhs of eval_tcol, all-zero between iterations */
    double *v, *u;            /* PSE work vectors [m]                            */
    double *yk, *yk2, *wk, *zn; /* kernel-space work [ldt]                       */
    double *T;
    int *rslot, *slot_pos, *cslot, *slot_row;
    Key *scratch;
    unsigned int *bar;
    long long *prof_cyc;      /* optional: SM cycles per phase, CTA 0 (NULL = off)  */
    double *prof_bytes;       /* optional: algorithmic bytes per phase              */
};

struct EngCtx {
    int G, cta, tid, lane, warp, gtid, gsize, gwarp, nwarp;
    unsigned int epoch;
    long long t_last;
    double *sh_d;             /* [dcap] staging values / dense vector            */
    int *sh_i;                /* [ENG_LCAP] staging indices                      */
};

/* grid-wide barrier: monotone arrival counter, zeroed by the host before the launch */
__device__ __forceinline__ void eng_bar(EngCtx &X, const EngArgs &A)
{
    __syncthreads();
    if (X.G > 1) {
        X.epoch += (unsigned int)X.G;
        if (X.tid == 0) {
            __threadfence();
            atomicAdd(A.bar, 1u);
            while (*((volatile unsigned int *)A.bar) < X.epoch) { }
            __threadfence();
        }
        __syncthreads();
    }
}

/* phase accounting for bench.py's roofline leg: CTA 0 leaves every barrier
   last-or-together with the grid, so its cycle stamps bound the phase */
__device__ __forceinline__ void eng_mark(EngCtx &X, const EngArgs &A, int phase, double bytes)
{
    if (A.prof_cyc != nullptr && X.cta == 0 && X.tid == 0) {
        const long long t = clock64();
        A.prof_cyc[phase] += t - X.t_last;
        A.prof_bytes[phase] += bytes;
        X.t_last = t;
    }
}

/* CTA partial -> slot -> barrier -> every CTA combines the G partials */
template <class Comb>
__device__ Key eng_allreduce(EngCtx &X, const EngArgs &A, Key v, const Key &none, int slot, Comb comb)
{
    __shared__ Key sh_res;
    v = block_reduce(v, none, comb);
    Key *sl = A.scratch + slot * ENG_MAXG;
    if (X.G > 1) {
        if (X.tid == 0) sl[X.cta] = v;
        eng_bar(X, A);
        if (X.tid < 32) {
            Key r = none;
            for (int i = X.tid; i < X.G; i += 32) {
                Key o;
                o.a = __ldcg(&sl[i].a); o.b = __ldcg(&sl[i].b); o.c = __ldcg(&sl[i].c);
                o.pos = __ldcg(&sl[i].pos); o.aux = __ldcg(&sl[i].aux);
                comb(r, o);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                Key o = key_shfl_down(r, off);
                comb(r, o);
            }
            if (X.tid == 0) sh_res = r;
        }
    } else {
        if (X.tid == 0) sh_res = v;
    }
    __syncthreads();
    Key r = sh_res;
    __syncthreads();
    return r;
}

/* ordered compaction inside a CTA: returns the position of this thread's
   element among the flagged ones and the total; all threads must call */
__device__ __forceinline__ int eng_compact(const EngCtx &X, bool flag, int &total)
{
    __shared__ int cnt[33];
    const unsigned int b = __ballot_sync(FULLMASK, flag);
    if (X.lane == 0) cnt[X.warp] = __popc(b);
    __syncthreads();
    if (X.tid < 32) {
        int c = cnt[X.tid], s = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int o = __shfl_up_sync(FULLMASK, s, off);
            if (X.lane >= off) s += o;
        }
        cnt[X.tid] = s - c;
        if (X.tid == 31) cnt[32] = s;
    }
    __syncthreads();
    const int pos = cnt[X.warp] + __popc(b & ((1u << X.lane) - 1u));
    total = cnt[32];
    __syncthreads();
    return pos;
}

/* y[b] (+)= sum_e T[b, idx[e]] * val[e], b < k, for the list (idx, val) of L
   entries (idx == NULL: idx[e] = e, the dense case).  A CTA owns blocks of RB
   consecutive rows; its 32 warps split the list, partial sums meet in shared
   memory in a fixed order.  Streams 8 L k bytes of T. */
__device__ void eng_gemv_rows(const EngCtx &X, const EngArgs &A, int k, int L, const int *idx,
                              const double *val, double *y, bool accumulate)
{
    __shared__ double red[32][33];
    int RB = 8;
    if (k >= 32 * 2 * X.G) RB = 32; else if (k >= 16 * 2 * X.G) RB = 16;
    const int NSUB = 32 / RB;
    const int r = X.lane & (RB - 1), sub = X.lane / RB;
    const int nrb = (k + RB - 1) / RB;
    const int stride = NSUB * 32;
    const size_t ldt = (size_t)A.ldt;
    for (int rb = X.cta; rb < nrb; rb += X.G) {
        const int b = rb * RB + r;
        const bool inb = b < k;
        const double *Tb = A.T + (inb ? b : 0);
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
        int e = sub + NSUB * X.warp;
        if (idx) {
            for (; e + 3 * stride < L; e += 4 * stride) {
                int c0 = idx[e], c1 = idx[e + stride], c2 = idx[e + 2 * stride], c3 = idx[e + 3 * stride];
                double t0 = __ldcg(Tb + c0 * ldt), t1 = __ldcg(Tb + c1 * ldt);
                double t2 = __ldcg(Tb + c2 * ldt), t3 = __ldcg(Tb + c3 * ldt);
                acc0 += t0 * val[e]; acc1 += t1 * val[e + stride];
                acc2 += t2 * val[e + 2 * stride]; acc3 += t3 * val[e + 3 * stride];
            }
            for (; e < L; e += stride) acc0 += __ldcg(Tb + idx[e] * ldt) * val[e];
        } else {
            for (; e + 7 * stride < L; e += 8 * stride) {
                double t[8];
#pragma unroll
                for (int x = 0; x < 8; x++) t[x] = __ldcg(Tb + (size_t)(e + x * stride) * ldt);
                acc0 += t[0] * val[e]; acc1 += t[1] * val[e + stride];
                acc2 += t[2] * val[e + 2 * stride]; acc3 += t[3] * val[e + 3 * stride];
                acc0 += t[4] * val[e + 4 * stride]; acc1 += t[5] * val[e + 5 * stride];
                acc2 += t[6] * val[e + 6 * stride]; acc3 += t[7] * val[e + 7 * stride];
            }
            for (; e < L; e += stride) acc0 += __ldcg(Tb + (size_t)e * ldt) * val[e];
        }
        double acc = (acc0 + acc1) + (acc2 + acc3);
        if (!inb) acc = 0.0;
        for (int off = RB; off < 32; off <<= 1) acc += __shfl_xor_sync(FULLMASK, acc, off);
        if (sub == 0) red[X.warp][r] = acc;
        __syncthreads();
        if (X.tid < RB) {
            double s = 0.0;
#pragma unroll 8
            for (int w = 0; w < 32; w++) s += red[w][X.tid];
            const int b2 = rb * RB + X.tid;
            if (b2 < k) y[b2] = accumulate ? y[b2] + s : s;
        }
        __syncthreads();
    }
}

/* FTRAN, first half, for the right-hand side h = -N_q (eval_tcol,
   lib/glpspx01.js:690-727): y = T h_N over the entries of column q that fall
   on rows of R_N.  CTA 0 also scatters h into the dense vector hz that the
   second half reads. */
__device__ void eng_ftran_head_col(EngCtx &X, const EngArgs &A, int k, int kq, double *y)
{
    const int m = A.m;
    const int RBmin = 8;
    const bool work = (X.cta * RBmin < k);
    if (!work && X.cta != 0) return;
    if (kq < m) {
        if (X.tid == 0) {
            X.sh_i[0] = A.cslot[kq]; X.sh_d[0] = -1.0;
            if (X.cta == 0) A.hz[kq] = -1.0;
        }
        __syncthreads();
        eng_gemv_rows(X, A, k, 1, X.sh_i, X.sh_d, y, false);
        return;
    }
    const int beg = __ldg(A.a_ptr + (kq - m)), end = __ldg(A.a_ptr + (kq - m) + 1);
    bool first = true;
    for (int seg = beg; seg < end || first; seg += ENG_LCAP) {
        const int segend = min(end, seg + ENG_LCAP);
        int L = 0;
        for (int base = seg; base < segend; base += ENG_NT) {
            const int e = base + X.tid;
            const bool valid = e < segend;
            int r = 0, cs = -1;
            double a = 0.0;
            if (valid) {
                r = __ldg(A.a_ind + e); a = __ldg(A.a_val + e); cs = A.cslot[r];
                if (X.cta == 0) A.hz[r] = a;
            }
            int tot;
            const int pos = eng_compact(X, valid && cs >= 0, tot);
            if (valid && cs >= 0) { X.sh_i[L + pos] = cs; X.sh_d[L + pos] = a; }
            L += tot;
        }
        __syncthreads();
        eng_gemv_rows(X, A, k, L, X.sh_i, X.sh_d, y, !first);
        first = false;
    }
}

/* FTRAN, second half: x[i] for every basic position from y = T h_N.  One
   group of GL lanes per position gathers the row of A of a basic auxiliary
   variable.  PREP (primal): the reductions of k_primal_prep ride along. */
template <int GL, bool PREP>
__device__ __forceinline__ void eng_ftran_tail(const EngCtx &X, const EngArgs &A, const Ctrl &S,
                                               const double *h, const double *y, double *x, Key &acc)
{
    const int m = A.m;
    const int ngroups = X.gsize / GL, g = X.gtid / GL, lane = X.tid % GL;
    const bool pse = PREP && gamma_on(&S);
    for (int i0 = 0; i0 < m; i0 += ngroups) {
        const int i = i0 + g;
        double a = 0.0;
        int kk = m;
        if (i < m) {
            kk = A.head[i];
            if (kk < m) {
                const int beg = __ldg(A.at_ptr + kk), end = __ldg(A.at_ptr + kk + 1);
                for (int ptr = beg + lane; ptr < end; ptr += GL) {
                    const int pb = A.bind[m + __ldg(A.at_ind + ptr)];
                    if (pb < m) a += __ldg(A.at_val + ptr) * y[A.rslot[pb]];
                }
            }
        }
        a = group_sum<GL>(a);
        if (i < m && lane == 0) {
            const double t = (kk < m) ? h[kk] + a : y[A.rslot[i]];
            x[i] = t;
            if (PREP) {
                acc.a += A.coef[kk] * t;
                acc.c = fmax(acc.c, fabs(t));
                if (pse) {
                    const double vv = A.refsp[kk] ? t : 0.0;
                    A.v[i] = vv;
                    acc.b += vv * vv;
                }
            }
        }
    }
}

/* BTRAN, first half: w[b] = c[pos_b] + sum_{r in R_B} A[r, j_b] c[bind[r]] */
template <int GL>
__device__ __forceinline__ void eng_btran_head(const EngCtx &X, const EngArgs &A, int k, const double *c, double *w)
{
    const int m = A.m;
    const int ngroups = X.gsize / GL, g = X.gtid / GL, lane = X.tid % GL;
    for (int b0 = 0; b0 < k; b0 += ngroups) {
        const int b = b0 + g;
        double a = 0.0;
        int i = 0;
        if (b < k) {
            i = A.slot_pos[b];
            const int j = A.head[i] - m;
            const int beg = __ldg(A.a_ptr + j), end = __ldg(A.a_ptr + j + 1);
            for (int ptr = beg + lane; ptr < end; ptr += GL) {
                const int pr = A.bind[__ldg(A.a_ind + ptr)];
                if (pr < m) a += __ldg(A.a_val + ptr) * c[pr];
            }
        }
        a = group_sum<GL>(a);
        if (b < k && lane == 0) w[b] = c[i] + a;
    }
}

/* zn[cs] = sum_b T[b, cs] w[b]: one warp per column, coalesced.  8 k^2 bytes. */
__device__ __forceinline__ void eng_gemvT(const EngCtx &X, const EngArgs &A, int k, const double *w, double *zn)
{
    for (int cs = X.gwarp; cs < k; cs += X.nwarp) {
        const double *col = A.T + (size_t)cs * A.ldt;
        double a0 = 0.0, a1 = 0.0;
        int b = X.lane;
        for (; b + 32 < k; b += 64) {
            double t0 = __ldcg(col + b), t1 = __ldcg(col + b + 32);
            a0 += t0 * w[b]; a1 += t1 * w[b + 32];
        }
        if (b < k) a0 += __ldcg(col + b) * w[b];
        double a = a0 + a1;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(FULLMASK, a, off);
        if (X.lane == 0) zn[cs] = a;
    }
}

/* rho = row p of inv(B) (eval_rho, lib/glpspx01.js:1030-1042), read out of T */
__device__ void eng_rho(EngCtx &X, const EngArgs &A, int k, int p)
{
    const int m = A.m;
    const int kp = A.head[p];
    for (int r = X.gtid; r < m; r += X.gsize)
        if (A.cslot[r] < 0) A.rho[r] = (A.bind[r] == p) ? 1.0 : 0.0;
    if (kp >= m) {
        const double *row = A.T + A.rslot[p];
        for (int cs = X.gtid; cs < k; cs += X.gsize) A.rho[A.slot_row[cs]] = __ldcg(row + (size_t)cs * A.ldt);
        return;
    }
    if (X.cta * (ENG_NT / 32) >= k) return;          /* no column of T for this CTA */
    const int beg = __ldg(A.at_ptr + kp), end = __ldg(A.at_ptr + kp + 1);
    bool first = true;
    for (int seg = beg; seg < end || first; seg += ENG_LCAP) {
        const int segend = min(end, seg + ENG_LCAP);
        int L = 0;
        for (int base = seg; base < segend; base += ENG_NT) {
            const int e = base + X.tid;
            const bool valid = e < segend;
            int pb = m;
            double a = 0.0;
            if (valid) { pb = A.bind[m + __ldg(A.at_ind + e)]; a = __ldg(A.at_val + e); }
            int tot;
            const int pos = eng_compact(X, valid && pb < m, tot);
            if (valid && pb < m) { X.sh_i[L + pos] = A.rslot[pb]; X.sh_d[L + pos] = a; }
            L += tot;
        }
        __syncthreads();
        for (int cs = X.gwarp; cs < k; cs += X.nwarp) {
            const double *col = A.T + (size_t)cs * A.ldt;
            double a = 0.0;
            for (int e = X.lane; e < L; e += 32) a += __ldcg(col + X.sh_i[e]) * X.sh_d[e];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(FULLMASK, a, off);
            if (X.lane == 0) {
                const int r = A.slot_row[cs];
                A.rho[r] = first ? a : A.rho[r] + a;
            }
        }
        first = false;
        __syncthreads();
    }
}

/* pivot row (eval_trow): trow[j] = -rho' N_j for the non-basic non-fixed
   columns; with PSE also s_j = N_j' u (primal update_gamma) from the same pass.
   acc.c collects |trow|_inf and acc.a the sum of trow_j^2 over the reference
   space (dual). */
template <int GL, bool DUAL>
__device__ __forceinline__ void eng_trow(const EngCtx &X, const EngArgs &A, const Ctrl &S, Key &acc)
{
    const int m = A.m, n = A.n;
    const int ngroups = X.gsize / GL, g = X.gtid / GL, lane = X.tid % GL;
    const bool pse = gamma_on(&S);
    const bool want_s = !DUAL && pse;
    for (int j0 = 0; j0 < n; j0 += ngroups) {
        const int j = j0 + g;
        double t = 0.0, s = 0.0;
        int k = 0;
        const bool live = (j < n) && A.stat[j] != GLP_NS;
        if (live) {
            k = A.head[m + j];
            if (k < m) {
                if (lane == 0) { t = -A.rho[k]; if (want_s) s = A.u[k]; }
            } else {
                const int beg = __ldg(A.a_ptr + (k - m)), end = __ldg(A.a_ptr + (k - m) + 1);
                for (int ptr = beg + lane; ptr < end; ptr += GL) {
                    const int r = __ldg(A.a_ind + ptr);
                    const double a = __ldg(A.a_val + ptr);
                    t += A.rho[r] * a;
                    if (want_s) s -= a * A.u[r];
                }
            }
        }
        t = group_sum<GL>(t);
        if (want_s) s = group_sum<GL>(s);
        if (j < n && lane == 0) {
            A.trow[j] = t;
            if (want_s) A.svec[j] = s;
            if (DUAL) {
                acc.c = fmax(acc.c, fabs(t));
                if (pse && live && t != 0.0 && A.refsp[k]) acc.a += t * t;
            }
        }
    }
}

/* dual update_gamma, first half (lib/glpspx02.js:1103-1132), by rows:
   v[r] = sum_{j in C, non-basic} N_j[r] trow_j; the entries on rows of R_N are
   also written in kernel order (wk) for the dense product with T */
template <int GL>
__device__ __forceinline__ void eng_gamma_rhs(const EngCtx &X, const EngArgs &A)
{
    const int m = A.m;
    const int ngroups = X.gsize / GL, g = X.gtid / GL, lane = X.tid % GL;
    for (int r0 = 0; r0 < m; r0 += ngroups) {
        const int row = r0 + g;
        double a = 0.0;
        if (row < m) {
            const int beg = __ldg(A.at_ptr + row), end = __ldg(A.at_ptr + row + 1);
            for (int ptr = beg + lane; ptr < end; ptr += GL) {
                const int jj = __ldg(A.at_ind + ptr);
                const int pb = A.bind[m + jj];
                if (pb >= m && A.refsp[m + jj]) a -= A.trow[pb - m] * __ldg(A.at_val + ptr);
            }
        }
        a = group_sum<GL>(a);
        if (row < m && lane == 0) {
            const int pr = A.bind[row];
            const double val = a + ((pr >= m && A.refsp[row]) ? A.trow[pr - m] : 0.0);
            A.v[row] = val;
            const int cs = A.cslot[row];
            if (cs >= 0) A.wk[cs] = val;
        }
    }
}

/* description of a basis change, identical in every CTA */
struct EngChange {
    int p, q, kp, kq, LS, ES, csq, bp, k, knew, ctgt, bnew;
    double tp;
};

__device__ __forceinline__ void eng_describe_change(const EngArgs &A, const Ctrl &S, EngChange &C)
{
    const int m = A.m;
    C.p = S.p; C.q = S.q; C.k = S.k;
    C.kp = A.head[C.p]; C.kq = A.head[m + C.q];
    C.LS = (C.kp < m); C.ES = (C.kq < m);
    C.tp = A.tcol[C.p];
    C.csq = C.ES ? A.cslot[C.kq] : -1;
    C.bp = C.LS ? -1 : A.rslot[C.p];
    C.ctgt = C.LS ? (C.ES ? C.csq : C.k) : -1;
    C.bnew = (C.LS && !C.ES) ? C.k : -1;
    C.knew = C.k;
    if (C.LS && !C.ES) C.knew = C.k + 1;
    if (!C.LS && C.ES) C.knew = C.k - 1;
}

/* The basis change on T in ONE pass (k_update_rank1 + the matrix part of
   k_update_fix): inv(B)' = E inv(B) restricted to the structural kernel.
   Every destination cell (b, cs) of the new kernel is computed from a source
   cell of the old one; when a row/column leaves, the last row/column moves
   into the hole by reading from it (the last row/column is never written).
   Algorithmic bytes: 16 k^2. */
#define ENG_UR 128
#define ENG_UC 32
__device__ void eng_update_T(const EngCtx &X, const EngArgs &A, const EngChange &C)
{
    const int k = C.k;
    const size_t ldt = (size_t)A.ldt;
    const bool removal = (!C.LS && C.ES);
    const int kd = removal ? k - 1 : k;                 /* destination range */
    const double tp = C.tp;
    const int r = X.tid & (ENG_UR - 1), cg = X.tid / ENG_UR;      /* 128 rows x 8 column groups */
    const int ntr = (kd + ENG_UR - 1) / ENG_UR, ntc = (kd + ENG_UC - 1) / ENG_UC;
    for (int tile = X.cta; tile < ntr * ntc; tile += X.G) {
        const int b = (tile % ntr) * ENG_UR + r;
        const int c0 = (tile / ntr) * ENG_UC;
        if (b >= kd) continue;
        const int sb = (removal && b == C.bp) ? k - 1 : b;
        const int i = A.slot_pos[sb];
        const bool isp = (i == C.p);
        const double f = isp ? -1.0 / tp : A.tcol[i] / tp;
#pragma unroll
        for (int x = 0; x < ENG_UC / 8; x++) {
            const int cs = c0 + cg + 8 * x;
            if (cs >= kd) continue;
            double *dst = A.T + (size_t)cs * ldt + b;
            if (C.LS && C.ES && cs == C.csq) { *dst = -A.tcol[i] / tp; continue; }
            const int scs = (removal && cs == C.csq) ? k - 1 : cs;
            const double rs = A.rho[A.slot_row[scs]];
            if (isp) { *dst = rs * f; continue; }
            const bool moved = (sb != b) || (scs != cs);
            if (f != 0.0 || moved) {
                const double old = __ldcg(A.T + (size_t)scs * ldt + sb);
                *dst = old - f * rs;
            }
        }
    }
    if (C.bnew >= 0) {
        /* a row and a column join: column k (rows 0..k) and row k (columns 0..k-1) */
        for (int t = X.gtid; t < 2 * k + 1; t += X.gsize) {
            if (t < k) A.T[(size_t)k * ldt + t] = -A.tcol[A.slot_pos[t]] / tp;
            else if (t < 2 * k) A.T[(size_t)(t - k) * ldt + k] = -A.rho[A.slot_row[t - k]] / tp;
            else A.T[(size_t)k * ldt + k] = -1.0 / tp;
        }
    }
}

/* slot maps, basis header and the new non-basic status (change_basis,
   lib/glpspx01.js:1310-1371 / lib/glpspx02.js:1259-1294); one thread */
__device__ void eng_bookkeep(const EngArgs &A, const EngChange &C, int new_stat, bool drop_refsp)
{
    const int m = A.m, k = C.k;
    if (C.LS) { A.cslot[C.kp] = C.ctgt; A.slot_row[C.ctgt] = C.kp; }
    if (C.ES) {
        A.cslot[C.kq] = -1;
        if (!C.LS && C.csq != k - 1) { int rl = A.slot_row[k - 1]; A.slot_row[C.csq] = rl; A.cslot[rl] = C.csq; }
    }
    if (C.bnew >= 0) { A.rslot[C.p] = C.bnew; A.slot_pos[C.bnew] = C.p; }
    if (!C.LS && C.ES) {
        A.rslot[C.p] = -1;
        if (C.bp != k - 1) { int pl = A.slot_pos[k - 1]; A.slot_pos[C.bp] = pl; A.rslot[pl] = C.bp; }
    }
    A.head[C.p] = C.kq; A.head[m + C.q] = C.kp;
    A.bind[C.kq] = C.p; A.bind[C.kp] = m + C.q;
    A.stat[C.q] = (signed char)new_stat;
    if (drop_refsp) A.refsp[C.kp] = 0;
}

#define ENG_GROUPS(G, CALL)                                                    \
    do {                                                                       \
        if ((G) == 32) { constexpr int GG = 32; CALL; }                        \
        else if ((G) == 8) { constexpr int GG = 8; CALL; }                     \
        else { constexpr int GG = 4; CALL; }                                   \
    } while (0)

__device__ __forceinline__ void eng_init(EngCtx &X, const EngArgs &A, double *dyn)
{
    X.G = gridDim.x; X.cta = blockIdx.x; X.tid = threadIdx.x;
    X.lane = X.tid & 31; X.warp = X.tid >> 5;
    X.gtid = X.cta * ENG_NT + X.tid; X.gsize = X.G * ENG_NT;
    X.gwarp = X.gtid >> 5; X.nwarp = X.gsize >> 5;
    X.epoch = 0u;
    X.t_last = clock64();
    X.sh_d = dyn;
    X.sh_i = (int *)(dyn + A.dcap);
}

/* ------------------------------------------------------------------ */
/* primal engine: lib/glpspx01.js:1868-2056                           */
/* ------------------------------------------------------------------ */
__global__ void __launch_bounds__(ENG_NT, 1) k_engine_primal(EngArgs A)
{
    extern __shared__ double eng_dyn[];
    __shared__ Ctrl S;
    __shared__ EngChange C;
    EngCtx X;
    eng_init(X, A, eng_dyn);
    const int m = A.m, n = A.n;
    if (X.tid == 0) S = *A.ctrl;
    __syncthreads();
    int qprev = -1, stprev = 0;
    for (int it = 0; it < A.max_iters && S.status == ST_OK; it++) {
        /* ---- P1: chuzc ---- */
        {
            Key none = {0.0, 0.0, 0.0, INT_MAX, 0};
            Key v = none;
            scan_chuzc_primal(v, X.gtid, X.gsize, n, A.stat, A.cbar, A.gamma, A.tol_dj, qprev, stprev);
            Key r = eng_allreduce(X, A, v, none, 0, CombArgMax());
            if (X.tid == 0) {
                S.q = (r.a > 0.0 && r.pos != INT_MAX) ? r.pos : P_NONE;
                S.big = 0.0;
                if (S.q == P_NONE) S.status = ST_NONE1;
            }
            __syncthreads();
            eng_mark(X, A, 0, 17.0 * n);
            if (S.status != ST_OK) break;
        }
        const int q = S.q;
        const int kq = A.head[m + q];
        const double nnz_q = (kq < m) ? 1.0 : (double)(__ldg(A.a_ptr + (kq - m) + 1) - __ldg(A.a_ptr + (kq - m)));
        /* ---- P2: tcol, first half ---- */
        eng_ftran_head_col(X, A, S.k, kq, A.yk);
        eng_bar(X, A);
        eng_mark(X, A, 1, 12.0 * nnz_q + 8.0 * S.k * nnz_q * ((double)S.k / m) + 8.0 * S.k);
        /* ---- P3: tcol, second half + the reductions of k_primal_prep ---- */
        {
            Key none = {0.0, 0.0, 0.0, 0, 0};
            Key acc = none;
            ENG_GROUPS(A.gr, (eng_ftran_tail<GG, true>(X, A, S, A.hz, A.yk, A.tcol, acc)));
            Key r = eng_allreduce(X, A, acc, none, 1, CombSum2());
            if (X.tid == 0) {
                const double big = r.c;
                S.tcol_max = big;
                S.eps = A.tol_piv * (1.0 + 0.01 * big);
                double d1 = A.cbar[q];
                const double d2 = A.coef[kq] + r.a;
                S.d2 = d2;
                const double dq = (A.refsp[kq] ? 1.0 : 0.0);
                S.delta_q = dq;
                S.gamma_q = dq + r.b;
                if (fabs(d1 - d2) > 1e-5 * (1.0 + fabs(d2)) ||
                    !((d1 < 0.0 && d2 < 0.0) || (d1 > 0.0 && d2 > 0.0)))
                    S.status = ST_D1D2;      /* the engine never runs in rigorous mode */
                else {
                    if (d1 > 0.0) d1 = (d2 > 0.0 ? d2 : +DBL_EPSILON);
                    else d1 = (d2 < 0.0 ? d2 : -DBL_EPSILON);
                    S.d1 = d1;
                }
            }
            __syncthreads();
            eng_mark(X, A, 2, 12.0 * (double)__ldg(A.at_ptr + m) * (1.0 - (double)S.k / m) + 29.0 * m);
            if (S.status != ST_OK) break;
        }
        const double sgn = (S.d1 > 0.0 ? -1.0 : +1.0);
        const bool pse = gamma_on(&S);
        /* ---- P4: Harris pass 1 (+ first half of u = inv(B') v) ---- */
        {
            Key none = {DBL_MAX, 0.0, 0.0, INT_MAX, 0};
            Key v = none;
            if (X.gtid == 0 && A.type[kq] == GLP_DB) {
                v.a = __dsub_rn(A.ub[kq], A.lb[kq]); v.b = 1.0; v.pos = -1; v.aux = 0;
            }
            scan_ratio_primal(v, X.gtid, X.gsize, 1, S.phase, sgn, S.eps, 0.0, A.rtol, A.type, A.lb, A.ub,
                              A.coef, A.head, A.bbar, A.tcol, nullptr, m);
            if (pse) ENG_GROUPS(A.gc, (eng_btran_head<GG>(X, A, S.k, A.v, A.wk)));
            Key r = eng_allreduce(X, A, v, none, 2, CombRatio1());
            /* reeval_cost result (lib/glpspx01.js:1915-1918); every CTA has read the old value by now */
            if (X.cta == 0 && X.tid == 0) A.cbar[q] = S.d1;
            if (X.tid == 0) fin_ratio_primal(&S, r, 1, sgn, A.rtol, A.type, A.head, A.tcol, nullptr);
            __syncthreads();
            eng_mark(X, A, 3, 45.0 * m + (pse ? 12.0 * (double)__ldg(A.a_ptr + n) * ((double)S.k / n) + 16.0 * S.k : 0.0));
            if (S.status != ST_OK) break;
        }
        /* ---- P5: Harris pass 2 (+ second half of u) ---- */
        {
            const bool pass2 = !S.skip2;
            Key none = {DBL_MAX, 0.0, 0.0, INT_MAX, 0};
            Key v = none;
            if (pass2)
                scan_ratio_primal(v, X.gtid, X.gsize, 2, S.phase, sgn, S.eps, S.tmax, A.rtol, A.type, A.lb, A.ub,
                                  A.coef, A.head, A.bbar, A.tcol, nullptr, m);
            if (pse && S.p != P_FLIP) eng_gemvT(X, A, S.k, A.wk, A.zn);
            if (pass2) {
                Key r = eng_allreduce(X, A, v, none, 3, CombRatio2());
                if (X.tid == 0) fin_ratio_primal(&S, r, 2, sgn, A.rtol, A.type, A.head, A.tcol, nullptr);
                __syncthreads();
            } else
                eng_bar(X, A);
            eng_mark(X, A, 4, (pass2 ? 45.0 * m : 0.0) + (pse ? 8.0 * S.k * (double)S.k : 0.0));
            if (S.status != ST_OK) break;
        }
        const int p = S.p;
        if (p >= 0) {
            /* ---- P6: rho and the tail of u ---- */
            eng_rho(X, A, S.k, p);
            if (pse)
                for (int r = X.gtid; r < m; r += X.gsize) {
                    const int cs = A.cslot[r];
                    A.u[r] = (cs >= 0) ? A.zn[cs] : A.v[A.bind[r]];
                }
            eng_bar(X, A);
            eng_mark(X, A, 5, 12.0 * m + 8.0 * S.k + (pse ? 20.0 * m : 0.0));
            /* ---- P7: pivot row and PSE inner products ---- */
            {
                Key dummy = {0.0, 0.0, 0.0, 0, 0};
                ENG_GROUPS(A.gc, (eng_trow<GG, false>(X, A, S, dummy)));
                eng_bar(X, A);
                if (X.tid == 0) {
                    /* k_primal_piv: lib/glpspx01.js:1985-2006 */
                    const double piv1 = A.tcol[p], piv2 = A.trow[q];
                    S.piv1 = piv1; S.piv2 = piv2;
                    if (fabs(piv1 - piv2) > 1e-8 * (1.0 + fabs(piv1)) ||
                        !((piv1 > 0.0 && piv2 > 0.0) || (piv1 < 0.0 && piv2 < 0.0)))
                        S.status = ST_PIV12;
                    else
                        S.new_dq = S.d1 / piv2;
                    if (S.status == ST_OK) eng_describe_change(A, S, C);
                }
                __syncthreads();
                eng_mark(X, A, 6, 12.0 * (double)__ldg(A.a_ptr + n) * (1.0 - (double)S.k / n) + 13.0 * n + 8.0 * m);
                if (S.status != ST_OK) break;
            }
        }
        /* ---- P8: update_bbar / update_cbar / update_gamma / basis change ---- */
        {
            const double teta = S.teta;
            const double xq = get_xN(A.stat, A.head, A.lb, A.ub, m, q);
            if (p >= 0) {
                const double new_dq = S.new_dq, pivot = A.trow[q];
                const int kp = C.kp;
                const int phase = S.phase;
                for (int t = X.gtid; t < n; t += X.gsize) {
                    if (t == q) {
                        double c = new_dq;
                        if (phase == 1) c -= A.coef[kp];
                        A.cbar[q] = c;
                        if (pse) {
                            double g = 1.0;
                            if (A.type[kp] != GLP_FX) {
                                g = S.gamma_q / (pivot * pivot);
                                if (g < DBL_EPSILON) g = DBL_EPSILON;
                            }
                            A.gamma[q] = g;
                        }
                    } else {
                        const double tr = A.trow[t];
                        if (tr != 0.0) {
                            A.cbar[t] -= tr * new_dq;
                            if (pse) {
                                const double tt = tr / pivot;
                                const int k = A.head[m + t];
                                const double t1 = A.gamma[t] + tt * tt * S.gamma_q + 2.0 * tt * A.svec[t];
                                const double t2 = (A.refsp[k] ? 1.0 : 0.0) + S.delta_q * tt * tt;
                                double g = (t1 >= t2 ? t1 : t2);
                                if (g < DBL_EPSILON) g = DBL_EPSILON;
                                A.gamma[t] = g;
                            }
                        }
                    }
                }
            }
            for (int t = X.gtid; t < m; t += X.gsize) {
                if (t == p) A.bbar[t] = xq + teta;
                else if (teta != 0.0) {
                    const double tc = A.tcol[t];
                    if (tc != 0.0) A.bbar[t] += tc * teta;
                }
            }
            /* keep the dense right-hand side of eval_tcol all-zero */
            if (X.cta == 0) {
                if (kq < m) { if (X.tid == 0) A.hz[kq] = 0.0; }
                else
                    for (int ptr = __ldg(A.a_ptr + (kq - m)) + X.tid; ptr < __ldg(A.a_ptr + (kq - m) + 1); ptr += ENG_NT)
                        A.hz[__ldg(A.a_ind + ptr)] = 0.0;
            }
            if (p >= 0 && S.k + (C.bnew >= 0) > 0) eng_update_T(X, A, C);
            /* the O(1) remainder runs after the barrier; readers of the next phase use (qprev, stprev) */
            int new_stat;
            if (p >= 0) new_stat = S.p_stat;
            else new_stat = (A.stat[q] == GLP_NL) ? GLP_NU : GLP_NL;
            eng_bar(X, A);
            eng_mark(X, A, 7, 24.0 * m + (p >= 0 ? 40.0 * n + 16.0 * S.k * (double)S.k : 0.0));
            if (X.cta == 0 && X.tid == 0) {
                if (p >= 0) {
                    if (S.phase == 1) A.coef[C.kp] = 0.0;       /* lib/glpspx01.js:2016-2020 */
                    eng_bookkeep(A, C, new_stat, false);
                } else
                    A.stat[q] = (signed char)new_stat;
            }
            if (X.tid == 0) {
                if (p >= 0) S.k = C.knew;
                iter_end(&S, p >= 0, 0);
            }
            qprev = q; stprev = new_stat;
            __syncthreads();
        }
    }
    if (X.cta == 0 && X.tid == 0) *A.ctrl = S;
}

/* ------------------------------------------------------------------ */
/* dual engine: lib/glpspx02.js:1780-1966                             */
/* ------------------------------------------------------------------ */
__global__ void __launch_bounds__(ENG_NT, 1) k_engine_dual(EngArgs A)
{
    extern __shared__ double eng_dyn[];
    __shared__ Ctrl S;
    __shared__ EngChange C;
    EngCtx X;
    eng_init(X, A, eng_dyn);
    const int m = A.m, n = A.n;
    if (X.tid == 0) S = *A.ctrl;
    __syncthreads();
    int pprev = -1, kprev = 0;
    for (int it = 0; it < A.max_iters && S.status == ST_OK; it++) {
        /* ---- D1: chuzr ---- */
        {
            Key none = {0.0, 0.0, 0.0, INT_MAX, 0};
            Key v = none;
            scan_chuzr_dual(v, X.gtid, X.gsize, m, A.type, A.lb, A.ub, A.head, A.bbar, A.gamma, A.tol_bnd,
                            pprev, kprev);
            Key r = eng_allreduce(X, A, v, none, 0, CombArgMax());
            if (X.tid == 0) {
                const bool found = (r.a > 0.0 && r.pos != INT_MAX);
                S.p = found ? r.pos : P_NONE;
                S.delta = found ? r.b : 0.0;
                S.big = 0.0;
                if (!found) S.status = ST_NONE1;
            }
            __syncthreads();
            eng_mark(X, A, 0, 37.0 * m);
            if (S.status != ST_OK) break;
        }
        const int p = S.p;
        const bool pse = gamma_on(&S);
        const double sgn = (S.delta > 0.0 ? +1.0 : -1.0);
        const double nnzA = (double)__ldg(A.a_ptr + n);
        /* ---- D2: rho ---- */
        eng_rho(X, A, S.k, p);
        eng_bar(X, A);
        eng_mark(X, A, 1, 12.0 * m + 8.0 * S.k);
        /* ---- D3: pivot row, |trow|_inf, sum of squares over the reference space ---- */
        {
            Key none = {0.0, 0.0, 0.0, 0, 0};
            Key acc = none;
            ENG_GROUPS(A.gc, (eng_trow<GG, true>(X, A, S, acc)));
            Key r = eng_allreduce(X, A, acc, none, 1, CombSum2());
            if (X.tid == 0) {
                S.trow_max = r.c;
                S.eps = A.tol_bnd * (1.0 + 0.01 * r.c);      /* sic: tol_bnd, lib/glpspx02.js:1851 */
                S.scal = r.a;                                 /* sum of trow_j^2, j in the reference space */
            }
            __syncthreads();
            eng_mark(X, A, 2, 12.0 * nnzA * (1.0 - (double)S.k / n) + 13.0 * n + 8.0 * m);
        }
        /* ---- D4: Harris pass 1 (+ right-hand side of update_gamma) ---- */
        {
            Key none = {DBL_MAX, 0.0, 0.0, INT_MAX, 0};
            Key v = none;
            scan_ratio_dual(v, X.gtid, X.gsize, 1, sgn, S.eps, 0.0, A.rtol, A.stat, A.cbar, A.trow, nullptr, n);
            if (pse) ENG_GROUPS(A.gr, (eng_gamma_rhs<GG>(X, A)));
            Key r = eng_allreduce(X, A, v, none, 2, CombRatio1());
            if (X.tid == 0) fin_ratio_dual(&S, r, 1, sgn, A.rtol, A.trow, nullptr);
            __syncthreads();
            eng_mark(X, A, 3, 17.0 * n + (pse ? 17.0 * nnzA + 16.0 * m : 0.0));
            if (S.status != ST_OK) break;
        }
        /* ---- D5: Harris pass 2 (+ dense product y2 = T v_N of update_gamma) ---- */
        {
            const bool pass2 = !S.skip2;
            Key none = {DBL_MAX, 0.0, 0.0, INT_MAX, 0};
            Key v = none;
            if (pass2)
                scan_ratio_dual(v, X.gtid, X.gsize, 2, sgn, S.eps, S.tmax, A.rtol, A.stat, A.cbar, A.trow, nullptr, n);
            if (pse && S.k > 0) {
                const int k = S.k;
                const double *val = A.wk;
                if (k <= A.dcap && X.cta * 8 < k) {
                    for (int e = X.tid; e < k; e += ENG_NT) X.sh_d[e] = A.wk[e];
                    __syncthreads();
                    val = X.sh_d;
                }
                eng_gemv_rows(X, A, k, k, nullptr, val, A.yk2, false);
            }
            if (pass2) {
                Key r = eng_allreduce(X, A, v, none, 3, CombRatio2());
                if (X.tid == 0) fin_ratio_dual(&S, r, 2, sgn, A.rtol, A.trow, nullptr);
                __syncthreads();
            } else
                eng_bar(X, A);
            eng_mark(X, A, 4, (pass2 ? 17.0 * n : 0.0) + (pse ? 8.0 * S.k * (double)S.k + 16.0 * S.k : 0.0));
            if (S.status != ST_OK) break;
        }
        const int q = S.q;
        const int kq = A.head[m + q];
        /* ---- D6: tcol, first half (+ tail of u = inv(B) v) ---- */
        eng_ftran_head_col(X, A, S.k, kq, A.yk);
        if (pse) {
            Key dummy = {0.0, 0.0, 0.0, 0, 0};
            ENG_GROUPS(A.gr, (eng_ftran_tail<GG, false>(X, A, S, A.v, A.yk2, A.u, dummy)));
        }
        eng_bar(X, A);
        eng_mark(X, A, 5, (pse ? 12.0 * nnzA * (1.0 - (double)S.k / m) + 16.0 * m : 0.0) + 8.0 * S.k);
        /* ---- D7: tcol, second half ---- */
        {
            Key dummy = {0.0, 0.0, 0.0, 0, 0};
            ENG_GROUPS(A.gr, (eng_ftran_tail<GG, false>(X, A, S, A.hz, A.yk, A.tcol, dummy)));
            eng_bar(X, A);
            if (X.tid == 0) {
                /* k_dual_prep: lib/glpspx02.js:1913-1938, :1103-1115 */
                double piv1 = A.tcol[p];
                const double piv2 = A.trow[q];
                S.piv1 = piv1; S.piv2 = piv2;
                if (fabs(piv1 - piv2) > 1e-8 * (1.0 + fabs(piv1)) ||
                    !((piv1 > 0.0 && piv2 > 0.0) || (piv1 < 0.0 && piv2 < 0.0)))
                    S.status = ST_PIV12;
                else {
                    const double eta = (A.refsp[A.head[p]] ? 1.0 : 0.0);
                    S.delta_q = eta;
                    S.gamma_q = eta + (pse ? S.scal : 0.0);
                    S.teta = S.delta / piv1;
                    if (S.phase == 2) S.obj += (A.cbar[q] / S.zeta) * (S.delta / piv1);
                    eng_describe_change(A, S, C);
                }
            }
            __syncthreads();
            eng_mark(X, A, 6, 12.0 * nnzA * (1.0 - (double)S.k / m) + 16.0 * m);
            if (S.status != ST_OK) break;
        }
        /* ---- D8: update_cbar / update_bbar / update_gamma / basis change ---- */
        {
            const double teta = S.teta, new_dq = S.new_dq;
            const int kp = C.kp;
            const double pivot = C.tp;
            const double xq = get_xN(A.stat, A.head, A.lb, A.ub, m, q);
            const bool drop = (A.type[kp] == GLP_FX && A.refsp[kp]);
            for (int t = X.gtid; t < n; t += X.gsize) {
                if (t == q) A.cbar[q] = new_dq;
                else if (new_dq != 0.0) {
                    const double tr = A.trow[t];
                    if (tr != 0.0) A.cbar[t] -= tr * new_dq;
                }
            }
            for (int t = X.gtid; t < m; t += X.gsize) {
                if (t == p) {
                    A.bbar[p] = xq + teta;
                    if (pse) {
                        double g = 1.0;
                        if (A.type[kq] != GLP_FR) {
                            g = S.gamma_q / (pivot * pivot);
                            if (g < DBL_EPSILON) g = DBL_EPSILON;
                            if (drop) {
                                const double tt = 1.0 / pivot;
                                g -= tt * tt;
                                if (g < DBL_EPSILON) g = DBL_EPSILON;
                            }
                        }
                        A.gamma[p] = g;
                    }
                } else {
                    const double tc = A.tcol[t];
                    if (tc != 0.0) {
                        if (teta != 0.0) A.bbar[t] += tc * teta;
                        const int k = A.head[t];
                        if (pse && A.type[k] != GLP_FR) {
                            const double tt = tc / pivot;
                            const double t1 = A.gamma[t] + tt * tt * S.gamma_q + 2.0 * tt * A.u[t];
                            const double t2 = (A.refsp[k] ? 1.0 : 0.0) + S.delta_q * tt * tt;
                            double g = (t1 >= t2 ? t1 : t2);
                            if (g < DBL_EPSILON) g = DBL_EPSILON;
                            if (drop) {
                                g -= tt * tt;
                                if (g < DBL_EPSILON) g = DBL_EPSILON;
                            }
                            A.gamma[t] = g;
                        }
                    }
                }
            }
            if (X.cta == 0) {
                if (kq < m) { if (X.tid == 0) A.hz[kq] = 0.0; }
                else
                    for (int ptr = __ldg(A.a_ptr + (kq - m)) + X.tid; ptr < __ldg(A.a_ptr + (kq - m) + 1); ptr += ENG_NT)
                        A.hz[__ldg(A.a_ind + ptr)] = 0.0;
            }
            if (S.k + (C.bnew >= 0) > 0) eng_update_T(X, A, C);
            eng_bar(X, A);
            eng_mark(X, A, 7, 24.0 * n + 40.0 * m + 16.0 * S.k * (double)S.k);
            const int new_stat = (A.type[kp] == GLP_FX) ? GLP_NS : (S.delta > 0.0 ? GLP_NL : GLP_NU);
            if (X.cta == 0 && X.tid == 0) eng_bookkeep(A, C, new_stat, pse && drop);
            if (X.tid == 0) {
                S.k = C.knew;
                iter_end(&S, true, 1);
            }
            pprev = p; kprev = kq;
            __syncthreads();
        }
    }
    if (X.cta == 0 && X.tid == 0) *A.ctrl = S;
}

#endif /* GLPB_ENGINE_CUH */
