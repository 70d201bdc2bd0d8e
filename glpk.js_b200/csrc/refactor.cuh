/* refactor.cuh -- refactorisation of the basis inverse in ONE cooperative
 * launch (replaces invert_B -> bfd_factorize -> luf_factorize of the reference,
 * lib/glpspx01.js:177-181, lib/glpluf.js:1105-1225, for the explicit inverse
 * T = -inv(A[R_N, J_B]) that libglpb200 keeps; see glpb_internal.cuh).
 *
 * Blocked Gauss-Jordan with partial pivoting, NB pivots per round:
 *   1. panel: the rows of the NB panel columns are distributed over the CTAs
 *      and live in shared memory for the whole round.  One grid-wide
 *      arg-reduction per pivot: every CTA publishes its best candidate row
 *      TOGETHER with that row's panel entries (and the owner of row t publishes
 *      row t), so that after the reduction every CTA can swap and eliminate its
 *      own rows without a second exchange.
 *   2. the round's row swaps on every other column (a gather/scatter over the
 *      <= 2 NB affected rows, not NB sequential swaps) and a copy of their
 *      pivot-row entries;
 *   3. rank-NB update of the other columns  X[i,c] = (i in P ? 0 : X[i,c]) +
 *      sum_t Pf[i,t] Xp[t,c]  -- one read+write of X per round.
 * The final column permutation that undoes the row pivoting and the negation
 * are one out-of-place streaming pass into the second buffer (the handle flips
 * its T pointers).  Ties in the pivot search go to the lowest row index, so the
 * result does not depend on the grid size or on scheduling.
 */
#ifndef GLPB_REFACTOR_CUH
#define GLPB_REFACTOR_CUH
#include "engine.cuh"

#define REF_NT 1024
#define REF_NB 32            /* pivots per round                              */
#define REF_CB 16            /* columns per update block                      */
#define REF_RMAX 256         /* distributed panel: max rows per CTA            */
#define REF_PS 132           /* shared row strides of the DMMA operands        */
#define REF_XS 68
#define REF_UPD_SMEM ((REF_NB * REF_XS) * sizeof(double))

struct __align__(128) RefSlot {
    double absval;
    int row;
    unsigned int flag;
    double cand[REF_NB];     /* panel entries of the candidate row            */
    double rowt[REF_NB];     /* panel entries of row t (owner of t only)      */
};

struct RefArgs {
    Ctrl *ctrl;
    int m, ldt;
    int nbr;                 /* pivots per round (8, 16 or 32)                */
    int single;              /* 1: the whole panel lives in CTA 0's shared memory (k nbr 8 bytes fit);
                                0: its rows are spread over all CTAs          */
    int pg;                  /* distributed mode: CTAs that hold the panel (a few dozen: the per-pivot
                                exchange gets cheaper with fewer participants)                       */
    double *xp;              /* [REF_NB][ldt] pivot rows of the round (after the swaps) by column    */
    int *plan;               /* [2 + 4 REF_NB] naff, sing, aff_row, aff_src published by CTA 0 for the
                                CTAs that did not follow the pivots                                  */
    long long *prof_cyc;     /* optional [8]: SM cycles of CTA 0 per phase (build, panel, panel i/o, wait, swap+update, finish) */
    double *X;               /* working matrix (in: nothing, built here)      */
    double *T2;              /* out: T = -inv(M), column-major, ld = ldt      */
    const int *a_ptr, *a_ind; const double *a_val;
    const int *head, *slot_pos, *cslot;
    int *piv;                /* [ldt] pivot rows                              */
    RefSlot *slots;          /* [ENG_RING][ENG_MAXG], zeroed by the host      */
    unsigned int *flags;     /* [ENG_RING][ENG_MAXG*32] plain barrier flags, one line each, zeroed */
};

struct RefCtx {
    int G, cta, tid, lane, warp;
    unsigned int seq;        /* grid barriers (all CTAs)                         */
    unsigned int pseq;       /* per-pivot exchanges (panel CTAs only)            */
};

__device__ __forceinline__ void ref_bar(RefCtx &X, const RefArgs &A)
{
    __syncthreads();
    X.seq++;
    if (X.G > 1) {
        unsigned int *ring = A.flags + (size_t)(X.seq & (ENG_RING - 1)) * ENG_MAXG * 32;
        if (X.tid == 0) eng_st_release(ring + X.cta * 32, X.seq);
        if (X.tid < X.G) {
            while (eng_ld_relaxed(ring + X.tid * 32) < X.seq) { }
            eng_fence_acq();
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(REF_NT, 1) k_refactor(RefArgs A)
{
    extern __shared__ double ref_dyn[];       /* pan[R][nbr], fcol[R] | later: int map[k], pv[k] */
    __shared__ double rowR[REF_NB], rowT[REF_NB], srow[REF_NB];
    __shared__ double xs[REF_NB][REF_CB];
    __shared__ int s_piv[REF_NB];
    __shared__ int aff_row[2 * REF_NB], aff_src[2 * REF_NB];
    __shared__ int s_naff, s_r, s_wc, s_sing;
    __shared__ Key wres[ENG_MAXG / 32];
    __shared__ double s_max[32];

    long long t_last = clock64();
#define REF_MARK(ph)                                                           \
    do {                                                                       \
        if (A.prof_cyc != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {    \
            const long long t_now = clock64();                                 \
            A.prof_cyc[ph] += t_now - t_last;                                  \
            t_last = t_now;                                                    \
        }                                                                      \
    } while (0)
    RefCtx X;
    X.G = gridDim.x; X.cta = blockIdx.x; X.tid = threadIdx.x; X.lane = X.tid & 31; X.warp = X.tid >> 5;
    X.seq = 0u; X.pseq = 0u;
    const int G = X.G, cta = X.cta, tid = X.tid;
    const int k = A.ctrl->k;
    if (A.ctrl->sing || k <= 0) return;
    const size_t ldt = (size_t)A.ldt;
    const int m = A.m;
    double *pan = ref_dyn;

    /* ---- build M[cs, b] = A[row_cs, j_b] (column b contiguous) ---- */
    {
        double mx = 0.0;
        for (int b = cta; b < k; b += G) {
            double *col = A.X + (size_t)b * ldt;
            for (int cs = tid; cs < k; cs += REF_NT) col[cs] = 0.0;
            __syncthreads();
            const int j = A.head[A.slot_pos[b]] - m;
            for (int ptr = __ldg(A.a_ptr + j) + tid; ptr < __ldg(A.a_ptr + j + 1); ptr += REF_NT) {
                const int cs = A.cslot[__ldg(A.a_ind + ptr)];
                if (cs >= 0) { const double a = __ldg(A.a_val + ptr); col[cs] = a; mx = fmax(mx, fabs(a)); }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(FULLMASK, mx, off));
        if (X.lane == 0) s_max[X.warp] = mx;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 32; w++) mx = fmax(mx, s_max[w]);
            if (mx > 0.0) atomic_max_abs(&A.ctrl->max_a, mx);
        }
    }
    ref_bar(X, A);
    REF_MARK(0);
    const double max_a = __ldcg(&A.ctrl->max_a);
    const double tiny = 1e-13 * fmax(max_a, 1e-300);

    /* panel ownership: PG CTAs hold R rows each */
    const int PG = A.single ? 1 : min(G, A.pg);
    const int NBR = A.nbr;
    const int R = (k + PG - 1) / PG;
    const int r0 = (cta < PG) ? min(k, cta * R) : k, r1 = min(k, r0 + R);
    const int nloc = r1 - r0;
    const int PS = NBR + 1;                   /* odd row stride: column walks are bank-conflict free */
    double *fcol = pan + (size_t)R * PS;

    for (int c0 = 0; c0 < k; c0 += NBR) {
        const int nb = min(NBR, k - c0);
        /* ---- panel into shared memory ---- */
        for (int e = tid; e < nloc * nb; e += REF_NT) {
            const int i = e % nloc, cc = e / nloc;
            pan[i * PS + cc] = A.X[(size_t)(c0 + cc) * ldt + r0 + i];
        }
        if (tid < nb) { aff_row[tid] = c0 + tid; aff_src[tid] = c0 + tid; }
        if (tid == 0) { s_naff = nb; s_sing = 0; }
        __syncthreads();
        REF_MARK(2);
        if (A.single && cta == 0) {
            /* ---- whole panel in this CTA's shared memory, warp-synchronous:
               every warp owns a contiguous block of rows (lane = panel column), so a
               pivot needs two block-wide barriers: candidates -> pivot row -> eliminate.
               Rows do not move inside the panel; the swaps are recorded (piv, gather
               list) and applied when the panel is written back. ---- */
            unsigned char *mark = (unsigned char *)(pan + (size_t)R * PS);
            int *rowdest = (int *)(mark + ((k + 15) & ~15));
            for (int i = tid; i < k; i += REF_NT) { mark[i] = 0; rowdest[i] = i; }
            __syncthreads();
            const int RW = (k + 31) / 32;
            const int w0 = min(k, X.warp * RW), w1 = min(k, w0 + RW);
            const int RS = 32 / NBR, sub = X.lane / NBR, cc = X.lane % NBR;
            /* candidates of the first pivot column; later ones are tracked while eliminating */
            double bv = -1.0;
            int bi = INT_MAX;
            for (int i = w0 + X.lane; i < w1; i += 32)
                if (i >= c0) {
                    const double a = fabs(pan[i * PS]);
                    if (a > bv) { bv = a; bi = i; }
                }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_xor_sync(FULLMASK, bv, off);
                const int oi = __shfl_xor_sync(FULLMASK, bi, off);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            int *s_wi = (int *)&xs[1][0];
            for (int tt = 0; tt < nb; tt++) {
                if (X.lane == 0) { s_max[X.warp] = bv; s_wi[X.warp] = bi; }
                __syncthreads();
                bv = s_max[X.lane]; bi = s_wi[X.lane];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const double ov = __shfl_xor_sync(FULLMASK, bv, off);
                    const int oi = __shfl_xor_sync(FULLMASK, bi, off);
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                REF_MARK(6);
                if (!(bv > tiny)) { if (tid == 0) s_sing = 1; break; }
                const int r = bi;
                const double rr = (X.lane < nb) ? pan[r * PS + X.lane] : 0.0;
                const double ipv = 1.0 / __shfl_sync(FULLMASK, rr, tt);
                const double sr = (X.lane == tt) ? ipv : rr * ipv;
                if (X.warp == 0) {
                    /* position currently holding physical row r, and the transposition (c0+tt <-> it) */
                    const int na = s_naff;
                    const bool h0 = (X.lane < na) && (aff_src[X.lane] == r);
                    const bool h1 = (32 + X.lane < na) && (aff_src[32 + X.lane] == r);
                    const unsigned int b0 = __ballot_sync(FULLMASK, h0), b1 = __ballot_sync(FULLMASK, h1);
                    int ib;
                    if (b0) ib = __ffs(b0) - 1;
                    else if (b1) ib = 32 + __ffs(b1) - 1;
                    else {
                        ib = na;
                        if (X.lane == 0) { aff_row[na] = r; aff_src[na] = r; s_naff = na + 1; }
                    }
                    __syncwarp();
                    if (X.lane == 0) {
                        s_piv[tt] = aff_row[ib];
                        const int tmp = aff_src[tt]; aff_src[tt] = aff_src[ib]; aff_src[ib] = tmp;
                        mark[r] = 1;
                    }
                }
                __syncthreads();
                REF_MARK(7);
                /* eliminate this warp's rows, four at a time; the lane of column tt+1
                   keeps the best unpivoted entry of the next pivot column */
                const double scc = __shfl_sync(FULLMASK, sr, cc);
                const bool track = (cc == tt + 1) && (tt + 1 < nb);
                bv = -1.0; bi = INT_MAX;
                for (int i0 = w0; i0 < w1; i0 += 4 * RS) {          /* uniform trip count per warp */
                    double f[4], old[4];
                    bool act[4];
#pragma unroll
                    for (int x = 0; x < 4; x++) {
                        const int i = i0 + x * RS + sub;
                        act[x] = (i < w1) && (cc < nb);
                        const double *px = pan + (act[x] ? i : w0) * PS;
                        f[x] = px[tt]; old[x] = px[act[x] ? cc : 0];
                    }
                    __syncwarp();
#pragma unroll
                    for (int x = 0; x < 4; x++) {
                        const int i = i0 + x * RS + sub;
                        if (act[x]) {
                            const double nv = (i == r) ? scc : ((cc == tt) ? -f[x] * scc : old[x] - f[x] * scc);
                            pan[i * PS + cc] = nv;
                            if (track && i >= c0 && !mark[i]) {
                                const double a = fabs(nv);
                                if (a > bv) { bv = a; bi = i; }
                            }
                        }
                    }
                }
                for (int off = NBR; off < 32; off <<= 1) {
                    const double ov = __shfl_xor_sync(FULLMASK, bv, off);
                    const int oi = __shfl_xor_sync(FULLMASK, bi, off);
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                /* lane (tt+1) of sub-row 0 holds the warp's candidate: hand it to lane 0 */
                bv = __shfl_sync(FULLMASK, bv, (tt + 1) & 31);
                bi = __shfl_sync(FULLMASK, bi, (tt + 1) & 31);
                __syncwarp();
                REF_MARK(1);
            }
            __syncthreads();
            REF_MARK(1);
            /* write back; position aff_row[a] receives physical row aff_src[a] */
            if (tid < s_naff) rowdest[aff_src[tid]] = aff_row[tid];
            __syncthreads();
            for (int e = tid; e < k * nb; e += REF_NT) {
                const int i = e % k, c = e / k;
                A.X[(size_t)(c0 + c) * ldt + rowdest[i]] = pan[i * PS + c];
            }
        }
        for (int tt = 0; tt < nb && cta < PG && !A.single; tt++) {
            const int t = c0 + tt;
            /* best local candidate among rows >= t */
            Key none = {-1.0, 0.0, 0.0, INT_MAX, 0};
            Key v = none;
            for (int i = tid; i < nloc; i += REF_NT)
                if (r0 + i >= t) {
                    Key c = {fabs(pan[i * PS + tt]), 0.0, 0.0, r0 + i, 0};
                    CombArgMax()(v, c);
                }
            v = block_reduce(v, none, CombArgMax());
            if (tid == 0) s_r = v.pos;
            __syncthreads();
            int r, wc;
            if (PG > 1) {
                X.pseq++;
                RefSlot *ring = A.slots + (X.pseq & (ENG_RING - 1)) * ENG_MAXG;
                RefSlot *mine = ring + cta;
                const int lr = s_r;
                if (tid < nb) {
                    if (lr != INT_MAX) __stcg(&mine->cand[tid], pan[(lr - r0) * PS + tid]);
                    if (t >= r0 && t < r1) __stcg(&mine->rowt[tid], pan[(t - r0) * PS + tid]);
                }
                if (tid == 0) { __stcg(&mine->absval, v.a); __stcg(&mine->row, v.pos); }
                __syncthreads();
                if (tid == 0) eng_st_release(&mine->flag, X.pseq);
                const int nw = (PG + 31) >> 5;
                if (X.warp < nw) {
                    Key c = none;
                    if (tid < PG) {
                        while (eng_ld_relaxed(&ring[tid].flag) < X.pseq) { }
                        eng_fence_acq();
                        c.a = __ldcg(&ring[tid].absval); c.pos = __ldcg(&ring[tid].row); c.aux = tid;
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        Key o = key_shfl_down(c, off);
                        CombArgMax()(c, o);
                    }
                    if (X.lane == 0) wres[X.warp] = c;
                }
                __syncthreads();
                if (tid == 0) {
                    Key c = wres[0];
                    for (int w = 1; w < nw; w++) CombArgMax()(c, wres[w]);
                    s_r = c.pos; s_wc = c.aux;
                    s_sing = !(c.a > tiny);
                }
                __syncthreads();
                r = s_r; wc = s_wc;
                if (s_sing) break;             /* reported after the round's barrier */
                const int ot = min(PG - 1, t / R);
                if (tid < nb) {
                    rowR[tid] = __ldcg(&ring[wc].cand[tid]);
                    rowT[tid] = __ldcg(&ring[ot].rowt[tid]);
                }
            } else {
                if (tid == 0) s_sing = !(v.a > tiny);
                __syncthreads();
                r = s_r; wc = 0;
                if (s_sing) break;             /* reported after the round's barrier */
                if (tid < nb) { rowR[tid] = pan[r * PS + tid]; rowT[tid] = pan[t * PS + tid]; }
            }
            __syncthreads();
            if (tid == 0) s_piv[tt] = r;
            /* the round's net row permutation, kept as a gather list: slots
               0..nb-1 are the panel rows, further slots the outside pivot rows */
            if (X.warp == 0) {
                int ib;
                if (r < c0 + nb) ib = r - c0;
                else {
                    const int na = s_naff;
                    const bool hit = (nb + X.lane < na) && (aff_row[nb + X.lane] == r);
                    const unsigned int bm = __ballot_sync(FULLMASK, hit);
                    if (bm) ib = nb + __ffs(bm) - 1;
                    else {
                        ib = na;
                        if (X.lane == 0) { aff_row[na] = r; aff_src[na] = r; s_naff = na + 1; }
                    }
                    __syncwarp();
                }
                if (X.lane == 0 && r != t) { const int tmp = aff_src[tt]; aff_src[tt] = aff_src[ib]; aff_src[ib] = tmp; }
            }
            /* swap rows t and r inside the panel */
            if (r != t && tid < nb) {
                if (r >= r0 && r < r1) pan[(r - r0) * PS + tid] = rowT[tid];
                if (t >= r0 && t < r1) pan[(t - r0) * PS + tid] = rowR[tid];
            }
            const double ipv = 1.0 / rowR[tt];
            if (tid < nb) srow[tid] = (tid == tt) ? ipv : rowR[tid] * ipv;
            __syncthreads();
            for (int i = tid; i < nloc; i += REF_NT) fcol[i] = pan[i * PS + tt];
            __syncthreads();
            for (int e = tid; e < nloc * nb; e += REF_NT) {
                const int i = e / nb, cc = e - i * nb;
                double *x = pan + i * PS + cc;
                const double f = fcol[i];
                if (r0 + i == t) *x = srow[cc];
                else if (cc == tt) *x = -f * srow[cc];
                else *x -= f * srow[cc];
            }
            __syncthreads();
        }
        if (!A.single) REF_MARK(1);
        /* ---- panel back to memory; the round's row permutation as a gather ---- */
        for (int e = tid; e < nloc * nb && !A.single; e += REF_NT) {
            const int i = e % nloc, cc = e / nloc;
            A.X[(size_t)(c0 + cc) * ldt + r0 + i] = pan[i * PS + cc];
        }
        if (cta == 0 && tid < nb) A.piv[c0 + tid] = s_piv[tid];
        if (cta == 0) {
            /* the CTAs that did not follow the pivots take the gather list from here */
            __syncthreads();
            if (tid < 2 * REF_NB) { A.plan[2 + tid] = aff_row[tid]; A.plan[2 + 2 * REF_NB + tid] = aff_src[tid]; }
            if (tid == 0) { A.plan[0] = s_naff; A.plan[1] = s_sing; }
        }
        REF_MARK(2);
        ref_bar(X, A);
        REF_MARK(3);
        if (cta >= PG) {
            if (tid < 2 * REF_NB) { aff_row[tid] = __ldcg(A.plan + 2 + tid); aff_src[tid] = __ldcg(A.plan + 2 + 2 * REF_NB + tid); }
            if (tid == 0) { s_naff = __ldcg(A.plan + 0); s_sing = __ldcg(A.plan + 1); }
        }
        __syncthreads();
        if (s_sing) { if (cta == 0 && tid == 0) A.ctrl->sing = 1; return; }
        /* ---- the other columns: row swaps, pivot rows, rank-nb update ---- */
        const int naff = s_naff;
        /* (2) the round's row permutation on the other columns, 16 columns per CTA step;
               their pivot rows (after the swaps) go to xp[t][col] */
        for (int cb = cta; cb < (k + REF_CB - 1) / REF_CB; cb += G) {
            const int a = tid / REF_CB, c = tid % REF_CB;
            const int col = cb * REF_CB + c;
            const bool inpanel = (col >= c0 && col < c0 + nb);
            const bool act = (a < naff) && (col < k) && !inpanel;
            double val = 0.0;
            if (act) val = A.X[(size_t)col * ldt + aff_src[a]];
            __syncthreads();
            if (act && aff_src[a] != aff_row[a]) A.X[(size_t)col * ldt + aff_row[a]] = val;
            __syncthreads();
            if (tid < REF_NB * REF_CB) {
                const int t = tid / REF_CB;
                if (col < (int)ldt)
                    A.xp[(size_t)t * ldt + col] = (t < nb && col < k && !inpanel) ? A.X[(size_t)col * ldt + c0 + t] : 0.0;
            }
        }
        ref_bar(X, A);
        /* (3) rank-nb update on the fp64 tensor cores (DMMA m8n8k4).  A unit is a block of
               64 columns times a chunk of rows; the 32 x 64 slice of the pivot rows is staged
               in shared memory (row stride 68: bank-conflict-free fragment loads), then every
               warp walks its own 8-row blocks with no block-wide barrier, the Pf fragments
               coming straight from L2: the DRAM latency of one warp's tile overlaps the
               arithmetic of the others */
        double *Rs2 = ref_dyn;                     /* [REF_NB][REF_XS] */
        const int ncb = (k + 63) >> 6;
        int nrc = (3 * G + ncb - 1) / ncb;
        nrc = max(1, min(nrc, (k + 127) >> 7));
        const int RCH = ((((k + nrc - 1) / nrc) + 127) >> 7) << 7;
        const int ksteps = (nb + 3) >> 2;
        const int g = X.lane >> 2, t4 = X.lane & 3;
        const int rbk = X.warp & 15, cq = X.warp >> 4;
        for (int u = cta; u < ncb * nrc; u += G) {
            const int cc0 = (u % ncb) << 6, q0 = (u / ncb) * RCH, q1 = min(k, q0 + RCH);
            __syncthreads();
            for (int e = tid; e < REF_NB * 64; e += REF_NT) {
                const int t = e >> 6, c = e & 63, col = cc0 + c;
                Rs2[t * REF_XS + c] = (t < nb && col < k) ? __ldcg(A.xp + (size_t)t * ldt + col) : 0.0;
            }
            __syncthreads();
            for (int i0 = q0; i0 < q1; i0 += 128) {
                const int row = i0 + rbk * 8 + g;
                const bool rok = row < q1;
                const bool in_p = (row >= c0 && row < c0 + nb);
                double acc[4][2];
                double *tp[4];
                bool ok[4][2];
#pragma unroll
                for (int x = 0; x < 4; x++) {
                    const int col = cc0 + (cq * 4 + x) * 8 + 2 * t4;
                    tp[x] = A.X + (size_t)col * ldt + row;
                    ok[x][0] = rok && (col < k) && !(col >= c0 && col < c0 + nb);
                    ok[x][1] = rok && (col + 1 < k) && !(col + 1 >= c0 && col + 1 < c0 + nb);
                    acc[x][0] = (ok[x][0] && !in_p) ? __ldcg(tp[x]) : 0.0;
                    acc[x][1] = (ok[x][1] && !in_p) ? __ldcg(tp[x] + ldt) : 0.0;
                }
                double af[8];
#pragma unroll
                for (int kk = 0; kk < 8; kk++)
                    af[kk] = (4 * kk + t4 < nb && rok) ? __ldcg(A.X + (size_t)(c0 + 4 * kk + t4) * ldt + row) : 0.0;
#pragma unroll
                for (int kk = 0; kk < 8; kk++) {
                    if (kk < ksteps) {
#pragma unroll
                        for (int x = 0; x < 4; x++) {
                            const double b = Rs2[(4 * kk + t4) * REF_XS + (cq * 4 + x) * 8 + g];
                            eng_dmma(acc[x][0], acc[x][1], af[kk], b);
                        }
                    }
                }
#pragma unroll
                for (int x = 0; x < 4; x++) {
                    if (ok[x][0]) tp[x][0] = acc[x][0];
                    if (ok[x][1]) tp[x][ldt] = acc[x][1];
                }
            }
        }
        __syncthreads();
        REF_MARK(4);
        ref_bar(X, A);
        REF_MARK(3);
    }
    /* ---- undo the row pivoting (a column permutation) and negate, out of place ---- */
    int *map = (int *)ref_dyn;
    int *pv = map + k;
    for (int j = tid; j < k; j += REF_NT) { map[j] = j; pv[j] = A.piv[j]; }
    __syncthreads();
    if (tid == 0)
        for (int t = k - 1; t >= 0; t--) {
            const int r = pv[t];
            if (r != t) { const int x = map[t]; map[t] = map[r]; map[r] = x; }
        }
    __syncthreads();
    for (int j = cta; j < k; j += G) {
        const double *src = A.X + (size_t)map[j] * ldt;
        double *dst = A.T2 + (size_t)j * ldt;
        for (int b = tid; b < k; b += REF_NT) dst[b] = -src[b];
    }
    REF_MARK(5);
#undef REF_MARK
}

#endif /* GLPB_REFACTOR_CUH */
