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

#ifndef ENG_NT
#define ENG_NT 1024          /* threads per CTA                                  */
#endif
#define ENG_MAXG 160         /* scratch slot width (>= number of SMs)            */
#define ENG_LCAP 2048        /* staged entries of one sparse column / row        */
#define ENG_RING 4           /* rotating barrier slots                           */
#ifndef ENG_SU
#define ENG_SU 8            /* columns per trip of the dense stream's load loop   */
#endif
#ifndef ENG_DB
#define ENG_DB 32            /* deferred basis changes before T is rewritten (a multiple of 32) */
#endif

/* one 128-byte line per (ring slot, CTA): the arrival flag and the partial
   result that travels with it.  A line of its own keeps the G pollers of a
   flag from queueing behind the pollers of its neighbours in one L2 slice. */
struct __align__(128) EngSlot {
    Key key;
    unsigned int flag;
    unsigned int pad[23];
};

struct EngArgs {
    Ctrl *ctrl;
    int m, n, ldt, max_iters;
    int dcap;                 /* doubles of dynamic shared memory for staging    */
    double avg_col, avg_row;  /* average column / row length of A                */
    double tol_bnd, tol_dj, tol_piv, rtol;
    const int *a_ptr, *a_ind; const double *a_val;
    const int *at_ptr, *at_ind; const double *at_val;
    signed char *type, *stat, *refsp;
    const signed char *orig_type; /* types of the problem itself (dual phase 1 works on modified ones) */
    double *lb, *ub, *coef;
    int *head, *bind;
    double *bbar, *cbar, *gamma, *tcol, *trow, *rho, *svec;
    double *hz;               /* dense rhs of eval_tcol, all-zero between iterations */
    double *v, *u;            /* PSE work vectors [m]                            */
    double *yk, *yk2, *wk, *zn; /* kernel-space work [ldt]                       */
    /* copies indexed by structural column / by row, zero where the variable is
       not in the set: they turn the gathers of the sparse passes into one load */
    double *ycol, *ycol2;     /* [n] FTRAN results y by basic column (0: non-basic)        */
    double *vrow;             /* [m] primal: masked tcol by row of a basic auxiliary       */
    double *trowcol;          /* [n] dual: trow over the reference space by non-basic column */
    double *T;
    /* deferred basis changes (dual engine, large problems): the kernel in use is
       T + sum_{j<nd} Fd_j Rd_j', flushed into T every ENG_DB changes */
    double *Fd, *Rd;          /* [ENG_DB][ldt] each                              */
    double *zbuf;             /* [ENG_DB] Rd_j' v for the dense product          */
    int defer;                /* 1: deferral allowed                             */
    int use_tma;              /* 1: TMA-staged dense T*v stream where it applies */
    int hdr_smem;             /* 1: keep per-CTA copies of the basis header in shared memory */
    int local_max;            /* ratio tests up to this length are replicated per CTA */
    int pf_dist;              /* dense T*v stream: L2 prefetch distance in column steps (0 = off) */
    int use_async;            /* 1: dense T*v stream through cp.async into a shared-memory ring */
    int pf_first;             /* first of the 8 columns of a trip that is prefetched (experiments) */
    int tie_stop;             /* 1: an exact tie in a ratio test stops the engine (ST_TIE)  */
    int *rslot, *slot_pos, *cslot, *slot_row;
    EngSlot *slots;           /* [ENG_RING][ENG_MAXG] arrival flags + CTA partials, zeroed by the host */
    long long *prof_cyc;      /* optional: SM cycles per phase, CTA 0 (NULL = off)  */
    double *prof_bytes;       /* optional: algorithmic bytes per phase              */
};

template <bool HL> struct EngCtxT {
    int G, cta, tid, lane, warp, gtid, gsize;
    int vtid;                 /* grid-wide index of the plain grid-stride loops, see eng_init */
    unsigned int seq;         /* barrier sequence number, identical in every thread */
    unsigned int dseq;        /* sequence number of the header-update hand-off     */
    long long t_last;
    double *sh_d;             /* [dcap] staging values / dense vector            */
    int *sh_i;                /* [ENG_LCAP] staging indices                      */
    double *sh_red2;          /* [32][65] cross-warp sums of the dense T*v stream */
    /* the basis header and the slot maps as the engine reads them: the arrays in
       global memory, or -- small problems -- per-CTA copies in shared memory that
       every CTA keeps up to date itself (an index lookup then costs a shared-memory
       access instead of an L2 round trip inside every dependent-load chain) */
    int *head, *bind, *rslot, *slot_pos, *cslot, *slot_row;
    signed char *stat;
    int hdr_local;            /* 1: the pointers above are this CTA's private copies */
};

/* The basis header and the slot maps as the engine reads them: HL = true (small problems), per-CTA copies in
   shared memory whose pointers live in the context; HL = false, the arrays in global memory, addressed straight
   from the kernel parameters -- no pointer of the context stays live in a register (the 1024-thread kernel
   has 64 registers per thread; seven 64-bit pointers were a quarter of them). */
template <bool HL> __device__ __forceinline__ constexpr bool eng_hl(const EngCtxT<HL> &) { return HL; }
#define XHDR(f) (eng_hl(X) ? X.f : A.f)

__device__ __forceinline__ unsigned int eng_ld_relaxed(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void eng_st_release(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void eng_fence_acq() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

/* Grid-wide barrier without a shared counter: every CTA publishes the barrier
   sequence number in its own flag (release), thread i of every CTA polls the
   flag of CTA i.  No atomic is contended and the critical path is one store
   becoming visible in L2.  Slots rotate so that a CTA that runs ahead cannot
   overwrite a flag a slower CTA has not seen yet (it can be at most one
   barrier ahead). */
template <bool HL>
__device__ __forceinline__ void eng_arrive(EngCtxT<HL> &X, const EngArgs &A)
{
    __syncthreads();
    X.seq++;
    if (X.G > 1 && X.tid == 0) eng_st_release(&A.slots[(X.seq & (ENG_RING - 1)) * ENG_MAXG + X.cta].flag, X.seq);
}
template <bool HL>
__device__ __forceinline__ void eng_wait(EngCtxT<HL> &X, const EngArgs &A)
{
    if (X.G > 1) {
        if (X.tid < X.G) {
            const unsigned int *f = &A.slots[(X.seq & (ENG_RING - 1)) * ENG_MAXG + X.tid].flag;
            while (eng_ld_relaxed(f) < X.seq) { }
            eng_fence_acq();
        }
        __syncthreads();
    }
}
template <bool HL>
__device__ __forceinline__ void eng_bar(EngCtxT<HL> &X, const EngArgs &A)
{
    eng_arrive(X, A);
    eng_wait(X, A);
}

/* Hand-off of the O(1) header update (slot maps, head/bind, status of the
   entering column): CTA 0 performs it once EVERY CTA has arrived at the last
   barrier of the iteration -- nobody reads the old header any more -- and then
   raises this flag; nobody reads the new header before seeing it.  Costs one
   flag round trip instead of a second grid barrier. */
template <bool HL>
__device__ __forceinline__ void eng_header_done(EngCtxT<HL> &X, const EngArgs &A)
{
    X.dseq++;
    if (X.G > 1 && !HL) {
        unsigned int *done = &A.slots[ENG_RING * ENG_MAXG].flag;
        __syncthreads();        /* CTA 0: thread 0's header writes are ordered before the release */
        if (X.tid == 0) {
            if (X.cta == 0) eng_st_release(done, X.dseq);
            else {
                while (eng_ld_relaxed(done) < X.dseq) { }
                eng_fence_acq();
            }
        }
    }
    __syncthreads();
}

/* phase accounting for bench.py's roofline leg: CTA 0 leaves every barrier
   together with the grid, so its cycle stamps bound the phase */
template <bool HL>
__device__ __forceinline__ void eng_mark(EngCtxT<HL> &X, const EngArgs &A, int phase, double bytes)
{
    if (A.prof_cyc != nullptr && X.cta == 0 && X.tid == 0) {
        const long long t = clock64();
        A.prof_cyc[phase] += t - X.t_last;
        A.prof_bytes[phase] += bytes;
        X.t_last = t;
    }
}

/* Barrier that carries a reduction: the CTA partial travels with the arrival
   flag; thread i of every CTA picks up partial i as soon as flag i is up and the
   G partials are combined in index order (deterministic, same in every CTA). */
/* The two halves of the all-reduce, so that work which does not depend on its result can run between them:
   post = CTA partial into the slot + arrival flag (release: everything this CTA wrote so far is ordered before it),
   collect = wait for all G flags (acquire), combine the G partials in index order.  No other barrier or
   all-reduce may be started in between (the slots rotate on the shared sequence number). */
template <class Comb, bool HL>
__device__ void eng_allreduce_post(EngCtxT<HL> &X, const EngArgs &A, Key v, const Key &none, Comb comb, Key *wres)
{
    v = block_reduce(v, none, comb);      /* its __syncthreads also fence this CTA's phase writes */
    X.seq++;
    if (X.G == 1) {
        if (X.tid == 0) wres[0] = v;
        return;
    }
    EngSlot *ring = A.slots + (X.seq & (ENG_RING - 1)) * ENG_MAXG;
    if (X.tid == 0) {
        Key *d = &ring[X.cta].key;
        __stcg(&d->a, v.a); __stcg(&d->b, v.b); __stcg(&d->c, v.c); __stcg(&d->pos, v.pos); __stcg(&d->aux, v.aux);
        eng_st_release(&ring[X.cta].flag, X.seq);
    }
}
template <class Comb, bool HL>
__device__ Key eng_allreduce_collect(EngCtxT<HL> &X, const EngArgs &A, const Key &none, Comb comb, Key *wres)
{
    if (X.G == 1) {
        __syncthreads();
        Key r = wres[0];
        __syncthreads();
        return r;
    }
    EngSlot *ring = A.slots + (X.seq & (ENG_RING - 1)) * ENG_MAXG;
    const int nw = (X.G + 31) >> 5;
    if (X.warp < nw) {
        Key r = none;
        if (X.tid < X.G) {
            const unsigned int *f = &ring[X.tid].flag;
            while (eng_ld_relaxed(f) < X.seq) { }
            eng_fence_acq();
            const Key *sl = &ring[X.tid].key;
            r.a = __ldcg(&sl->a); r.b = __ldcg(&sl->b); r.c = __ldcg(&sl->c);
            r.pos = __ldcg(&sl->pos); r.aux = __ldcg(&sl->aux);
        }
        r = warp_reduce(r, none, comb);
        if (X.lane == 0) wres[X.warp] = r;
    }
    __syncthreads();
    Key r = wres[0];
    for (int w = 1; w < nw; w++) comb(r, wres[w]);
    __syncthreads();
    return r;
}
template <class Comb, bool HL>
__device__ Key eng_allreduce(EngCtxT<HL> &X, const EngArgs &A, Key v, const Key &none, Comb comb)
{
    __shared__ Key wres[ENG_MAXG / 32];
    eng_allreduce_post(X, A, v, none, comb, wres);
    return eng_allreduce_collect(X, A, none, comb, wres);
}

/* lanes per item for the segmented sums below: as many as the grid offers for
   nitems in ONE round (rounds of the multi-warp form are separated by
   __syncthreads), but no more than entries, and at least 4 */
template <bool HL>
__device__ __forceinline__ int eng_pick_lp(const EngCtxT<HL> &X, int nitems, double avg_len)
{
    int LP = 4;
    while (LP < 1024 && (long)nitems * LP * 2 <= (long)X.gsize && LP < avg_len) LP <<= 1;
    return LP;
}

/* Segmented sums over the grid: item i is summed by LP lanes (a power of two,
   4..1024).  LP <= 32: sub-warp groups and shuffles; LP > 32: LP/32 warps of one
   CTA, partial sums meet in shared memory in a fixed order.  body(i, l, LP, a)
   accumulates lane l's share into a[0..NV); out(i, a) runs in one thread.
   All threads of the grid must call with identical nitems and LP. */
template <int NV, class Body, class Out, bool HL>
__device__ __forceinline__ void eng_items(const EngCtxT<HL> &X, int nitems, int LP, Body body, Out out)
{
    __shared__ double part[NV][32];
    if (LP <= 32) {
        const int ngroups = X.gsize / LP, g = X.gtid / LP, l = X.tid & (LP - 1);
        for (int i0 = 0; i0 < nitems; i0 += ngroups) {
            const int item = i0 + g;
            double a[NV];
#pragma unroll
            for (int x = 0; x < NV; x++) a[x] = 0.0;
            if (item < nitems) body(item, l, LP, a);
            for (int off = LP >> 1; off > 0; off >>= 1) {
#pragma unroll
                for (int x = 0; x < NV; x++) a[x] += __shfl_xor_sync(FULLMASK, a[x], off);
            }
            if (item < nitems && l == 0) out(item, a);
        }
    } else {
        const int W = LP >> 5, IPC = 32 / W;
        const int sub = X.warp / W, prt = X.warp % W;
        for (int i0 = 0; i0 < nitems; i0 += X.G * IPC) {
            const int item = i0 + X.cta * IPC + sub;
            double a[NV];
#pragma unroll
            for (int x = 0; x < NV; x++) a[x] = 0.0;
            if (item < nitems) body(item, prt * 32 + X.lane, LP, a);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                for (int x = 0; x < NV; x++) a[x] += __shfl_xor_sync(FULLMASK, a[x], off);
            }
            if (X.lane == 0) {
#pragma unroll
                for (int x = 0; x < NV; x++) part[x][X.warp] = a[x];
            }
            __syncthreads();
            if (prt == 0 && X.lane == 0 && item < nitems) {
#pragma unroll
                for (int x = 0; x < NV; x++) {
                    double s = 0.0;
                    for (int w = 0; w < W; w++) s += part[x][sub * W + w];
                    a[x] = s;
                }
                out(item, a);
            }
            __syncthreads();
        }
    }
}

/* lane l's share of  sum_ptr val[ptr] * x[ind[ptr]]  over [beg, end); the
   matrix arrays are immutable (read-only path), x is not */
__device__ __forceinline__ double eng_spdot(const int *ind, const double *val, int beg, int end, int l, int LP,
                                            const double *x)
{
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int ptr = beg + l;
    for (; ptr + 3 * LP < end; ptr += 4 * LP) {
        const int c0 = __ldg(ind + ptr), c1 = __ldg(ind + ptr + LP), c2 = __ldg(ind + ptr + 2 * LP),
                  c3 = __ldg(ind + ptr + 3 * LP);
        const double v0 = __ldg(val + ptr), v1 = __ldg(val + ptr + LP), v2 = __ldg(val + ptr + 2 * LP),
                     v3 = __ldg(val + ptr + 3 * LP);
        a0 += v0 * x[c0]; a1 += v1 * x[c1]; a2 += v2 * x[c2]; a3 += v3 * x[c3];
    }
    for (; ptr < end; ptr += LP) a0 += __ldg(val + ptr) * x[__ldg(ind + ptr)];
    return (a0 + a1) + (a2 + a3);
}

/* sum_{j<n} term(j).  Unrolling by eight (ENG_SUM8: the loads of eight terms issued before the first product is
   consumed) was measured 2 % SLOWER on C3: the extra live registers cost the dense stream more than the
   shorter latency chains of these short loops gain (same-box A/B, profiles/r02s_ab.txt). */
template <class Term>
__device__ __forceinline__ double eng_sum8(int n, Term term)
{
#ifndef ENG_SUM8
    double s0 = 0.0;
    for (int jj = 0; jj < n; jj++) s0 += term(jj);
    return s0;
#endif
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int j = 0;
    for (; j + 7 < n; j += 8) {
        const double t0 = term(j), t1 = term(j + 1), t2 = term(j + 2), t3 = term(j + 3),
                     t4 = term(j + 4), t5 = term(j + 5), t6 = term(j + 6), t7 = term(j + 7);
        a0 += t0; a1 += t1; a2 += t2; a3 += t3; a0 += t4; a1 += t5; a2 += t6; a3 += t7;
    }
    if (j + 3 < n) {
        const double t0 = term(j), t1 = term(j + 1), t2 = term(j + 2), t3 = term(j + 3);
        a0 += t0; a1 += t1; a2 += t2; a3 += t3;
        j += 4;
    }
    for (; j < n; j++) a0 += term(j);
    return (a0 + a1) + (a2 + a3);
}

/* ordered compaction inside a CTA: returns the position of this thread's
   element among the flagged ones and the total; all threads must call */
template <bool HL>
__device__ __forceinline__ int eng_compact(const EngCtxT<HL> &X, bool flag, int &total)
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

/* ---- the dense stream through cp.async (LDGSTS): the 16-byte pieces of the next ENG_AD columns of a lane travel
   from global to ITS OWN slots of a shared-memory ring without passing through registers, so the bytes in
   flight per lane are bounded by shared memory (ENG_AD x 16) instead of by the three 16-byte loads the register
   allocation of the 1024-thread kernel leaves room for.  Every lane reads back only what it copied itself:
   no barrier, only cp.async.wait_group. ---- */
#ifndef ENG_AD
#define ENG_AD 6
#endif
__device__ __forceinline__ void eng_cp_async16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned int)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void eng_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void eng_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

/* one lane's share: columns e0, e0 + 32, ... < L of the two rows at Tb2; slot s of the lane at ring + s * sstride */
__device__ __forceinline__ void eng_dense_stream_async(const double *Tb2, size_t ldt, int e0, int L, const double *val,
                                                       double *ring, int sstride, double &ax, double &ay)
{
    const int cnt = (L - e0 + 31) >> 5;
    const size_t step = 32 * ldt;
    const double *src = Tb2 + (size_t)e0 * ldt;
#pragma unroll
    for (int s = 0; s < ENG_AD; s++) {
        if (s < cnt) eng_cp_async16(ring + s * sstride, src + s * step);
        eng_cp_async_commit();
    }
    const double *nxt = src + ENG_AD * step;
    const double *v = val + e0;
    double a0 = 0.0, b0 = 0.0, a1 = 0.0, b1 = 0.0;
    int i = 0;
    for (; i + ENG_AD <= cnt; i += ENG_AD) {
#pragma unroll
        for (int s = 0; s < ENG_AD; s++) {
            eng_cp_async_wait<ENG_AD - 1>();
            const double2 t = *(const double2 *)(ring + s * sstride);
            const double vv = v[(i + s) * 32];
            if (s & 1) { a1 += t.x * vv; b1 += t.y * vv; } else { a0 += t.x * vv; b0 += t.y * vv; }
            if (i + s + ENG_AD < cnt) eng_cp_async16(ring + s * sstride, nxt + (size_t)(i + s) * step);
            eng_cp_async_commit();
        }
    }
#pragma unroll
    for (int s = 0; s < ENG_AD; s++) {
        if (i + s < cnt) {
            eng_cp_async_wait<ENG_AD - 1>();
            const double2 t = *(const double2 *)(ring + s * sstride);
            const double vv = v[(i + s) * 32];
            a0 += t.x * vv; b0 += t.y * vv;
        }
        eng_cp_async_commit();
    }
    eng_cp_async_wait<0>();
    ax = a0 + a1; ay = b0 + b1;
}

/* y[b] (+)= sum_e T[b, idx[e]] * val[e] + sum_{j<nd} Fd_j[b] z[j], b < k, for the
   list (idx, val) of L entries (idx == NULL: idx[e] = e, the dense case).
   Every CTA owns one contiguous range of rows (a multiple of 4, i.e. whole 32-byte
   sectors) and walks it in chunks of RB rows; its 32 warps split the list and
   the partial sums meet in shared memory in a fixed order.  Streams 8 L k bytes. */
template <bool HL, class Pre>
__device__ void eng_gemv_rows(const EngCtxT<HL> &X, const EngArgs &A, int k, int L, const int *idx,
                              const double *val, double *y, double *ycol, bool accumulate,
                              int nd, const double *z, Pre pre)
{
    /* pre(): called ONCE by every thread of every CTA, after the first chunk has been streamed and before
       anything reads z (or at the end, for a CTA without rows): the place where a pending all-reduce is collected */
    bool pre_done = false;
    __shared__ double red[32][33];
    const int RPC = (((k + X.G - 1) / X.G) + 3) & ~3;          /* rows per CTA */
    const int q0 = min(k, X.cta * RPC), q1 = min(k, q0 + RPC);
    const size_t ldt = (size_t)A.ldt;
    int RB = 32;
    for (int b0 = q0; b0 < q1; b0 += RB) {
        /* chunk height by what is left: a short last chunk puts two or four columns in one warp load */
        const int rem = q1 - b0;
        RB = (rem >= 24) ? 32 : (rem >= 12 ? 16 : 8);
        if (idx == nullptr && rem >= 24) {
            /* dense stream: 16-byte loads, two rows per lane, eight loads in flight per
               thread.  Up to 32 rows left: half a warp per column, two columns per warp
               instruction; 33..64 rows: the whole chunk in ONE pass, a warp per column
               (each column of T is touched once, in one contiguous piece). */
#ifndef ENG_WIDE_MIN
#define ENG_WIDE_MIN 32
#endif
            const bool wide = rem > ENG_WIDE_MIN;
            const int CH = wide ? min(rem, 64) : 32;             /* rows of this chunk */
            const int r2 = wide ? X.lane : (X.lane & 15), sub2 = wide ? 0 : (X.lane >> 4);
            const int cstep = wide ? 32 : 64, cfirst = wide ? X.warp : sub2 + 2 * X.warp;
            const int ba = b0 + 2 * r2;
            const int lim = min(b0 + CH, q1);
            const bool in0 = ba < lim, in1 = ba + 1 < lim;
            const double *Tb2 = A.T + (in0 ? ba : q0);
            double ax0 = 0.0, ay0 = 0.0, ax1 = 0.0, ay1 = 0.0;
            int e = cfirst;
            /* ring of the cp.async variant: behind the staged vector in the dynamic shared memory, one slot row of
               nact 16-byte pieces per (warp, stage) */
            const int nact = (lim - b0 + 1) >> 1;
            const int kpad = (L + 1) & ~1;
            const bool use_async = wide && A.use_async && val == X.sh_d &&
                                   (size_t)(A.dcap - kpad) * 8 >= (size_t)32 * ENG_AD * nact * 16;
            if (use_async) {
                if (in0) {
                    double *ring = X.sh_d + kpad + ((size_t)X.warp * ENG_AD * nact + r2) * 2;
                    eng_dense_stream_async(Tb2, ldt, cfirst, L, val, ring, nact * 2, ax0, ay0);
                }
            } else
            if (in0) {
                /* Under the 64-register cap of a 1024-thread CTA ptxas keeps only three of the eight loads of
                   a trip in flight (SASS: the destinations R16/R20/R24 are reused), which leaves the stream
                   short of memory-level parallelism.  The sectors are therefore requested from DRAM a little
                   ahead of the loads with prefetch.global.L2 -- no register, no shared memory: every even lane
                   asks for the 32-byte sector it shares with its neighbour, pf_dist column steps ahead
                   (measured on C3: distance 1-2 steps = 32-64 columns, 3-7 MB in flight over the chip, -3.6 %
                   solve time; 8 trips ahead and more thrash L2 and cost 8 %; profiles/r02j_pf_sweep.txt). */
                const size_t pfo = (size_t)A.pf_dist * cstep * ldt;
                const bool pfl = A.pf_dist > 0 && !(X.lane & 1);
                for (; e + (ENG_SU - 1) * cstep < L; e += ENG_SU * cstep) {
                    if (pfl && e + (A.pf_dist + ENG_SU) * cstep <= L) {
#pragma unroll
                        for (int x = 0; x < ENG_SU; x++)
                            if (x >= A.pf_first)
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(Tb2 + (size_t)(e + x * cstep) * ldt + pfo));
                    }
                    double2 t[ENG_SU];
#pragma unroll
                    for (int x = 0; x < ENG_SU; x++) t[x] = __ldcg((const double2 *)(Tb2 + (size_t)(e + x * cstep) * ldt));
#pragma unroll
                    for (int x = 0; x < ENG_SU; x += 2) {
                        const double v0 = val[e + x * cstep], v1 = val[e + (x + 1) * cstep];
                        ax0 += t[x].x * v0; ay0 += t[x].y * v0;
                        ax1 += t[x + 1].x * v1; ay1 += t[x + 1].y * v1;
                    }
                }
                for (; e < L; e += cstep) {
                    const double2 t0 = __ldcg((const double2 *)(Tb2 + (size_t)e * ldt));
                    const double v0 = val[e];
                    ax0 += t0.x * v0; ay0 += t0.y * v0;
                }
            }
            double ax = in0 ? ax0 + ax1 : 0.0, ay = in1 ? ay0 + ay1 : 0.0;
            if (!wide) {
                ax += __shfl_xor_sync(FULLMASK, ax, 16);
                ay += __shfl_xor_sync(FULLMASK, ay, 16);
            }
            double (*red2)[65] = (double (*)[65])X.sh_red2;      /* [32][65] */
            if (sub2 == 0) { red2[X.warp][2 * r2] = ax; red2[X.warp][2 * r2 + 1] = ay; }
            __syncthreads();
            if (!pre_done) { pre(); pre_done = true; }
            if (X.tid < CH) {
                double s = 0.0;
#pragma unroll 8
                for (int w = 0; w < 32; w++) s += red2[w][X.tid];
                const int b2 = b0 + X.tid;
                if (b2 < q1) {
                    s += eng_sum8(nd, [&](int j) { return A.Fd[(size_t)j * ldt + b2] * z[j]; });
                    if (accumulate) s += y[b2];
                    y[b2] = s;
                    ycol[XHDR(head)[XHDR(slot_pos)[b2]] - A.m] = s;
                }
            }
            __syncthreads();
            RB = CH;
            continue;
        }
        const int NSUB = 32 / RB;
        const int r = X.lane & (RB - 1), sub = X.lane / RB;
        const int stride = NSUB * 32;
        const int b = b0 + r;
        const bool inb = b < q1;
        const double *Tb = A.T + (inb ? b : q0);
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
        if (!pre_done) { pre(); pre_done = true; }
        if (X.tid < RB) {
            double s = 0.0;
#pragma unroll 8
            for (int w = 0; w < 32; w++) s += red[w][X.tid];
            const int b2 = b0 + X.tid;
            if (b2 < q1) {
                s += eng_sum8(nd, [&](int j) { return A.Fd[(size_t)j * ldt + b2] * z[j]; });
                if (accumulate) s += y[b2];
                y[b2] = s;
                ycol[XHDR(head)[XHDR(slot_pos)[b2]] - A.m] = s;      /* the same value by basic column */
            }
        }
        __syncthreads();
    }
    if (!pre_done) pre();
}

/* ---- TMA (bulk asynchronous copy) + mbarrier helpers ---- */
__device__ __forceinline__ unsigned int eng_smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void eng_mbar_init(unsigned long long *bar, unsigned int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(eng_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void eng_mbar_inval(unsigned long long *bar)
{
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(eng_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void eng_mbar_expect_tx(unsigned long long *bar, unsigned int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(eng_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void eng_mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(eng_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void eng_mbar_wait(unsigned long long *bar, unsigned int parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(eng_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void eng_bulk_g2s(void *dst, const void *src, unsigned int bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(eng_smem_u32(dst)), "l"(src), "r"(bytes), "r"(eng_smem_u32(bar)) : "memory");
}

/* The dense stream y = T v (8 k^2 bytes, the dominant pass of large dual
   problems) with TMA/shared-memory staging of column segments: every CTA owns a
   contiguous range of <= 64 rows; the piece of each column of T that falls in
   it (one contiguous run of 16-byte multiples) is brought into shared memory by
   cp.async.bulk, ENG_TNC columns per stage, ENG_TST stages in flight, completion
   on an mbarrier per stage -- the bytes in flight no longer depend on registers.
   The 1024 threads consume a stage as 64 row slots x 16 column groups; a second
   mbarrier per stage (one arrival per warp) hands the slot back to the producer
   (warp 0). */
#define ENG_TNC 64
#define ENG_TST 4
#define ENG_TMA_SMEM (ENG_TST * ENG_TNC * 64)       /* doubles */
template <bool HL>
__device__ __forceinline__ bool eng_gemv_tma_ok(const EngCtxT<HL> &X, const EngArgs &A, int k)
{
    const int RPC = (((k + X.G - 1) / X.G) + 3) & ~3;
    return RPC > 16 && RPC <= 64 && A.dcap >= ENG_TMA_SMEM && k >= 4 * ENG_TNC;
}
template <bool HL>
__device__ void eng_gemv_dense_tma(const EngCtxT<HL> &X, const EngArgs &A, int k, const double *v, double *y, double *ycol,
                                   int nd, const double *z)
{
    __shared__ __align__(8) unsigned long long full_bar[ENG_TST], empty_bar[ENG_TST];
    const size_t ldt = (size_t)A.ldt;
    const int RPC = (((k + X.G - 1) / X.G) + 3) & ~3;
    const int q0 = min(k, X.cta * RPC), q1 = min(k, q0 + RPC);
    const int rows = q1 - q0;
    if (rows <= 0) return;                                    /* uniform per CTA: no grid barrier inside */
    const int rowsp = (rows + 1) & ~1;                         /* 16-byte multiples */
    const unsigned int rowbytes = (unsigned int)rowsp * 8u;
    double *stage = X.sh_d;                                    /* [ENG_TST][ENG_TNC][rowsp] */
    const int ntile = (k + ENG_TNC - 1) / ENG_TNC;
    if (X.tid == 0) {
        for (int s = 0; s < ENG_TST; s++) { eng_mbar_init(&full_bar[s], 1u); eng_mbar_init(&empty_bar[s], ENG_NT / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const double *Tq = A.T + q0;
    /* producer: warp 0 fills one stage */
    auto fill = [&](int tile) {
        const int s = tile % ENG_TST;
        const int c0 = tile * ENG_TNC, nc = min(ENG_TNC, k - c0);
        if (X.lane == 0) eng_mbar_expect_tx(&full_bar[s], (unsigned int)nc * rowbytes);
        __syncwarp();
        double *dst = stage + (size_t)s * ENG_TNC * rowsp;
        for (int c = X.lane; c < nc; c += 32)
            eng_bulk_g2s(dst + (size_t)c * rowsp, Tq + (size_t)(c0 + c) * ldt, rowbytes, &full_bar[s]);
    };
    if (X.warp == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int t = 0; t < min(ENG_TST, ntile); t++) fill(t);
    }
    const int r = X.tid & 63, cg = X.tid >> 6;                 /* 64 row slots x 16 column groups */
    double acc0 = 0.0, acc1 = 0.0;
    for (int tile = 0; tile < ntile; tile++) {
        const int s = tile % ENG_TST;
        const unsigned int par = (unsigned int)((tile / ENG_TST) & 1);
        eng_mbar_wait(&full_bar[s], par);
        const int c0 = tile * ENG_TNC, nc = min(ENG_TNC, k - c0);
        const double *src = stage + (size_t)s * ENG_TNC * rowsp + r;
        if (r < rows) {
            int c = cg;
            for (; c + 16 < nc; c += 32) {
                acc0 += src[(size_t)c * rowsp] * v[c0 + c];
                acc1 += src[(size_t)(c + 16) * rowsp] * v[c0 + c + 16];
            }
            if (c < nc) acc0 += src[(size_t)c * rowsp] * v[c0 + c];
        }
        __syncwarp();
        if (X.lane == 0) eng_mbar_arrive(&empty_bar[s]);       /* this warp is done with the slot */
        if (X.warp == 0 && tile + ENG_TST < ntile) {
            eng_mbar_wait(&empty_bar[s], par);                 /* every warp is done with it */
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            fill(tile + ENG_TST);
        }
    }
    /* column groups meet in shared memory (fixed order), then the epilogue of eng_gemv_rows */
    double (*red2)[65] = (double (*)[65])X.sh_red2;            /* [32][65]: 16 groups used */
    red2[cg][r] = acc0 + acc1;
    __syncthreads();
    if (X.tid < rows) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < 16; w++) sum += red2[w][X.tid];
        const int b2 = q0 + X.tid;
        sum += eng_sum8(nd, [&](int j) { return A.Fd[(size_t)j * ldt + b2] * z[j]; });
        y[b2] = sum;
        ycol[XHDR(head)[XHDR(slot_pos)[b2]] - A.m] = sum;
    }
    __syncthreads();
    if (X.tid == 0)
        for (int s = 0; s < ENG_TST; s++) { eng_mbar_inval(&full_bar[s]); eng_mbar_inval(&empty_bar[s]); }
    __syncthreads();
}

/* FTRAN, first half, for the right-hand side h = -N_q (eval_tcol,
   lib/glpspx01.js:690-727): y = T h_N over the entries of column q that fall
   on rows of R_N.  CTA 0 also scatters h into the dense vector hz that the
   second half reads. */
template <bool HL>
__device__ void eng_ftran_head_col(EngCtxT<HL> &X, const EngArgs &A, int k, int kq, double *y, double *ycol, int nd)
{
    __shared__ double zs[ENG_DB];
    const int m = A.m;
    const int RPC = (((k + X.G - 1) / X.G) + 3) & ~3;
    const bool work = (X.cta * RPC < k);
    if (!work && X.cta != 0) return;
    if (kq < m) {
        if (X.tid == 0) {
            X.sh_i[0] = XHDR(cslot)[kq]; X.sh_d[0] = -1.0;
            if (X.cta == 0) A.hz[kq] = -1.0;
        }
        __syncthreads();
        if (X.tid < nd) zs[X.tid] = -A.Rd[(size_t)X.tid * A.ldt + X.sh_i[0]];
        __syncthreads();
        eng_gemv_rows(X, A, k, 1, X.sh_i, X.sh_d, y, ycol, false, nd, zs, [] {});
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
                r = __ldg(A.a_ind + e); a = __ldg(A.a_val + e); cs = XHDR(cslot)[r];
                if (X.cta == 0) A.hz[r] = a;
            }
            int tot;
            const int pos = eng_compact(X, valid && cs >= 0, tot);
            if (valid && cs >= 0) { X.sh_i[L + pos] = cs; X.sh_d[L + pos] = a; }
            L += tot;
        }
        __syncthreads();
        if (X.tid < nd) {
            const double *Rj = A.Rd + (size_t)X.tid * A.ldt;
            double zz = 0.0;
            zz = eng_sum8(L, [&](int e) { return X.sh_d[e] * Rj[X.sh_i[e]]; });
            zs[X.tid] = zz;
        }
        __syncthreads();
        eng_gemv_rows(X, A, k, L, X.sh_i, X.sh_d, y, ycol, !first, nd, zs, [] {});
        first = false;
    }
}

/* FTRAN, second half: x[i] for every basic position from y = T h_N, given by
   basic column (ycol, zero on non-basic columns): a basic auxiliary variable
   gathers its row of A, a basic structural one reads its own entry.
   PREP (primal): the reductions of k_primal_prep ride along. */
template <bool PREP, bool HL>
__device__ __forceinline__ void eng_ftran_tail(const EngCtxT<HL> &X, const EngArgs &A, const Ctrl &S,
                                               const double *h, const double *ycol, double *x, Key &acc)
{
    const int m = A.m;
    const bool pse = PREP && gamma_on(&S);
    const int LP = eng_pick_lp(X, m, A.avg_row);
    eng_items<1>(X, m, LP,
        [&](int i, int l, int lp, double *a) {
            const int kk = XHDR(head)[i];
            if (kk < m) a[0] = eng_spdot(A.at_ind, A.at_val, __ldg(A.at_ptr + kk), __ldg(A.at_ptr + kk + 1), l, lp, ycol);
        },
        [&](int i, const double *a) {
            const int kk = XHDR(head)[i];
            const double t = (kk < m) ? h[kk] + a[0] : ycol[kk - m];
            x[i] = t;
            if (PREP) {
                acc.a += A.coef[kk] * t;
                acc.c = fmax(acc.c, fabs(t));
                if (pse) {
                    const double vv = A.refsp[kk] ? t : 0.0;
                    A.v[i] = vv;
                    if (kk < m) A.vrow[kk] = vv;
                    acc.b += vv * vv;
                }
            }
        });
}

/* two FTRAN tails in one pass over the rows of A (dual engine: tcol from (hz, ycol) and the tail of
   u = inv(B) v from (v, ycol2)): the row's indices and values are loaded once, the dependent chain
   head -> at_ptr -> (at_ind, at_val) -> gather is walked once instead of twice */
template <bool HL>
__device__ __forceinline__ void eng_ftran_tail2(const EngCtxT<HL> &X, const EngArgs &A,
                                                const double *h1, const double *ycol1, double *x1,
                                                const double *h2, const double *ycol2, double *x2)
{
    const int m = A.m;
    const int LP = eng_pick_lp(X, m, A.avg_row);
    eng_items<2>(X, m, LP,
        [&](int i, int l, int lp, double *a) {
            const int kk = XHDR(head)[i];
            if (kk >= m) return;
            const int beg = __ldg(A.at_ptr + kk), end = __ldg(A.at_ptr + kk + 1);
            double p0 = 0.0, p1 = 0.0, q0 = 0.0, q1 = 0.0;
            int ptr = beg + l;
            for (; ptr + lp < end; ptr += 2 * lp) {
                const int c0 = __ldg(A.at_ind + ptr), c1 = __ldg(A.at_ind + ptr + lp);
                const double v0 = __ldg(A.at_val + ptr), v1 = __ldg(A.at_val + ptr + lp);
                const double y0 = ycol1[c0], y1 = ycol1[c1], z0 = ycol2[c0], z1 = ycol2[c1];
                p0 += v0 * y0; p1 += v1 * y1; q0 += v0 * z0; q1 += v1 * z1;
            }
            if (ptr < end) {
                const int c0 = __ldg(A.at_ind + ptr);
                const double v0 = __ldg(A.at_val + ptr);
                p0 += v0 * ycol1[c0]; q0 += v0 * ycol2[c0];
            }
            a[0] = p0 + p1; a[1] = q0 + q1;
        },
        [&](int i, const double *a) {
            const int kk = XHDR(head)[i];
            if (kk < m) { x1[i] = h1[kk] + a[0]; x2[i] = h2[kk] + a[1]; }
            else { x1[i] = ycol1[kk - m]; x2[i] = ycol2[kk - m]; }
        });
}

/* BTRAN, first half: w[b] = c[pos_b] + sum_{r in R_B} A[r, j_b] c[bind[r]],
   with c given by position (v) and by row of a basic auxiliary (vrow) */
template <bool HL>
__device__ __forceinline__ void eng_btran_head(const EngCtxT<HL> &X, const EngArgs &A, int k)
{
    const int m = A.m;
    const int LP = eng_pick_lp(X, k, A.avg_col);
    eng_items<1>(X, k, LP,
        [&](int b, int l, int lp, double *a) {
            const int j = XHDR(head)[XHDR(slot_pos)[b]] - m;
            a[0] = eng_spdot(A.a_ind, A.a_val, __ldg(A.a_ptr + j), __ldg(A.a_ptr + j + 1), l, lp, A.vrow);
        },
        [&](int b, const double *a) { A.wk[b] = A.v[XHDR(slot_pos)[b]] + a[0]; });
}

/* zn[cs] = sum_b T[b, cs] w[b]: column dots, coalesced.  8 k^2 bytes. */
template <bool HL>
__device__ __forceinline__ void eng_gemvT(const EngCtxT<HL> &X, const EngArgs &A, int k, const double *w, double *zn)
{
    const int LP = max(32, eng_pick_lp(X, k, (double)k / 2));
    eng_items<1>(X, k, LP,
        [&](int cs, int l, int lp, double *a) {
            const double *col = A.T + (size_t)cs * A.ldt;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int b = l;
            for (; b + 3 * lp < k; b += 4 * lp) {
                const double t0 = __ldcg(col + b), t1 = __ldcg(col + b + lp);
                const double t2 = __ldcg(col + b + 2 * lp), t3 = __ldcg(col + b + 3 * lp);
                a0 += t0 * w[b]; a1 += t1 * w[b + lp]; a2 += t2 * w[b + 2 * lp]; a3 += t3 * w[b + 3 * lp];
            }
            for (; b < k; b += lp) a0 += __ldcg(col + b) * w[b];
            a[0] = (a0 + a1) + (a2 + a3);
        },
        [&](int cs, const double *a) { zn[cs] = a[0]; });
}

/* rho = row p of inv(B) (eval_rho, lib/glpspx01.js:1030-1042), read out of T
   (plus the nd deferred rank-1 terms) */
template <bool HL>
__device__ void eng_rho(EngCtxT<HL> &X, const EngArgs &A, int k, int p, int nd)
{
    __shared__ double fs[ENG_DB];
    const int m = A.m;
    const size_t ldt = (size_t)A.ldt;
    /* head[p] and rslot[p] are fetched together, and the deferred factors Fd_j[bp] are read (as broadcasts)
       next to the Rd_j[cs] they multiply: two levels of dependent loads instead of four, no staging barrier */
    const int kp = XHDR(head)[p];
    const int bp0 = XHDR(rslot)[p];
    for (int r = X.vtid; r < m; r += X.gsize) {
        const int c = XHDR(cslot)[r], b = XHDR(bind)[r];
        if (c < 0) A.rho[r] = (b == p) ? 1.0 : 0.0;
    }
    if (kp >= m) {
        const int bp = bp0;
        const double *row = A.T + bp;
        const double *Fp = A.Fd + bp;
        for (int cs = X.vtid; cs < k; cs += X.gsize) {
            const double t0 = __ldcg(row + (size_t)cs * ldt);
            const int r = XHDR(slot_row)[cs];
            const double *Rc = A.Rd + cs;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int j = 0;
            for (; j + 3 < nd; j += 4) {
                const double r0 = Rc[(size_t)j * ldt], r1 = Rc[(size_t)(j + 1) * ldt], r2 = Rc[(size_t)(j + 2) * ldt],
                             r3 = Rc[(size_t)(j + 3) * ldt];
                const double f0 = Fp[(size_t)j * ldt], f1 = Fp[(size_t)(j + 1) * ldt], f2 = Fp[(size_t)(j + 2) * ldt],
                             f3 = Fp[(size_t)(j + 3) * ldt];
                a0 += f0 * r0; a1 += f1 * r1; a2 += f2 * r2; a3 += f3 * r3;
            }
            for (; j < nd; j++) a0 += Fp[(size_t)j * ldt] * Rc[(size_t)j * ldt];
            A.rho[r] = t0 + ((a0 + a1) + (a2 + a3));
        }
        return;
    }
    /* an auxiliary variable leaves: rho_N = w' T with w the basic part of row kp of A */
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
            if (valid) { pb = XHDR(bind)[m + __ldg(A.at_ind + e)]; a = __ldg(A.at_val + e); }
            int tot;
            const int pos = eng_compact(X, valid && pb < m, tot);
            if (valid && pb < m) { X.sh_i[L + pos] = XHDR(rslot)[pb]; X.sh_d[L + pos] = a; }
            L += tot;
        }
        __syncthreads();
        if (X.tid < nd) {
            const double *Fj = A.Fd + (size_t)X.tid * ldt;
            double ww = 0.0;
            ww = eng_sum8(L, [&](int e) { return X.sh_d[e] * Fj[X.sh_i[e]]; });
            fs[X.tid] = ww;
        }
        if (nd > 0) __syncthreads();
        const int LP = eng_pick_lp(X, k, (double)L);
        const bool acc = !first;
        eng_items<1>(X, k, LP,
            [&](int cs, int l, int lp, double *a) {
                const double *col = A.T + (size_t)cs * ldt;
                double a0 = 0.0, a1 = 0.0;
                int e = l;
                for (; e + lp < L; e += 2 * lp) {
                    const double t0 = __ldcg(col + X.sh_i[e]), t1 = __ldcg(col + X.sh_i[e + lp]);
                    a0 += t0 * X.sh_d[e]; a1 += t1 * X.sh_d[e + lp];
                }
                if (e < L) a0 += __ldcg(col + X.sh_i[e]) * X.sh_d[e];
                a[0] = a0 + a1;
            },
            [&](int cs, const double *a) {
                const int r = XHDR(slot_row)[cs];
                double v = a[0];
                v += eng_sum8(nd, [&](int j) { return fs[j] * A.Rd[(size_t)j * ldt + cs]; });
                A.rho[r] = acc ? A.rho[r] + v : v;
            });
        first = false;
        __syncthreads();
    }
}

/* pivot row (eval_trow): trow[j] = -rho' N_j for the non-basic non-fixed
   columns; with PSE also s_j = N_j' u (primal update_gamma) from the same pass.
   Dual: acc.c collects |trow|_inf, acc.a the sum of trow_j^2 over the reference
   space, and trowcol gets the reference-space part of trow by column. */
template <bool DUAL, bool HL>
__device__ __forceinline__ void eng_trow(const EngCtxT<HL> &X, const EngArgs &A, const Ctrl &S, Key &acc)
{
    const int m = A.m, n = A.n;
    const bool pse = gamma_on(&S);
    const bool want_s = !DUAL && pse;
    const int LP = eng_pick_lp(X, n, A.avg_col);
    eng_items<2>(X, n, LP,
        [&](int j, int l, int lp, double *a) {
            if (XHDR(stat)[j] == GLP_NS) return;
            const int k = XHDR(head)[m + j];
            if (k < m) {
                if (l == 0) { a[0] = -A.rho[k]; if (want_s) a[1] = A.u[k]; }
                return;
            }
            const int beg = __ldg(A.a_ptr + (k - m)), end = __ldg(A.a_ptr + (k - m) + 1);
            double t0 = 0.0, t1 = 0.0, s0 = 0.0, s1 = 0.0;
            int ptr = beg + l;
            for (; ptr + 3 * lp < end; ptr += 4 * lp) {
                const int r0 = __ldg(A.a_ind + ptr), r1 = __ldg(A.a_ind + ptr + lp), r2 = __ldg(A.a_ind + ptr + 2 * lp),
                          r3 = __ldg(A.a_ind + ptr + 3 * lp);
                const double v0 = __ldg(A.a_val + ptr), v1 = __ldg(A.a_val + ptr + lp),
                             v2 = __ldg(A.a_val + ptr + 2 * lp), v3 = __ldg(A.a_val + ptr + 3 * lp);
                const double x0 = A.rho[r0], x1 = A.rho[r1], x2 = A.rho[r2], x3 = A.rho[r3];
                t0 += x0 * v0; t1 += x1 * v1; t0 += x2 * v2; t1 += x3 * v3;
                if (want_s) {
                    const double u0 = A.u[r0], u1 = A.u[r1], u2 = A.u[r2], u3 = A.u[r3];
                    s0 -= v0 * u0; s1 -= v1 * u1; s0 -= v2 * u2; s1 -= v3 * u3;
                }
            }
            for (; ptr < end; ptr += lp) {
                const int r0 = __ldg(A.a_ind + ptr);
                const double v0 = __ldg(A.a_val + ptr);
                t0 += A.rho[r0] * v0;
                if (want_s) s0 -= v0 * A.u[r0];
            }
            a[0] = t0 + t1; a[1] = s0 + s1;
        },
        [&](int j, const double *a) {
            const double t = a[0];
            A.trow[j] = t;
            if (want_s) A.svec[j] = a[1];
            if (DUAL) {
                acc.c = fmax(acc.c, fabs(t));
                const int k = XHDR(head)[m + j];
                const bool in = pse && t != 0.0 && A.refsp[k];
                if (in) acc.a += t * t;
                if (k >= m) A.trowcol[k - m] = in ? t : 0.0;
            }
        });
}

/* dual update_gamma, first half (lib/glpspx02.js:1103-1132), by rows:
   v[r] = sum_{j in C, non-basic} N_j[r] trow_j; the entries on rows of R_N are
   also written in kernel order (wk) for the dense product with T */
template <bool HL>
__device__ __forceinline__ void eng_gamma_rhs(const EngCtxT<HL> &X, const EngArgs &A)
{
    const int m = A.m;
    const int LP = eng_pick_lp(X, m, A.avg_row);
    eng_items<1>(X, m, LP,
        [&](int row, int l, int lp, double *a) {
            a[0] = -eng_spdot(A.at_ind, A.at_val, __ldg(A.at_ptr + row), __ldg(A.at_ptr + row + 1), l, lp, A.trowcol);
        },
        [&](int row, const double *a) {
            const int pr = XHDR(bind)[row];
            const double val = a[0] + ((pr >= m && A.refsp[row]) ? A.trow[pr - m] : 0.0);
            A.v[row] = val;
            const int cs = XHDR(cslot)[row];
            if (cs >= 0) A.wk[cs] = val;
        });
}

/* description of a basis change, identical in every CTA */
struct EngChange {
    int p, q, kp, kq, LS, ES, csq, bp, k, knew, ctgt, bnew;
    double tp;
};

template <bool HL>
__device__ __forceinline__ void eng_describe_change(const EngCtxT<HL> &X, const EngArgs &A, const Ctrl &S, EngChange &C)
{
    const int m = A.m;
    C.p = S.p; C.q = S.q; C.k = S.k;
    C.kp = XHDR(head)[C.p]; C.kq = XHDR(head)[m + C.q];
    C.LS = (C.kp < m); C.ES = (C.kq < m);
    C.tp = A.tcol[C.p];
    C.csq = C.ES ? XHDR(cslot)[C.kq] : -1;
    C.bp = C.LS ? -1 : XHDR(rslot)[C.p];
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
template <bool HL>
__device__ void eng_update_T(const EngCtxT<HL> &X, const EngArgs &A, const EngChange &C)
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
        const int i = XHDR(slot_pos)[sb];
        const bool isp = (i == C.p);
        const double f = isp ? -1.0 / tp : A.tcol[i] / tp;
#pragma unroll
        for (int x = 0; x < ENG_UC / 8; x++) {
            const int cs = c0 + cg + 8 * x;
            if (cs >= kd) continue;
            double *dst = A.T + (size_t)cs * ldt + b;
            if (C.LS && C.ES && cs == C.csq) { *dst = -A.tcol[i] / tp; continue; }
            const int scs = (removal && cs == C.csq) ? k - 1 : cs;
            const double rs = A.rho[XHDR(slot_row)[scs]];
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
        for (int t = X.vtid; t < 2 * k + 1; t += X.gsize) {
            if (t < k) A.T[(size_t)k * ldt + t] = -A.tcol[XHDR(slot_pos)[t]] / tp;
            else if (t < 2 * k) A.T[(size_t)(t - k) * ldt + k] = -A.rho[XHDR(slot_row)[t - k]] / tp;
            else A.T[(size_t)k * ldt + k] = -1.0 / tp;
        }
    }
}

/* The basis change in deferred form: instead of rewriting T (16 k^2 bytes) the
   rank-1 term (g, rho_N) is appended to (Fd, Rd); only the O(k) structural part
   touches T: a replaced row or column is written directly (and its entries in
   the earlier terms are zeroed), a leaving row/column is filled from the last
   one, in T and in every stored term alike. */
template <bool HL>
__device__ void eng_defer_apply(const EngCtxT<HL> &X, const EngArgs &A, const EngChange &C, int nd)
{
    const int k = C.k;
    const size_t ldt = (size_t)A.ldt;
    const double tp = C.tp;
    const bool removal = (!C.LS && C.ES);
    double *Fn = A.Fd + (size_t)nd * ldt, *Rn = A.Rd + (size_t)nd * ldt;
    /* entries that the structural part below rewrites are left to it (one writer per cell) */
    const int skipF = removal ? C.bp : -1;
    const int skipR = (removal || (C.LS && C.ES)) ? C.csq : -1;
    for (int t = X.vtid; t < k; t += X.gsize) {
        const int i = XHDR(slot_pos)[t];
        if (t != skipF) Fn[t] = (i == C.p) ? 0.0 : -A.tcol[i] / tp;    /* row of p: replaced below */
        if (t != skipR) Rn[t] = A.rho[XHDR(slot_row)[t]];
    }
    if (C.LS && C.ES) {
        /* column csq is replaced: T[:, csq] = -tcol_S / tp */
        for (int t = X.vtid; t < k; t += X.gsize) A.T[(size_t)C.csq * ldt + t] = -A.tcol[XHDR(slot_pos)[t]] / tp;
        for (int j = X.vtid; j <= nd; j += X.gsize) A.Rd[(size_t)j * ldt + C.csq] = 0.0;
    } else if (C.LS && !C.ES) {
        /* a row and a column join at slot k */
        for (int t = X.vtid; t < 2 * k + 1; t += X.gsize) {
            if (t < k) A.T[(size_t)k * ldt + t] = -A.tcol[XHDR(slot_pos)[t]] / tp;
            else if (t < 2 * k) A.T[(size_t)(t - k) * ldt + k] = -A.rho[XHDR(slot_row)[t - k]] / tp;
            else A.T[(size_t)k * ldt + k] = -1.0 / tp;
        }
        for (int j = X.vtid; j <= nd; j += X.gsize) { A.Fd[(size_t)j * ldt + k] = 0.0; A.Rd[(size_t)j * ldt + k] = 0.0; }
    } else if (removal) {
        /* row bp and column csq leave: the last row/column moves into the hole */
        for (int t = X.vtid; t < 2 * k; t += X.gsize) {
            if (t < k) {            /* destination (bp, c = t), c < k - 1 */
                const int c = t;
                if (C.bp != k - 1 && c < k - 1) {
                    const int sc = (c == C.csq) ? k - 1 : c;
                    A.T[(size_t)c * ldt + C.bp] = __ldcg(A.T + (size_t)sc * ldt + (k - 1));
                }
            } else {                /* destination (s = t - k, csq), s < k - 1, s != bp */
                const int sr = t - k;
                if (C.csq != k - 1 && sr < k - 1 && sr != C.bp)
                    A.T[(size_t)C.csq * ldt + sr] = __ldcg(A.T + (size_t)(k - 1) * ldt + sr);
            }
        }
        for (int j = X.vtid; j <= nd; j += X.gsize) {
            if (C.bp != k - 1) A.Fd[(size_t)j * ldt + C.bp] = (j == nd) ? -A.tcol[XHDR(slot_pos)[k - 1]] / tp
                                                                       : A.Fd[(size_t)j * ldt + (k - 1)];
            if (C.csq != k - 1) A.Rd[(size_t)j * ldt + C.csq] = (j == nd) ? A.rho[XHDR(slot_row)[k - 1]]
                                                                         : A.Rd[(size_t)j * ldt + (k - 1)];
        }
    } else {
        /* a structural variable replaces a structural one: row bp = -rho_N / tp */
        for (int c = X.vtid; c < k; c += X.gsize) A.T[(size_t)c * ldt + C.bp] = -A.rho[XHDR(slot_row)[c]] / tp;
        for (int j = X.vtid; j < nd; j += X.gsize) A.Fd[(size_t)j * ldt + C.bp] = 0.0;
    }
}

/* fp64 tensor-core tile product: D(8x8) = A(8x4) B(4x8) + C.  Fragment layout of
   mma.m8n8k4.f64: a = A[lane/4][lane%4], b = B[lane%4][lane/4],
   c0,c1 = C[lane/4][2*(lane%4) + {0,1}]. */
__device__ __forceinline__ void eng_dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

/* T += sum_{j<nd} Fd_j Rd_j' over the k x k kernel: one read+write of T for up
   to ENG_DB basis changes -- the one dense contraction of the iteration loop, so
   it runs on the fp64 tensor cores (DMMA m8n8k4).  A unit is a block of 64
   columns times a chunk of rows: the 32 x 64 slice of Rd is staged in shared
   memory (row stride 68: bank-conflict-free fragment loads), then every warp
   walks its own 8-row blocks with no block-wide barrier, the Fd fragments coming
   straight from L2, so that the DRAM latency of one warp's tile overlaps the
   arithmetic of the others. */
#define ENG_RS 68
#define ENG_FLUSH_SMEM (ENG_DB * ENG_RS)
template <bool HL>
__device__ void eng_flush(const EngCtxT<HL> &X, const EngArgs &A, int k, int nd)
{
    if (nd <= 0 || k <= 0) return;
    const size_t ldt = (size_t)A.ldt;
    double *Rs = X.sh_d;                    /* [ENG_DB][ENG_RS]: Rs[j][c] = Rd_j[c0 + c] */
    const int g = X.lane >> 2, t = X.lane & 3;
    const int rbk = X.warp & 15, cq = X.warp >> 4;           /* 8-row block of a 128-row step, column half */
    const int ncb = (k + 63) >> 6;
    int nrc = (3 * X.G + ncb - 1) / ncb;                      /* row chunks: about three units per CTA */
    nrc = max(1, min(nrc, (k + 127) >> 7));
    const int RCH = ((((k + nrc - 1) / nrc) + 127) >> 7) << 7;
    const int ksteps = (nd + 3) >> 2;
    for (int u = X.cta; u < ncb * nrc; u += X.G) {
        const int c0 = (u % ncb) << 6, r0 = (u / ncb) * RCH, r1 = min(k, r0 + RCH);
        __syncthreads();
        for (int e = X.tid; e < ENG_DB * 64; e += ENG_NT) {
            const int j = e >> 6, c = e & 63, cc = c0 + c;
            Rs[j * ENG_RS + c] = (j < nd && cc < k) ? A.Rd[(size_t)j * ldt + cc] : 0.0;
        }
        __syncthreads();
        for (int i0 = r0; i0 < r1; i0 += 128) {
            const int row = i0 + rbk * 8 + g;
            const bool rok = row < r1;
            double acc[4][2];
            double *tp[4];
#pragma unroll
            for (int x = 0; x < 4; x++) {
                const int col = c0 + (cq * 4 + x) * 8 + 2 * t;
                tp[x] = A.T + (size_t)col * ldt + row;
                acc[x][0] = (rok && col < k) ? __ldcg(tp[x]) : 0.0;
                acc[x][1] = (rok && col + 1 < k) ? __ldcg(tp[x] + ldt) : 0.0;
            }
            /* the deferred terms in batches of 32 (eight k-steps of the m8n8k4 tile product) */
#pragma unroll
            for (int kb = 0; kb < ENG_DB / 4; kb += 8) {
                if (kb < ksteps) {
                    double af[8];
#pragma unroll
                    for (int kk = 0; kk < 8; kk++)
                        af[kk] = (4 * (kb + kk) + t < nd && rok) ? __ldcg(A.Fd + (size_t)(4 * (kb + kk) + t) * ldt + row) : 0.0;
#pragma unroll
                    for (int kk = 0; kk < 8; kk++) {
                        if (kb + kk < ksteps) {
#pragma unroll
                            for (int x = 0; x < 4; x++) {
                                const double b = Rs[(4 * (kb + kk) + t) * ENG_RS + (cq * 4 + x) * 8 + g];
                                eng_dmma(acc[x][0], acc[x][1], af[kk], b);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int x = 0; x < 4; x++) {
                const int col = c0 + (cq * 4 + x) * 8 + 2 * t;
                if (rok && col < k) tp[x][0] = acc[x][0];
                if (rok && col + 1 < k) tp[x][ldt] = acc[x][1];
            }
        }
    }
    __syncthreads();
}

/* slot maps, basis header and the new non-basic status (change_basis,
   lib/glpspx01.js:1310-1371 / lib/glpspx02.js:1259-1294) on one copy of the
   header arrays; one thread */
__device__ void eng_bookkeep_hdr(int *head, int *bind, int *rslot, int *slot_pos, int *cslot, int *slot_row,
                                 signed char *stat, int m, const EngChange &C, int new_stat)
{
    const int k = C.k;
    if (C.LS) { cslot[C.kp] = C.ctgt; slot_row[C.ctgt] = C.kp; }
    if (C.ES) {
        cslot[C.kq] = -1;
        if (!C.LS && C.csq != k - 1) { int rl = slot_row[k - 1]; slot_row[C.csq] = rl; cslot[rl] = C.csq; }
    }
    if (C.bnew >= 0) { rslot[C.p] = C.bnew; slot_pos[C.bnew] = C.p; }
    if (!C.LS && C.ES) {
        rslot[C.p] = -1;
        if (C.bp != k - 1) { int pl = slot_pos[k - 1]; slot_pos[C.bp] = pl; rslot[pl] = C.bp; }
    }
    head[C.p] = C.kq; head[m + C.q] = C.kp;
    bind[C.kq] = C.p; bind[C.kp] = m + C.q;
    stat[C.q] = (signed char)new_stat;
}

/* the whole O(1) remainder of a basis change: thread 0 of every CTA updates its
   private header (if any), thread 0 of CTA 0 the arrays in global memory and the
   vectors whose zero pattern follows the basis */
template <bool HL>
__device__ void eng_bookkeep(const EngCtxT<HL> &X, const EngArgs &A, const EngChange &C, int new_stat, bool drop_refsp)
{
    const int m = A.m;
    if (HL)
        eng_bookkeep_hdr(X.head, X.bind, X.rslot, X.slot_pos, X.cslot, X.slot_row, X.stat, m, C, new_stat);
    if (X.cta != 0) return;
    eng_bookkeep_hdr(A.head, A.bind, A.rslot, A.slot_pos, A.cslot, A.slot_row, A.stat, m, C, new_stat);
    if (drop_refsp) A.refsp[C.kp] = 0;
    /* zero where the variable left the set the copy is indexed over */
    if (C.LS) A.vrow[C.kp] = 0.0;
    else { A.ycol[C.kp - m] = 0.0; A.ycol2[C.kp - m] = 0.0; }
    if (!C.ES) A.trowcol[C.kq - m] = 0.0;
}

template <bool HL>
__device__ __forceinline__ void eng_init(EngCtxT<HL> &X, const EngArgs &A, double *dyn)
{
    X.G = gridDim.x; X.cta = blockIdx.x; X.tid = threadIdx.x;
    X.lane = X.tid & 31; X.warp = X.tid >> 5;
    /* vtid: grid-wide thread index with groups of four warps interleaved over the CTAs (virtual group =
       (tid / 128) * G + cta): the first N entries of a grid-stride loop land on threads 0..127 of every CTA,
       then on threads 128..255, ... instead of filling CTA 0, 1, 2, ... to the brim.  The vectors of the small
       phases (m, n, k entries against 150 k threads) are gathers bound by one SM's load/store path: on C3 the
       update loops ran on 16 of 148 SMs.  128 consecutive entries stay together so that the warps of a group
       still share the 128-byte lines of the index arrays in L1 (interleaving single warps cost more in L2
       requests than it gained: profiles/r02m_interleave.txt). */
    X.gtid = X.cta * ENG_NT + X.tid; X.gsize = X.G * ENG_NT;
    X.vtid = ((((X.tid >> 7) * X.G + X.cta) << 7) | (X.tid & 127));
    X.seq = 0u;
    X.dseq = 0u;
    X.t_last = clock64();
    X.sh_d = dyn;
    X.sh_i = (int *)(dyn + A.dcap);
    X.sh_red2 = (double *)(X.sh_i + ENG_LCAP);
    X.hdr_local = HL ? 1 : 0;
    if (HL) {
        const int m = A.m, n = A.n, ldt = A.ldt;
        int *base = (int *)(X.sh_red2 + 32 * 65);
        X.head = base; X.bind = X.head + (m + n); X.rslot = X.bind + (m + n); X.cslot = X.rslot + m;
        X.slot_pos = X.cslot + m; X.slot_row = X.slot_pos + ldt;
        X.stat = (signed char *)(X.slot_row + ldt);
        for (int t = X.tid; t < m + n; t += ENG_NT) { X.head[t] = A.head[t]; X.bind[t] = A.bind[t]; }
        for (int t = X.tid; t < m; t += ENG_NT) { X.rslot[t] = A.rslot[t]; X.cslot[t] = A.cslot[t]; }
        for (int t = X.tid; t < ldt; t += ENG_NT) { X.slot_pos[t] = A.slot_pos[t]; X.slot_row[t] = A.slot_row[t]; }
        for (int t = X.tid; t < n; t += ENG_NT) X.stat[t] = A.stat[t];
        __syncthreads();
    }
}

/* block-wide reduction whose result every thread of the CTA gets */
template <class Comb, bool HL>
__device__ Key eng_blockall(const EngCtxT<HL> &X, Key v, const Key &none, Comb comb)
{
    __shared__ Key res;
    v = block_reduce(v, none, comb);
    if (X.tid == 0) res = v;
    __syncthreads();
    Key r = res;
    __syncthreads();
    return r;
}

/* vectors up to this length are scanned by EVERY CTA on its own (replicated
   ratio test: two block-wide reductions, no grid barrier); longer ones by the
   grid with two barrier-reductions */
#define ENG_LOCAL_MAX 8192

enum { /* phase slots of the cycle accounting (12 per engine) */
    PP_PRICE0 = 0, PP_A, PP_B, PP_R1, PP_R2, PP_C, PP_D, PP_E, PP_F, PP_B_TAIL, PP_B_REDUCE, PP_B_BTRAN,
    PD_PRICE0 = 0, PD_RHO, PD_TROW, PD_R1, PD_R2, PD_X1, PD_TCOL1, PD_TCOL2, PD_UPD, PD_FLUSH
};

/* ------------------------------------------------------------------ */
/* primal engine: lib/glpspx01.js:1868-2056                           */
/* ------------------------------------------------------------------ */
template <bool HL>
__global__ void __launch_bounds__(ENG_NT, 1) k_engine_primal(EngArgs A)
{
    extern __shared__ __align__(128) double eng_dyn[];
    __shared__ Ctrl S;
    __shared__ EngChange C;
    EngCtxT<HL> X;
    eng_init(X, A, eng_dyn);
    const int m = A.m, n = A.n;
    const bool local_ratio = (m <= A.local_max);
    const double nnzA = (double)__ldg(A.a_ptr + n);
    if (X.tid == 0) S = *A.ctrl;
    __syncthreads();
    int qnext = P_NONE;
    for (int it = 0; it < A.max_iters && S.status == ST_OK; it++) {
        /* ---- pricing (chuzc): the first iteration of a launch prices on its own,
                later ones were priced by the update phase of their predecessor ---- */
        if (it == 0) {
            Key none = {0.0, 0.0, 0.0, INT_MAX, 0};
            Key v = none;
            scan_chuzc_primal(v, X.vtid, X.gsize, n, XHDR(stat), A.cbar, A.gamma, A.tol_dj, -1, 0);
            Key r = eng_allreduce(X, A, v, none, CombArgMax());
            qnext = (r.a > 0.0 && r.pos != INT_MAX) ? r.pos : P_NONE;
            eng_mark(X, A, PP_PRICE0, 17.0 * n);
        }
        if (X.tid == 0) {
            S.q = qnext;
            S.big = 0.0;
            if (qnext == P_NONE) S.status = ST_NONE1;
        }
        __syncthreads();
        if (S.status != ST_OK) break;
        const int q = S.q;
        const int kq = XHDR(head)[m + q];
        const double nnz_q = (kq < m) ? 1.0 : (double)(__ldg(A.a_ptr + (kq - m) + 1) - __ldg(A.a_ptr + (kq - m)));
        /* ---- A: tcol, first half ---- */
        eng_ftran_head_col(X, A, S.k, kq, A.yk, A.ycol, 0);
        eng_bar(X, A);
        eng_mark(X, A, PP_A, 12.0 * nnz_q + 8.0 * S.k * nnz_q * ((double)S.k / m) + 8.0 * S.k);
        /* ---- B: tcol, second half + the reductions of k_primal_prep ---- */
        {
            Key none = {0.0, 0.0, 0.0, 0, 0};
            Key acc = none;
            eng_ftran_tail<true>(X, A, S, A.hz, A.ycol, A.tcol, acc);
            eng_mark(X, A, PP_B_TAIL, 12.0 * (double)__ldg(A.at_ptr + m) * (1.0 - (double)S.k / m) + 29.0 * m);
            Key r = eng_allreduce(X, A, acc, none, CombSum2());
            eng_mark(X, A, PP_B_REDUCE, 0.0);
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
            if (S.status != ST_OK) { eng_mark(X, A, PP_B, 0.0); break; }
        }
        const double sgn = (S.d1 > 0.0 ? -1.0 : +1.0);
        const bool pse = gamma_on(&S);
        bool cbar_q_pending = true;
        bool pend_bar = false;        /* arrived at the barrier after the replicated ratio test, not waited yet */    /* reeval_cost result cbar[q] = d1 (lib/glpspx01.js:1915-1918) not stored yet */
        /* ---- Harris ratio test (chuzr); the first half of u = inv(B') v rides along ---- */
        if (local_ratio) {
            if (pse) eng_btran_head(X, A, S.k);
            eng_mark(X, A, PP_B_BTRAN, pse ? 12.0 * nnzA * ((double)S.k / n) + 16.0 * S.k : 0.0);
            Key none = {DBL_MAX, 0.0, 0.0, INT_MAX, 0};
            Key v = none;
            if (X.tid == 0 && A.type[kq] == GLP_DB) {
                v.a = __dsub_rn(A.ub[kq], A.lb[kq]); v.b = 1.0; v.pos = -1; v.aux = 0;
            }
            if (m <= 2 * ENG_NT) {
                /* both passes from ONE scan of memory: every thread keeps its (at most two)
                   candidates in registers, exact and relaxed ratio alike */
                RatioCand cand[2];
                bool ok[2];
#pragma unroll
                for (int x = 0; x < 2; x++) {
                    const int pos = X.tid + x * ENG_NT;
                    ok[x] = (pos < m) && ratio_primal_elem(cand[x], pos, S.phase, sgn, S.eps, A.rtol, A.type, A.lb, A.ub,
                                                           A.coef, XHDR(head), A.bbar, A.tcol);
                    if (ok[x]) { const Key c = ratio_primal_key(cand[x], 1, pos); CombRatio1()(v, c); }
                }
                Key r = eng_blockall(X, v, none, CombRatio1());
                if (X.tid == 0) fin_ratio_primal(&S, r, 1, sgn, A.rtol, nullptr, A.tie_stop != 0);
                __syncthreads();
                if (S.status == ST_OK && !S.skip2) {
                    const double tmax = S.tmax;
                    v = none;
#pragma unroll
                    for (int x = 0; x < 2; x++)
                        if (ok[x] && cand[x].t2 <= tmax) {
                            const Key c = ratio_primal_key(cand[x], 2, X.tid + x * ENG_NT);
                            CombRatio2()(v, c);
                        }
                    r = eng_blockall(X, v, none, CombRatio2());
                    if (X.tid == 0) fin_ratio_primal(&S, r, 2, sgn, A.rtol, nullptr, A.tie_stop != 0);
                    __syncthreads();
                }
            } else {
                scan_ratio_primal(v, X.tid, ENG_NT, 1, S.phase, sgn, S.eps, 0.0, A.rtol, A.type, A.lb, A.ub,
                                  A.coef, XHDR(head), A.bbar, A.tcol, nullptr, m);
                Key r = eng_blockall(X, v, none, CombRatio1());
                if (X.tid == 0) fin_ratio_primal(&S, r, 1, sgn, A.rtol, nullptr, A.tie_stop != 0);
                __syncthreads();
                if (S.status == ST_OK && !S.skip2) {
                    v = none;
                    scan_ratio_primal(v, X.tid, ENG_NT, 2, S.phase, sgn, S.eps, S.tmax, A.rtol, A.type, A.lb, A.ub,
                                      A.coef, XHDR(head), A.bbar, A.tcol, nullptr, m);
                    r = eng_blockall(X, v, none, CombRatio2());
                    if (X.tid == 0) fin_ratio_primal(&S, r, 2, sgn, A.rtol, nullptr, A.tie_stop != 0);
                    __syncthreads();
                }
            }
            eng_mark(X, A, PP_R1, 45.0 * m * (S.skip2 ? 1.0 : 2.0));
            /* one barrier: w (first half of u) is complete, and nobody still scans bbar/tcol
               when the update phase of a bound flip rewrites bbar.  Only the ARRIVAL happens here: rho, which
               needs p (known to every CTA from its own copy of the ratio test) but not w, is computed before
               the wait, so the barrier's round trip hides behind it. */
            eng_arrive(X, A);
            pend_bar = true;
        } else {
            eng_mark(X, A, PP_B, 0.0);
            Key none = {DBL_MAX, 0.0, 0.0, INT_MAX, 0};
            Key v = none;
            if (X.vtid == 0 && A.type[kq] == GLP_DB) {
                v.a = __dsub_rn(A.ub[kq], A.lb[kq]); v.b = 1.0; v.pos = -1; v.aux = 0;
            }
            scan_ratio_primal(v, X.vtid, X.gsize, 1, S.phase, sgn, S.eps, 0.0, A.rtol, A.type, A.lb, A.ub,
                              A.coef, XHDR(head), A.bbar, A.tcol, nullptr, m);
            if (pse) eng_btran_head(X, A, S.k);
            Key r = eng_allreduce(X, A, v, none, CombRatio1());
            if (X.cta == 0 && X.tid == 0) A.cbar[q] = S.d1;     /* every CTA has read the old value by now */
            cbar_q_pending = false;
            if (X.tid == 0) fin_ratio_primal(&S, r, 1, sgn, A.rtol, nullptr, A.tie_stop != 0);
            __syncthreads();
            eng_mark(X, A, PP_R1, 45.0 * m);
            if (S.status == ST_OK && !S.skip2) {
                v = none;
                scan_ratio_primal(v, X.vtid, X.gsize, 2, S.phase, sgn, S.eps, S.tmax, A.rtol, A.type, A.lb, A.ub,
                                  A.coef, XHDR(head), A.bbar, A.tcol, nullptr, m);
                r = eng_allreduce(X, A, v, none, CombRatio2());
                if (X.tid == 0) fin_ratio_primal(&S, r, 2, sgn, A.rtol, nullptr, A.tie_stop != 0);
                __syncthreads();
                eng_mark(X, A, PP_R2, 45.0 * m);
            }
        }
        if (S.status != ST_OK) {
            /* a slower CTA may still be reading cbar[q] in its own copy of the d1/d2 test:
               the value stored here makes that test come out the same way */
            if (cbar_q_pending && X.cta == 0 && X.tid == 0) A.cbar[q] = S.d1;
            if (pend_bar) { eng_wait(X, A); pend_bar = false; }
            break;
        }
        const int p = S.p;
        if (cbar_q_pending && X.cta == 0 && X.tid == 0) A.cbar[q] = S.d1;   /* see above: harmless for a CTA still at the d1/d2 test */
        if (p < 0 && pend_bar) { eng_wait(X, A); pend_bar = false; eng_mark(X, A, PP_B, 0.0); }
        if (p >= 0) {
            /* ---- C: rho and the second half of u ---- */
            eng_rho(X, A, S.k, p, 0);
            if (pend_bar) { eng_wait(X, A); pend_bar = false; eng_mark(X, A, PP_B, 0.0); }
            if (pse) {
                const int k = S.k;
                const int LP = max(32, eng_pick_lp(X, k, (double)k / 2));
                eng_items<1>(X, k, LP,
                    [&](int cs, int l, int lp, double *a) {
                        const double *col = A.T + (size_t)cs * A.ldt;
                        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                        int b = l;
                        for (; b + 3 * lp < k; b += 4 * lp) {
                            const double t0 = __ldcg(col + b), t1 = __ldcg(col + b + lp);
                            const double t2 = __ldcg(col + b + 2 * lp), t3 = __ldcg(col + b + 3 * lp);
                            a0 += t0 * A.wk[b]; a1 += t1 * A.wk[b + lp]; a2 += t2 * A.wk[b + 2 * lp]; a3 += t3 * A.wk[b + 3 * lp];
                        }
                        for (; b < k; b += lp) a0 += __ldcg(col + b) * A.wk[b];
                        a[0] = (a0 + a1) + (a2 + a3);
                    },
                    [&](int cs, const double *a) { A.u[XHDR(slot_row)[cs]] = a[0]; });
                for (int r = X.vtid; r < m; r += X.gsize)
                    if (XHDR(cslot)[r] < 0) A.u[r] = A.vrow[r];
            }
            eng_bar(X, A);
            eng_mark(X, A, PP_C, 12.0 * m + 8.0 * S.k + (pse ? 8.0 * S.k * (double)S.k + 20.0 * m : 0.0));
            /* ---- E: pivot row and PSE inner products ---- */
            {
                Key dummy = {0.0, 0.0, 0.0, 0, 0};
                eng_trow<false>(X, A, S, dummy);
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
                    if (S.status == ST_OK) eng_describe_change(X, A, S, C);
                }
                __syncthreads();
                eng_mark(X, A, PP_E, 12.0 * nnzA * (1.0 - (double)S.k / n) + 13.0 * n + 8.0 * m);
                if (S.status != ST_OK) break;
            }
        }
        /* ---- F: update_bbar / update_cbar / update_gamma / basis change, and the
                pricing of the next iteration from the values just written ---- */
        {
            const double teta = S.teta;
            const double xq = get_xN(XHDR(stat), XHDR(head), A.lb, A.ub, m, q);
            int new_stat;
            if (p >= 0) new_stat = S.p_stat;
            else new_stat = (XHDR(stat)[q] == GLP_NL) ? GLP_NU : GLP_NL;
            const double new_dq = S.new_dq;
            const double pivot = (p >= 0) ? A.trow[q] : 1.0;
            const int kp = (p >= 0) ? C.kp : 0;
            const int phase = S.phase;
            Key pnone = {0.0, 0.0, 0.0, INT_MAX, 0};
            Key pv = pnone;
            for (int t = X.vtid; t < n; t += X.gsize) {
                double dj, g;
                int st;
                if (p < 0) {
                    dj = A.cbar[t]; g = A.gamma[t]; st = XHDR(stat)[t];
                    if (t == q) { dj = S.d1; A.cbar[q] = dj; st = new_stat; }     /* reeval_cost result */
                }
                else if (t == q) {
                    dj = new_dq;
                    if (phase == 1) dj -= A.coef[kp];
                    A.cbar[q] = dj;
                    g = A.gamma[q];
                    if (pse) {
                        g = 1.0;
                        if (A.type[kp] != GLP_FX) {
                            g = S.gamma_q / (pivot * pivot);
                            if (g < DBL_EPSILON) g = DBL_EPSILON;
                        }
                        A.gamma[q] = g;
                    }
                    st = new_stat;
                } else {
                    const double tr = A.trow[t];
                    dj = A.cbar[t]; g = A.gamma[t]; st = XHDR(stat)[t];
                    if (tr != 0.0) {
                        dj -= tr * new_dq;
                        A.cbar[t] = dj;
                        if (pse) {
                            const double tt = tr / pivot;
                            const int k = XHDR(head)[m + t];
                            const double t1 = g + tt * tt * S.gamma_q + 2.0 * tt * A.svec[t];
                            const double t2 = (A.refsp[k] ? 1.0 : 0.0) + S.delta_q * tt * tt;
                            g = (t1 >= t2 ? t1 : t2);
                            if (g < DBL_EPSILON) g = DBL_EPSILON;
                            A.gamma[t] = g;
                        }
                    }
                }
                price_primal(pv, t, st, dj, g, A.tol_dj);
            }
            /* phase 1: does anything still violate its bound after this iteration (check_feas at the top of
               the reference's next iteration, lib/glpspx01.js:1768-1775)?  The count rides on the pricing key. */
            int ninf = 0;
            if (phase != 1) {
                for (int t = X.vtid; t < m; t += X.gsize) {
                    if (t == p) A.bbar[t] = xq + teta;
                    else if (teta != 0.0) {
                        const double tc = A.tcol[t];
                        if (tc != 0.0) A.bbar[t] += tc * teta;
                    }
                }
            } else {
                for (int t = X.vtid; t < m; t += X.gsize) {
                    double b;
                    if (t == p) { b = xq + teta; A.bbar[t] = b; continue; }      /* xN[q] enters with a zero auxiliary cost */
                    b = A.bbar[t];
                    if (teta != 0.0) {
                        const double tc = A.tcol[t];
                        if (tc != 0.0) { b += tc * teta; A.bbar[t] = b; }
                    }
                    const int k = XHDR(head)[t];
                    const double c = A.coef[k];
                    if (c < 0.0) ninf += (b < A.lb[k] - relax(A.tol_bnd, A.lb[k])) ? 1 : 0;
                    else if (c > 0.0) ninf += (b > A.ub[k] + relax(A.tol_bnd, A.ub[k])) ? 1 : 0;
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
            pv.aux = ninf;
            Key r = eng_allreduce(X, A, pv, pnone, CombArgMaxCnt());
            qnext = (r.a > 0.0 && r.pos != INT_MAX) ? r.pos : P_NONE;
            const bool feasible_now = (phase == 1 && r.aux == 0);
            eng_mark(X, A, PP_F, 24.0 * m + 17.0 * n + (p >= 0 ? 40.0 * n + 16.0 * S.k * (double)S.k : 0.0));
            /* the O(1) remainder: every CTA has arrived, nobody reads the old header any more */
            if (X.tid == 0) {
                if (p >= 0) {
                    if (phase == 1 && X.cta == 0) A.coef[C.kp] = 0.0;       /* lib/glpspx01.js:2016-2020 */
                    eng_bookkeep(X, A, C, new_stat, false);
                } else {
                    if (HL) X.stat[q] = (signed char)new_stat;
                    if (X.cta == 0) A.stat[q] = (signed char)new_stat;
                }
            }
            if (X.tid == 0) {
                if (p >= 0) S.k = C.knew;
                iter_end(&S, p >= 0, 0);
                if (feasible_now && S.status == ST_OK) S.status = ST_PHASE;
            }
            eng_header_done(X, A);
        }
    }
    if (X.cta == 0 && X.tid == 0) *A.ctrl = S;
}

/* ------------------------------------------------------------------ */
/* dual engine: lib/glpspx02.js:1780-1966                             */
/* ------------------------------------------------------------------ */
template <bool HL>
__global__ void __launch_bounds__(ENG_NT, 1) k_engine_dual(EngArgs A)
{
    extern __shared__ __align__(128) double eng_dyn[];
    __shared__ Ctrl S;
    __shared__ EngChange C;
    EngCtxT<HL> X;
    eng_init(X, A, eng_dyn);
    const int m = A.m, n = A.n;
    const bool local_ratio = (n <= A.local_max);
    const bool defer = A.defer && !local_ratio;     /* deferred basis changes need the grid-wide ratio phases */
    const double nnzA = (double)__ldg(A.a_ptr + n);
    if (X.tid == 0) S = *A.ctrl;
    __syncthreads();
    int pnext = P_NONE;
    double dnext = 0.0;
    int nd = 0;                   /* deferred rank-1 terms on top of T */
    __shared__ Key wres2[ENG_MAXG / 32];
    for (int it = 0; it < A.max_iters && S.status == ST_OK; it++) {
        /* ---- pricing (chuzr): see the primal engine ---- */
        if (it == 0) {
            Key none = {0.0, 0.0, 0.0, INT_MAX, 0};
            Key v = none;
            scan_chuzr_dual(v, X.vtid, X.gsize, m, A.type, A.lb, A.ub, XHDR(head), A.bbar, A.gamma, A.tol_bnd, -1, 0);
            Key r = eng_allreduce(X, A, v, none, CombArgMax());
            const bool found = (r.a > 0.0 && r.pos != INT_MAX);
            pnext = found ? r.pos : P_NONE;
            dnext = found ? r.b : 0.0;
            eng_mark(X, A, PD_PRICE0, 37.0 * m);
        }
        if (X.tid == 0) {
            S.p = pnext;
            S.delta = dnext;
            S.big = 0.0;
            if (pnext == P_NONE) S.status = ST_NONE1;
        }
        __syncthreads();
        if (S.status != ST_OK) break;
        const int p = S.p;
        const bool pse = gamma_on(&S);
        const double sgn = (S.delta > 0.0 ? +1.0 : -1.0);
        int pend = 0;                 /* 1: pass 2 of the ratio test posted, 2: barrier arrived -- not collected yet */
        /* ---- rho ---- */
        eng_mark(X, A, 10, 0.0);                  /* header hand-off of the previous iteration + loop top */
        eng_rho(X, A, S.k, p, nd);
        eng_bar(X, A);
        eng_mark(X, A, PD_RHO, 12.0 * m + 8.0 * S.k + 8.0 * nd * S.k);
        /* ---- pivot row, |trow|_inf, sum of squares over the reference space ---- */
        {
            Key none = {0.0, 0.0, 0.0, 0, 0};
            Key acc = none;
            eng_trow<true>(X, A, S, acc);
            const bool flush_now = (nd == ENG_DB);
            if (flush_now) {
                /* nothing else touches T in this phase: fold the deferred terms into it */
                eng_flush(X, A, S.k, nd);
                nd = 0;
            }
            Key r = eng_allreduce(X, A, acc, none, CombSum2());
            if (flush_now) eng_mark(X, A, PD_FLUSH, 16.0 * S.k * (double)S.k);
            if (X.tid == 0) {
                S.trow_max = r.c;
                S.eps = A.tol_bnd * (1.0 + 0.01 * r.c);      /* sic: tol_bnd, lib/glpspx02.js:1851 */
                S.scal = r.a;                                 /* sum of trow_j^2, j in the reference space */
            }
            __syncthreads();
            eng_mark(X, A, PD_TROW, 12.0 * nnzA * (1.0 - (double)S.k / n) + 13.0 * n + 8.0 * m);
        }
        /* ---- Harris ratio test (chuzc) and the right-hand side of update_gamma ---- */
        if (local_ratio) {
            Key none = {DBL_MAX, 0.0, 0.0, INT_MAX, 0};
            Key v = none;
            scan_ratio_dual(v, X.tid, ENG_NT, 1, sgn, S.eps, 0.0, A.rtol, XHDR(stat), A.cbar, A.trow, nullptr, n);
            Key r = eng_blockall(X, v, none, CombRatio1());
            if (X.tid == 0) fin_ratio_dual(&S, r, 1, sgn, A.rtol, nullptr, A.tie_stop != 0);
            __syncthreads();
            if (S.status == ST_OK && !S.skip2) {
                v = none;
                scan_ratio_dual(v, X.tid, ENG_NT, 2, sgn, S.eps, S.tmax, A.rtol, XHDR(stat), A.cbar, A.trow, nullptr, n);
                r = eng_blockall(X, v, none, CombRatio2());
                if (X.tid == 0) fin_ratio_dual(&S, r, 2, sgn, A.rtol, nullptr, A.tie_stop != 0);
                __syncthreads();
            }
            if (S.status != ST_OK) break;
            if (pse) {
                eng_gamma_rhs(X, A);
                eng_bar(X, A);
            }
            eng_mark(X, A, PD_X1, 34.0 * n + (pse ? 17.0 * nnzA + 16.0 * m : 0.0));
        } else {
            Key none = {DBL_MAX, 0.0, 0.0, INT_MAX, 0};
            Key v = none;
            scan_ratio_dual(v, X.vtid, X.gsize, 1, sgn, S.eps, 0.0, A.rtol, XHDR(stat), A.cbar, A.trow, nullptr, n);
            if (pse) eng_gamma_rhs(X, A);
            Key r = eng_allreduce(X, A, v, none, CombRatio1());
            if (X.tid == 0) fin_ratio_dual(&S, r, 1, sgn, A.rtol, nullptr, A.tie_stop != 0);
            __syncthreads();
            eng_mark(X, A, PD_R1, 17.0 * n + (pse ? 17.0 * nnzA + 16.0 * m : 0.0));
            if (S.status != ST_OK) break;
            const bool need_z = pse && nd > 0 && S.k > 0;
            if (need_z) {
                /* z_j = Rd_j' v for the dense product below: one CTA per term */
                for (int j = X.cta; j < nd; j += X.G) {
                    const double *Rj = A.Rd + (size_t)j * A.ldt;
                    const Key z0 = {0.0, 0.0, 0.0, 0, 0};
                    Key zk = z0;
                    for (int c = X.tid; c < S.k; c += ENG_NT) zk.a += Rj[c] * A.wk[c];
                    zk = block_reduce(zk, z0, CombSum2());
                    if (X.tid == 0) A.zbuf[j] = zk.a;
                    __syncthreads();
                }
            }
            /* Pass 2 is only POSTED here (its partial and arrival flag go out); the result is collected after the
               dense product below has streamed T: that product needs v, which is complete, but neither q nor z --
               the all-reduce's round trip and the wait for the slowest CTA hide behind 8 k^2 bytes of HBM traffic. */
            if (!S.skip2) {
                v = none;
                scan_ratio_dual(v, X.vtid, X.gsize, 2, sgn, S.eps, S.tmax, A.rtol, XHDR(stat), A.cbar, A.trow, nullptr, n);
                eng_allreduce_post(X, A, v, none, CombRatio2(), wres2);
                pend = 1;
                eng_mark(X, A, PD_R2, 17.0 * n + (need_z ? 16.0 * nd * S.k : 0.0));
            } else if (need_z) {
                eng_arrive(X, A);
                pend = 2;
                eng_mark(X, A, PD_R2, 16.0 * nd * S.k);
            }
        }
        /* second half of the exchange started above: afterwards q is known and every CTA's z_j is visible */
        auto finish_ratio = [&]() {
            if (pend == 1) {
                const Key none2 = {DBL_MAX, 0.0, 0.0, INT_MAX, 0};
                const Key r2 = eng_allreduce_collect(X, A, none2, CombRatio2(), wres2);
                if (X.tid == 0) fin_ratio_dual(&S, r2, 2, sgn, A.rtol, nullptr, A.tie_stop != 0);
                __syncthreads();
            } else if (pend == 2)
                eng_wait(X, A);
            pend = 0;
        };
        /* ---- the dense product y2 = T v_N of update_gamma, then tcol, first half ---- */
        if (pse && S.k > 0) {
            __shared__ double zd[ENG_DB];
            const int k = S.k;
            const double *val = A.wk;
            const bool tma = A.use_tma && eng_gemv_tma_ok(X, A, k);
            if (!tma && k <= A.dcap) {
                for (int e = X.tid; e < k; e += ENG_NT) X.sh_d[e] = A.wk[e];
                val = X.sh_d;
            }
            if (tma) {
                finish_ratio();
                if (X.tid < nd) zd[X.tid] = __ldcg(A.zbuf + X.tid);
                __syncthreads();
                eng_gemv_dense_tma(X, A, k, A.wk, A.yk2, A.ycol2, nd, zd);
            } else {
                __syncthreads();
                eng_gemv_rows(X, A, k, k, nullptr, val, A.yk2, A.ycol2, false, nd, zd, [&]() {
                    finish_ratio();
                    if (X.tid < nd) zd[X.tid] = __ldcg(A.zbuf + X.tid);
                    __syncthreads();
                });
            }
            __syncthreads();
        } else
            finish_ratio();
        if (S.status != ST_OK) break;
        const int q = S.q;
        const int kq = XHDR(head)[m + q];
        eng_ftran_head_col(X, A, S.k, kq, A.yk, A.ycol, nd);
        eng_bar(X, A);
        eng_mark(X, A, PD_TCOL1, (pse ? 8.0 * S.k * (double)S.k + 16.0 * S.k : 0.0) + 8.0 * S.k);
        /* ---- tcol, second half, and the tail of u = inv(B) v ---- */
        {
            Key dummy = {0.0, 0.0, 0.0, 0, 0};
            if (pse) eng_ftran_tail2(X, A, A.hz, A.ycol, A.tcol, A.v, A.ycol2, A.u);
            else eng_ftran_tail<false>(X, A, S, A.hz, A.ycol, A.tcol, dummy);
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
                    const double eta = (A.refsp[XHDR(head)[p]] ? 1.0 : 0.0);
                    S.delta_q = eta;
                    S.gamma_q = eta + (pse ? S.scal : 0.0);
                    S.teta = S.delta / piv1;
                    if (S.phase == 2) S.obj += (A.cbar[q] / S.zeta) * (S.delta / piv1);
                    eng_describe_change(X, A, S, C);
                }
            }
            __syncthreads();
            eng_mark(X, A, PD_TCOL2, (pse ? 2.0 : 1.0) * (12.0 * nnzA * (1.0 - (double)S.k / m) + 16.0 * m));
            if (S.status != ST_OK) break;
        }
        /* ---- update_cbar / update_bbar / update_gamma / basis change, and the
                pricing of the next iteration from the values just written ---- */
        {
            const double teta = S.teta, new_dq = S.new_dq;
            const int kp = C.kp;
            const double pivot = C.tp;
            const double xq = get_xN(XHDR(stat), XHDR(head), A.lb, A.ub, m, q);
            const bool drop = (A.type[kp] == GLP_FX && A.refsp[kp]);
            /* phase 1: is the basis dual feasible after this iteration (check_feas at the top of the reference's
               next iteration, lib/glpspx02.js:1681-1696)?  The count rides on the pricing key. */
            int ninf = 0;
            Key pnone = {0.0, 0.0, 0.0, INT_MAX, 0};
            Key pv = pnone;
            /* One pass over the reduced costs (n) and the basic values / weights (m): every entry's loads are
               issued before anything depends on them -- first (trow, cbar | gamma, head, bbar, tcol, u), then what
               hangs on head (type, lb, ub, refsp) -- so that the phase costs two rounds of L2 latency instead of
               one per nested condition.  Phase 1 adds the reference's check_feas on the updated reduced costs
               (lib/glpspx02.js:1681-1696); the count rides on the pricing key. */
            const bool ph1 = (S.phase == 1);
            const int nm = max(n, m);
            for (int t = X.vtid; t < nm; t += X.gsize) {
                const bool inn = t < n, inm = t < m;
                double tr = 0.0, dj = 0.0, g = 0.0, bi = 0.0, tc = 0.0, ut = 0.0;
                int k = 0, kk = 0;
                if (inn) { tr = A.trow[t]; dj = A.cbar[t]; if (ph1) kk = XHDR(head)[m + t]; }
                if (inm) { g = A.gamma[t]; k = XHDR(head)[t]; bi = A.bbar[t]; tc = A.tcol[t]; ut = A.u[t]; }
                if (inm && t == p) k = kq;
                int tk = 0, rk = 0, ty = 0;
                double lk = 0.0, uk = 0.0;
                if (inm) { tk = A.type[k]; lk = A.lb[k]; uk = A.ub[k]; rk = A.refsp[k]; }
                if (inn && ph1) ty = A.orig_type[t == q ? kp : kk];
                if (inn) {
                    if (t == q) { dj = new_dq; A.cbar[q] = dj; }
                    else if (new_dq != 0.0 && tr != 0.0) { dj -= tr * new_dq; A.cbar[t] = dj; }
                    if (ph1) {
                        bool bad = false;
                        if (dj < -A.tol_dj) bad = (ty == GLP_LO || ty == GLP_FR);
                        if (dj > +A.tol_dj) bad = bad || (ty == GLP_UP || ty == GLP_FR);
                        ninf += bad ? 1 : 0;
                    }
                }
                if (inm) {
                    if (t == p) {
                        bi = xq + teta;
                        A.bbar[p] = bi;
                        if (pse) {
                            g = 1.0;
                            if (tk != GLP_FR) {
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
                    } else if (tc != 0.0) {
                        if (teta != 0.0) { bi += tc * teta; A.bbar[t] = bi; }
                        if (pse && tk != GLP_FR) {
                            const double tt = tc / pivot;
                            const double t1 = g + tt * tt * S.gamma_q + 2.0 * tt * ut;
                            const double t2 = (rk ? 1.0 : 0.0) + S.delta_q * tt * tt;
                            g = (t1 >= t2 ? t1 : t2);
                            if (g < DBL_EPSILON) g = DBL_EPSILON;
                            if (drop) {
                                g -= tt * tt;
                                if (g < DBL_EPSILON) g = DBL_EPSILON;
                            }
                            A.gamma[t] = g;
                        }
                    }
                    price_dual(pv, t, tk, lk, uk, bi, g, A.tol_bnd);
                }
            }
            if (X.cta == 0) {
                if (kq < m) { if (X.tid == 0) A.hz[kq] = 0.0; }
                else
                    for (int ptr = __ldg(A.a_ptr + (kq - m)) + X.tid; ptr < __ldg(A.a_ptr + (kq - m) + 1); ptr += ENG_NT)
                        A.hz[__ldg(A.a_ind + ptr)] = 0.0;
            }
            if (defer) { eng_defer_apply(X, A, C, nd); nd++; }
            else if (S.k + (C.bnew >= 0) > 0) eng_update_T(X, A, C);
            pv.aux = ninf;
            eng_mark(X, A, 11, 0.0);              /* CTA 0's own share of the update; PD_UPD = all-reduce + header hand-off */
            Key r = eng_allreduce(X, A, pv, pnone, CombArgMaxCnt());
            {
                const bool found = (r.a > 0.0 && r.pos != INT_MAX);
                pnext = found ? r.pos : P_NONE;
                dnext = found ? r.b : 0.0;
            }
            const bool feasible_now = (S.phase == 1 && r.aux == 0);
            eng_mark(X, A, PD_UPD, 24.0 * n + 40.0 * m + 37.0 * m + (defer ? 40.0 * S.k : 16.0 * S.k * (double)S.k));
            const int new_stat = (A.type[kp] == GLP_FX) ? GLP_NS : (S.delta > 0.0 ? GLP_NL : GLP_NU);
            if (X.tid == 0) eng_bookkeep(X, A, C, new_stat, pse && drop);
            if (X.tid == 0) {
                S.k = C.knew;
                iter_end(&S, true, 1);
                if (feasible_now && S.status == ST_OK) S.status = ST_PHASE;
            }
            eng_header_done(X, A);
        }
    }
    /* the host-side kernels read T itself: fold what is still deferred.  Every exit
       from the loop follows a grid-wide barrier after the last access to T. */
    if (nd > 0) eng_flush(X, A, S.k, nd);
    if (X.cta == 0 && X.tid == 0) *A.ctrl = S;
}

#endif /* GLPB_ENGINE_CUH */
