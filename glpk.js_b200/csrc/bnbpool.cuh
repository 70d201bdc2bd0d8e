/* bnbpool.cuh -- host side of the batched branch-and-bound.
 *
 * The host keeps ONLY the tree: for every open node its slab number, local
 * bound, parent's LP objective, level (what ios_driver keeps in IOSNPD,
 * lib/glpios01.js:1-82).  The node states (type/lb/ub/stat of all m+n
 * variables) are device-resident in the node slab and never visit the host
 * except for the incumbent's solution vector.  One round =
 *     pick up to `batch` open nodes (ios_choose_node, lib/glpios12.js)
 *  -> ONE launch of k_bnb_nodes, one CTA per node (nodeengine.cuh)
 *  -> read the verdicts, insert children / delete fathomed nodes
 *     (ios_freeze_node / ios_delete_node / cleanup_the_tree,
 *      lib/glpios01.js:310-570, lib/glpios03.js:472-495, 907-945).
 * Sharding across GPUs: export / import of self-contained node records
 * (device buffers, the payload never touches the host) and an incumbent
 * cut-off; see glpk.js_b200/bnb.py.
 *
 * With -DNE_EMUL the same source compiles for the host (tests/ only).
 */
#ifndef GLPB_BNBPOOL_CUH
#define GLPB_BNBPOOL_CUH

#include "nodeengine.cuh"
#include <algorithm>
#include <cstdlib>
#include <set>
#include <tuple>
#include <vector>

#ifndef NE_EMUL
__global__ void __launch_bounds__(NE_NT, 1)
k_bnb_nodes(NeProb P, const NeTask *tasks, NeResult *res, double *xbuf)
{
    extern __shared__ __align__(16) unsigned char ne_smem[];
    NeT t = {(int)threadIdx.x, (int)blockDim.x, (int)(threadIdx.x & 31), (int)(threadIdx.x >> 5), (int)(blockDim.x >> 5), 32, 0};
    NeS S;
    ne_carve(S, ne_smem, P.m, P.n, P.ldb);
    if (P.a_in_smem) {
        double *As = (double *)(ne_smem + ne_state_bytes(P.m, P.n, P.ldb));
        const double2 *src = (const double2 *)P.As;
        double2 *dst = (double2 *)As;
        const int cnt = P.m * P.lda / 2;
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) dst[e] = src[e];
        S.A = As;
    } else S.A = P.As;
    __syncthreads();
    ne_process_node(t, P, S, tasks[blockIdx.x], res[blockIdx.x], xbuf + (size_t)blockIdx.x * (P.m + P.n));
}
#endif

struct BnbOpen { int slot, level; double bound, lp_obj, up_ii_sum; long seq; };

struct glpb_bnb {
    glpb_prob *P;
    glpb_iocp parm;
    int m, n, mn, batch, cap;
    NeProb np;
    size_t smem_bytes = 0;
    /* device */
    double *d_As = nullptr, *d_cw = nullptr, *d_obj = nullptr, *d_ucoef = nullptr, *d_rii = nullptr, *d_sjj = nullptr;
    signed char *d_kind = nullptr;
    double *d_x = nullptr, *d_inc = nullptr;
    int *d_inc_have = nullptr;
    NeTask *d_tasks = nullptr;
    NeResult *d_res = nullptr;
    std::vector<NeTask> h_tasks;
    std::vector<NeResult> h_res;
    /* tree */
    std::vector<BnbOpen> node;                         /* by slot */
    std::vector<char> alive;
    std::vector<int> free_slots;
    std::set<std::tuple<double, int, long, int>> by_bound;   /* (key, -level, seq, slot) */
    std::set<std::pair<long, int>> by_seq;
    long seq = 0;
    bool have_sol = false, have_cut = false;
    double mip_obj = 0.0;
    std::vector<double> mipx;
    long solved = 0, tasks_done = 0, rounds = 0, iters = 0, refacs = 0;
    long long cyc[6] = {0, 0, 0, 0, 0, 0};
    double tm_beg = 0.0;
    int fail_code = 0;

    double key_of(double bound) const { return P->dir == GLP_MIN ? bound : -bound; }
    int open_count() const { return (int)by_seq.size(); }

    /* ---- backend ---- */
#ifdef NE_EMUL
    static int dalloc(void **p, size_t b) { *p = calloc(1, b ? b : 1); return *p ? 0 : 1; }
    static void dfree(void *p) { free(p); }
    int h2d(void *d, const void *s, size_t b) { memcpy(d, s, b); return 0; }
    int d2h(void *d, const void *s, size_t b) { memcpy(d, s, b); return 0; }
    int d2d(void *d, const void *s, size_t b) { memcpy(d, s, b); return 0; }
    int dsync() { return 0; }
    int launch(int nt)
    {
        std::vector<unsigned char> buf(ne_state_bytes(m, n, np.ldb) + 64);
        for (int b = 0; b < nt; b++) {
            NeT t = {0, 1, 0, 0, 1, 1, 0};
            NeS S;
            ne_carve(S, buf.data(), m, n, np.ldb);
            S.A = np.As;
            ne_process_node(t, np, S, d_tasks[b], d_res[b], d_x + (size_t)b * mn);
        }
        return 0;
    }
#else
    /* Stream-ordered allocation from the device's default pool, which is told to keep what is freed: the node slab
       (17 KB per node, 4.5 GB at the default capacity) is then mapped once per process and handed back and forth
       between successive searches instead of going through cudaMalloc / cudaFree each time (50-300 ms of
       driver time per search, the run-to-run spread of the nodes/s figure). */
    int dalloc(void **p, size_t b)
    {
        static int pool_ready[64] = {0};
        const int dev = P->device;
        if (dev >= 0 && dev < 64 && !pool_ready[dev]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            pool_ready[dev] = 1;
        }
        if (cudaMallocAsync(p, b ? b : 1, P->stream) == cudaSuccess) return 0;
        cudaGetLastError();
        return cudaMalloc(p, b ? b : 1) == cudaSuccess ? 0 : 1;
    }
    void dfree(void *p) { if (p && cudaFreeAsync(p, P->stream) != cudaSuccess) { cudaGetLastError(); cudaFree(p); } }
    int h2d(void *d, const void *s, size_t b) { return cudaMemcpyAsync(d, s, b, cudaMemcpyHostToDevice, P->stream) != cudaSuccess; }
    int d2h(void *d, const void *s, size_t b) { return cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToHost, P->stream) != cudaSuccess; }
    int d2d(void *d, const void *s, size_t b) { return cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToDevice, P->stream) != cudaSuccess; }
    int dsync() { P->n_sync++; return cudaStreamSynchronize(P->stream) != cudaSuccess; }
    int launch(int nt)
    {
        k_bnb_nodes<<<nt, NE_NT, smem_bytes, P->stream>>>(np, d_tasks, d_res, d_x);
        P->n_launch++;
        return cudaGetLastError() != cudaSuccess;
    }
#endif

    ~glpb_bnb()
    {
        dfree(d_As); dfree(d_cw); dfree(d_obj); dfree(d_ucoef); dfree(d_rii); dfree(d_sjj); dfree(d_kind);
        dfree(np.slab_lb); dfree(np.slab_ub); dfree(np.slab_type); dfree(np.slab_stat); dfree(np.slab_bi); dfree(np.slab_upd); dfree(np.slab_head);
        dfree(d_x); dfree(d_inc); dfree(d_inc_have); dfree(d_tasks); dfree(d_res);
    }

    /* problems the one-CTA node engine takes */
    static bool eligible(const glpb_prob *P) { return P->m <= NE_MAXM && P->n <= NE_MAXN; }

    int init(glpb_prob *P_, const glpb_iocp &pr, int batch_, int cap_)
    {
        P = P_; parm = pr; m = P->m; n = P->n; mn = m + n;
        batch = batch_ > 0 ? batch_ : 592;
        cap = cap_ > 0 ? cap_ : 262144;
        if (cap < 2 * batch + 2) cap = 2 * batch + 2;
        memset(&np, 0, sizeof np);
        np.m = m; np.n = n; np.lda = (n + 1) & ~1; np.ldb = m | 1;
        np.dir = P->dir; np.c0 = P->c0;
        /* scaled dense matrix, costs, zeta (lib/glpspx02.js:89-190 init_csa) */
        std::vector<double> As((size_t)m * np.lda, 0.0), cw(n), ob(n), uc(n);
        std::vector<signed char> kind(n);
        for (int j = 0; j < n; j++)
            for (int p = P->h_aptr[j]; p < P->h_aptr[j + 1]; p++) {
                int i = P->h_aind[p];
                As[(size_t)i * np.lda + j] = P->h_rii[i] * P->h_aval[p] * P->h_sjj[j];
            }
        double cmax = 0.0;
        for (int j = 0; j < n; j++) { ob[j] = P->h_coef[j] * P->h_sjj[j]; uc[j] = P->h_coef[j]; if (cmax < fabs(ob[j])) cmax = fabs(ob[j]); kind[j] = (P->h_kind[j] == GLP_IV); }
        if (cmax == 0.0) cmax = 1.0;
        double zeta = (P->dir == GLP_MIN ? +1.0 : -1.0) / cmax;
        if (fabs(zeta) < 1.0) zeta *= 1000.0;
        np.zeta = zeta;
        np.unit_scale = 1;
        for (int i = 0; i < m; i++) if (P->h_rii[i] != 1.0) np.unit_scale = 0;
        for (int j = 0; j < n; j++) if (P->h_sjj[j] != 1.0) np.unit_scale = 0;
        for (int j = 0; j < n; j++) cw[j] = ob[j] * zeta;
        np.tol_bnd = 1e-7; np.tol_dj = 1e-7; np.tol_piv = 1e-10;       /* glp_init_smcp defaults, lib/glpios01.js:874-891 */
        np.tol_int = parm.tol_int; np.tol_obj = parm.tol_obj;
        np.pp_tech = parm.pp_tech; np.br_tech = parm.br_tech;
        np.it_max = 100 * (m + n) + 1000;
        np.refac_period = 50;
        size_t state = ne_state_bytes(m, n, np.ldb), abytes = (size_t)m * np.lda * 8;
        np.a_in_smem = (state + abytes <= 227u * 1024u) ? 1 : 0;
        smem_bytes = state + (np.a_in_smem ? abytes : 0);
        if (smem_bytes > 227u * 1024u) { glpb_set_error("branch-and-bound: node state does not fit shared memory"); return GLPB_EINVAL; }
#ifndef NE_EMUL
        if (cudaFuncSetAttribute(k_bnb_nodes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) {
            glpb_set_error("branch-and-bound: cannot reserve %zu bytes of shared memory", smem_bytes);
            return GLPB_ENODEV;
        }
#endif
        int bad = 0;
        bad |= dalloc((void **)&d_As, As.size() * 8); bad |= dalloc((void **)&d_cw, n * 8); bad |= dalloc((void **)&d_obj, n * 8);
        bad |= dalloc((void **)&d_ucoef, n * 8); bad |= dalloc((void **)&d_rii, m * 8); bad |= dalloc((void **)&d_sjj, n * 8);
        bad |= dalloc((void **)&d_kind, n);
        bad |= dalloc((void **)&np.slab_lb, (size_t)cap * mn * 8); bad |= dalloc((void **)&np.slab_ub, (size_t)cap * mn * 8);
        bad |= dalloc((void **)&np.slab_type, (size_t)cap * mn); bad |= dalloc((void **)&np.slab_stat, (size_t)cap * mn);
        bad |= dalloc((void **)&np.slab_bi, (size_t)cap * m * np.ldb * 8); bad |= dalloc((void **)&np.slab_upd, (size_t)cap * 4); bad |= dalloc((void **)&np.slab_head, (size_t)cap * m * 4);
        bad |= dalloc((void **)&d_x, (size_t)batch * mn * 8);
        bad |= dalloc((void **)&d_inc, 16); bad |= dalloc((void **)&d_inc_have, 16);
        bad |= dalloc((void **)&d_tasks, batch * sizeof(NeTask)); bad |= dalloc((void **)&d_res, batch * sizeof(NeResult));
        if (bad) { glpb_set_error("branch-and-bound: device allocation failed (slab of %d nodes)", cap); return GLPB_ENOMEM; }
        h2d(d_As, As.data(), As.size() * 8); h2d(d_cw, cw.data(), n * 8); h2d(d_obj, ob.data(), n * 8);
        h2d(d_ucoef, uc.data(), n * 8); h2d(d_rii, P->h_rii.data(), m * 8); h2d(d_sjj, P->h_sjj.data(), n * 8);
        h2d(d_kind, kind.data(), n);
        np.As = d_As; np.cw = d_cw; np.obj = d_obj; np.ucoef = d_ucoef; np.rii = d_rii; np.sjj = d_sjj; np.kind = d_kind;
        np.inc = d_inc; np.inc_have = d_inc_have;
        h_tasks.resize(batch); h_res.resize(batch);
        node.resize(cap); alive.assign(cap, 0);
        free_slots.reserve(cap);
        for (int s = cap - 1; s >= 0; s--) free_slots.push_back(s);
        mipx.assign(mn, 0.0);
        /* root node (ios_create_tree, lib/glpios01.js:83-174): the problem as it stands */
        int root = free_slots.back(); free_slots.pop_back();
        std::vector<signed char> ty(mn), st(mn);
        for (int k = 0; k < mn; k++) { ty[k] = (signed char)P->h_type[k]; st[k] = (signed char)P->h_stat[k]; }
        h2d(np.slab_lb + (size_t)root * mn, P->h_lb.data(), mn * 8); h2d(np.slab_ub + (size_t)root * mn, P->h_ub.data(), mn * 8);
        h2d(np.slab_type + (size_t)root * mn, ty.data(), mn); h2d(np.slab_stat + (size_t)root * mn, st.data(), mn);
        set_no_inverse(root);
        if (dsync()) { glpb_set_error("branch-and-bound: upload failed"); return GLPB_ENODEV; }
        double inf = (P->dir == GLP_MIN ? -DBL_MAX : +DBL_MAX);
        insert(BnbOpen{root, 0, inf, inf, 0.0, 0});
        return 0;
    }

    /* the node in `slot` carries no basis inverse (root, imported nodes): the engine inverts afresh */
    int set_no_inverse(int slot)
    {
        static const int minus1 = -1;
        return h2d(np.slab_upd + slot, &minus1, 4);
    }

    void insert(BnbOpen nd)
    {
        nd.seq = seq++;
        node[nd.slot] = nd; alive[nd.slot] = 1;
        by_bound.insert(std::make_tuple(key_of(nd.bound), -nd.level, nd.seq, nd.slot));
        by_seq.insert(std::make_pair(nd.seq, nd.slot));
    }

    void unlink(int slot)
    {
        const BnbOpen &nd = node[slot];
        by_bound.erase(std::make_tuple(key_of(nd.bound), -nd.level, nd.seq, nd.slot));
        by_seq.erase(std::make_pair(nd.seq, nd.slot));
        alive[slot] = 0;
    }

    void release(int slot) { free_slots.push_back(slot); }

    /* ios_is_hopeful: lib/glpios01.js:789-819 */
    bool is_hopeful(double bound) const
    {
        if (have_cut) {
            double eps = parm.tol_obj * (1.0 + fabs(mip_obj));
            if (P->dir == GLP_MIN) { if (bound >= mip_obj - eps) return false; }
            else { if (bound <= mip_obj + eps) return false; }
        } else {
            if (P->dir == GLP_MIN) { if (bound == +DBL_MAX) return false; }
            else { if (bound == -DBL_MAX) return false; }
        }
        return true;
    }

    /* cleanup_the_tree: lib/glpios03.js:472-495 -- the hopeless nodes are the
       tail of the bound order */
    void cleanup()
    {
        while (!by_bound.empty()) {
            auto it = std::prev(by_bound.end());
            int slot = std::get<3>(*it);
            if (is_hopeful(node[slot].bound)) break;
            unlink(slot);
            release(slot);
        }
    }

    /* best bound over the open nodes (ios_best_node) and ios_relative_gap,
       lib/glpios01.js:821-864 */
    double relative_gap() const
    {
        if (!have_cut) return DBL_MAX;
        if (by_bound.empty()) return 0.0;
        double best_bound = node[std::get<3>(*by_bound.begin())].bound;
        return fabs(mip_obj - best_bound) / (fabs(mip_obj) + DBL_EPSILON);
    }

    /* ios_choose_node (lib/glpios12.js): depth-first = newest, breadth-first =
       oldest, best local bound / best projection = best bound first; when the
       slab runs short the newest (deepest) nodes go first, which drains it */
    int pick()
    {
        bool pressure = (int)free_slots.size() < std::max(4 * batch, cap / 8);
        if (parm.bt_tech == GLP_BT_DFS || pressure) return std::prev(by_seq.end())->second;
        if (parm.bt_tech == GLP_BT_BFS) return by_seq.begin()->second;
        return std::get<3>(*by_bound.begin());
    }

    /* one round; returns 0 = pool exhausted before the round, 1 = round done,
       or a GLP_E* / GLPB_E* code */
    int round(long max_tasks, long *done)
    {
        if (done) *done = 0;
        if (fail_code) return fail_code;
        if (by_seq.empty()) return 0;
        if (parm.tm_lim < INT_MAX && (parm.tm_lim - 1) <= (glpb_now_ms() - tm_beg)) return GLP_ETMLIM;
        if (have_cut && parm.mip_gap > 0.0 && relative_gap() <= parm.mip_gap) return GLP_EMIPGAP;
        long want = batch;
        if (max_tasks >= 0 && max_tasks < want) want = max_tasks;
        if (parm.node_lim >= 0) { long left = parm.node_lim - solved; if (left <= 0) return GLP_ESTOP; if (left < want) want = left; }
        if (want <= 0) return 1;
        int nt = 0;
        while (nt < want && !by_seq.empty() && !free_slots.empty()) {
            int slot = pick();
            unlink(slot);
            if (!is_hopeful(node[slot].bound)) { release(slot); continue; }
            int child = free_slots.back(); free_slots.pop_back();
            const BnbOpen &nd = node[slot];
            h_tasks[nt] = NeTask{slot, child, nd.level, 0, nd.bound, nd.lp_obj};
            nt++;
        }
        if (nt == 0) {
            if (by_seq.empty()) return 0;
            glpb_set_error("branch-and-bound: node slab exhausted (%d nodes); raise GLPB_BNB_SLAB", cap);
            return fail_code = GLPB_ENOMEM;
        }
        double incv[2] = {have_cut ? mip_obj : 0.0, 0.0};
        int inch[4] = {have_cut ? 1 : 0, 0, 0, 0};
        int bad = 0;
        bad |= h2d(d_inc, incv, 16); bad |= h2d(d_inc_have, inch, 16);
        bad |= h2d(d_tasks, h_tasks.data(), nt * sizeof(NeTask));
        bad |= launch(nt);
        bad |= d2h(h_res.data(), d_res, nt * sizeof(NeResult));
        bad |= dsync();
        if (bad) {
#ifndef NE_EMUL
            glpb_set_error("branch-and-bound: device error: %s", cudaGetErrorString(cudaGetLastError()));
#endif
            return fail_code = GLPB_ENODEV;
        }
        rounds++;
        /* integral nodes first: the best one becomes the incumbent */
        int best_t = -1;
        for (int b = 0; b < nt; b++) {
            const NeResult &r = h_res[b];
            if (r.code != NE_R_INTEGRAL) continue;
            bool better = !have_cut || (P->dir == GLP_MIN ? r.obj < mip_obj : r.obj > mip_obj);
            if (better && (best_t < 0 || (P->dir == GLP_MIN ? r.obj < h_res[best_t].obj : r.obj > h_res[best_t].obj))) best_t = b;
        }
        if (best_t >= 0) {
            /* record_solution: lib/glpios03.js:118-139 */
            if (d2h(mipx.data(), d_x + (size_t)best_t * mn, mn * 8) || dsync()) return fail_code = GLPB_ENODEV;
            mip_obj = h_res[best_t].obj;
            have_sol = have_cut = true;
        }
        for (int b = 0; b < nt; b++) {
            const NeResult &r = h_res[b];
            const NeTask &tk = h_tasks[b];
            solved += r.solves; iters += r.iters; refacs += r.refacs; tasks_done++;
            for (int c = 0; c < 6; c++) cyc[c] += r.cyc[c];
            if (r.code == NE_R_BRANCH) {
                BnbOpen dn{tk.node, tk.level + 1, r.bound, r.dn_lp, r.ii_sum, 0};
                BnbOpen up{tk.child, tk.level + 1, r.bound, r.up_lp, r.ii_sum, 0};
                auto improve = [&](double &bd, double v) { if (P->dir == GLP_MIN) { if (bd < v) bd = v; } else { if (bd > v) bd = v; } };
                improve(dn.bound, r.dn_bnd); improve(up.bound, r.up_bnd);
                /* the child Driebeck-Tomlin suggests is inserted last = dived into first */
                BnbOpen first = (r.next == NE_DN_BRNCH) ? up : dn, second = (r.next == NE_DN_BRNCH) ? dn : up;
                if (is_hopeful(first.bound)) insert(first); else release(first.slot);
                if (is_hopeful(second.bound)) insert(second); else release(second.slot);
            } else if (r.code == NE_R_FATHOM || r.code == NE_R_INTEGRAL) {
                release(tk.node); release(tk.child);
            } else {
                release(tk.node); release(tk.child);
                glpb_set_error("branch-and-bound: node LP failed (simplex return code %d)", r.ret);
                fail_code = GLP_EFAIL;
            }
        }
        if (have_cut) cleanup();
        if (done) *done = nt;
        return fail_code ? fail_code : 1;
    }

    /* ---- migration between ranks: a record is
           [bound, lp_obj, level, up_ii_sum] (32 B) | lb[mn] | ub[mn] | type[mn] | stat[mn] | pad to 16 ---- */
    size_t record_bytes() const { return 32 + (size_t)mn * 16 + (((size_t)2 * mn + 15) & ~(size_t)15); }

    /* every second node of the bound order leaves, the best one stays */
    int export_nodes(int max_count, unsigned char *dev_buf, int *count)
    {
        int done_ = 0;
        const size_t rb = record_bytes();
        std::vector<int> out;
        int idx = 0;
        for (auto it = by_bound.begin(); it != by_bound.end() && (int)out.size() < max_count; ++it, ++idx)
            if (idx & 1) out.push_back(std::get<3>(*it));
        for (int slot : out) {
            const BnbOpen &nd = node[slot];
            double hdr[4] = {nd.bound, nd.lp_obj, (double)nd.level, nd.up_ii_sum};
            unsigned char *rec = dev_buf + (size_t)done_ * rb;
            int bad = h2d(rec, hdr, 32);
            bad |= d2d(rec + 32, np.slab_lb + (size_t)slot * mn, (size_t)mn * 8);
            bad |= d2d(rec + 32 + (size_t)mn * 8, np.slab_ub + (size_t)slot * mn, (size_t)mn * 8);
            bad |= d2d(rec + 32 + (size_t)mn * 16, np.slab_type + (size_t)slot * mn, mn);
            bad |= d2d(rec + 32 + (size_t)mn * 17, np.slab_stat + (size_t)slot * mn, mn);
            bad |= dsync();              /* hdr is a stack buffer */
            if (bad) return GLPB_ENODEV;
            unlink(slot); release(slot);
            done_++;
        }
        *count = done_;
        return 0;
    }

    int import_nodes(const unsigned char *dev_buf, int count)
    {
        const size_t rb = record_bytes();
        for (int c = 0; c < count; c++) {
            const unsigned char *rec = dev_buf + (size_t)c * rb;
            double hdr[4];
            if (d2h(hdr, rec, 32) || dsync()) return GLPB_ENODEV;
            if (!is_hopeful(hdr[0])) continue;
            if (free_slots.empty()) { glpb_set_error("branch-and-bound: node slab exhausted on import"); return GLPB_ENOMEM; }
            int slot = free_slots.back(); free_slots.pop_back();
            int bad = d2d(np.slab_lb + (size_t)slot * mn, rec + 32, (size_t)mn * 8);
            bad |= d2d(np.slab_ub + (size_t)slot * mn, rec + 32 + (size_t)mn * 8, (size_t)mn * 8);
            bad |= d2d(np.slab_type + (size_t)slot * mn, rec + 32 + (size_t)mn * 16, mn);
            bad |= d2d(np.slab_stat + (size_t)slot * mn, rec + 32 + (size_t)mn * 17, mn);
            bad |= set_no_inverse(slot);
            if (bad) return GLPB_ENODEV;
            insert(BnbOpen{slot, (int)hdr[2], hdr[0], hdr[1], hdr[3], 0});
        }
        return dsync() ? GLPB_ENODEV : 0;
    }

    /* drop every open node (a rank that starts empty and is fed by migration) */
    void clear()
    {
        while (!by_seq.empty()) { int slot = by_seq.begin()->second; unlink(slot); release(slot); }
    }

    void set_cutoff(double obj)
    {
        bool better = !have_cut || (P->dir == GLP_MIN ? obj < mip_obj : obj > mip_obj);
        if (better) { mip_obj = obj; have_cut = true; have_sol = false; cleanup(); }
    }
};

#endif /* GLPB_BNBPOOL_CUH */
