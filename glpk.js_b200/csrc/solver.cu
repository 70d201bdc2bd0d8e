/* solver.cu -- host side of libglpb200: device-resident problem handle, basis
 * solves, refactorisation, and the loop controllers that replay the control
 * flow of spx_primal (lib/glpspx01.js:1684-2056) and spx_dual
 * (lib/glpspx02.js:1592-1966) around the kernels in kernels.cuh.
 *
 * There is no CPU arithmetic on the solve path: the host only decides which
 * kernels to enqueue next from the device's status word, exactly where the
 * reference's loop branches.
 */
#include "kernels.cuh"
#include "engine.cuh"
#include "refactor.cuh"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

static thread_local char g_err[512] = "";

void glpb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

static double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

#define PROF_MAX_RECS 400000
#define REF_SMEM_MAX (200 * 1024)
#define RHO_SMEM_MAX (200 * 1024)

static inline void prof_begin(glpb_prob *P, const char *name)
{
    if (!P->prof || P->prof_recs.size() >= PROF_MAX_RECS) return;
    glpb_prob::ProfRec r;
    r.name = name; r.bytes = P->next_bytes;
    cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, P->stream);
    P->prof_recs.push_back(r);
}

static inline void prof_end(glpb_prob *P)
{
    P->next_bytes = 0.0;
    if (!P->prof || P->prof_recs.empty() || P->prof_recs.size() > PROF_MAX_RECS) return;
    cudaEventRecord(P->prof_recs.back().e1, P->stream);
}

/* fold the recorded event pairs into per-kernel totals (after a sync) */
static void prof_collect(glpb_prob *P)
{
    if (P->prof_recs.empty()) return;
    cudaStreamSynchronize(P->stream);
    for (auto &r : P->prof_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
            std::string nm(r.name);
            size_t lt = nm.find('<');
            if (lt != std::string::npos) nm = nm.substr(0, lt);
            auto &a = P->prof_acc[nm];
            a.ms += ms; a.bytes += r.bytes; a.count++;
        }
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    }
    P->prof_recs.clear();
}

#define LAUNCH(P, kern, grid, block, smem, ...)                                \
    do {                                                                       \
        prof_begin((P), #kern);                                                \
        kern<<<(grid), (block), (smem), (P)->stream>>>(__VA_ARGS__);            \
        prof_end(P);                                                           \
        (P)->n_launch++;                                                       \
    } while (0)

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

/* lanes per sparse vector, by average length */
static inline int pick_group(double avg) { return avg >= 48.0 ? 32 : (avg >= 12.0 ? 8 : 4); }

#define GROUP_DISPATCH(G, CALL)                                                \
    do {                                                                       \
        if ((G) == 32) { constexpr int GG = 32; CALL; }                        \
        else if ((G) == 8) { constexpr int GG = 8; CALL; }                     \
        else { constexpr int GG = 4; CALL; }                                   \
    } while (0)

template <class T> static int dalloc(T **p, size_t count)
{
    cudaError_t e = cudaMalloc((void **)p, (count ? count : 1) * sizeof(T));
    if (e != cudaSuccess) { glpb_set_error("cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e)); return GLPB_ENOMEM; }
    return 0;
}

template <class T> static int h2d(glpb_prob *P, T *dst, const T *src, size_t count)
{
    CK(cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyHostToDevice, P->stream));
    return 0;
}

static int sync_ctrl(glpb_prob *P)
{
    CK(cudaMemcpyAsync(P->h_ctrl, P->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, P->stream));
    CK(cudaStreamSynchronize(P->stream));
    P->n_sync++;
    P->k_host = P->h_ctrl->k;
    return 0;
}

/* ------------------------------------------------------------------ */
/* handle                                                             */
/* ------------------------------------------------------------------ */

extern "C" int glpb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char *glpb_last_error(void) { return g_err; }
extern "C" const char *glpb_version(void) { return "glpb200 0.1 (sm_100a); GLPK 4.49 semantics"; }

extern "C" void glpb_init_smcp(glpb_smcp *p)
{
    p->msg_lev = 3; p->meth = GLP_PRIMAL; p->pricing = GLP_PT_PSE; p->r_test = GLP_RT_HAR;
    p->tol_bnd = 1e-7; p->tol_dj = 1e-7; p->tol_piv = 1e-10;
    p->obj_ll = -DBL_MAX; p->obj_ul = +DBL_MAX;
    p->it_lim = INT_MAX; p->tm_lim = INT_MAX; p->out_frq = 500; p->out_dly = 0; p->presolve = 0;
}

extern "C" void glpb_init_iocp(glpb_iocp *p)
{
    p->msg_lev = 3; p->br_tech = GLP_BR_DTH; p->bt_tech = GLP_BT_BLB;
    p->tol_int = 1e-5; p->tol_obj = 1e-7; p->tm_lim = INT_MAX; p->out_frq = 5000;
    p->out_dly = 10000; p->pp_tech = GLP_PP_ALL; p->mip_gap = 0.0; p->presolve = 0;
    p->node_lim = -1;
}

extern "C" void glpb_destroy(glpb_prob *P)
{
    if (!P) return;
    cudaSetDevice(P->device);
    void *ptrs[] = {P->a_ptr, P->a_ind, P->at_ptr, P->at_ind, P->a_val, P->at_val, P->type,
                    P->orig_type, P->stat, P->refsp, P->lb, P->ub, P->coef, P->orig_lb, P->orig_ub,
                    P->obj, P->head, P->bind, P->bbar, P->cbar, P->gamma, P->tcol, P->trow, P->rho,
                    P->svec, P->w1, P->w2, P->w3, P->w4, P->w5, P->yk, P->wk, P->yk2, P->zn, P->eng_slots, P->eng_cols, P->eng_fr,
                    P->eng_cyc, P->eng_bytes, P->T, P->T2, P->ref_slots, P->ref_flags, P->ref_xp, P->partial,
                    P->rslot, P->slot_pos, P->cslot, P->slot_row, P->gj_piv, P->gj_row,
                    P->scratch, P->ctrl, P->sort_list, P->sort_tmp};
    for (void *p : ptrs) if (p) cudaFree(p);
    P->prof = 0; prof_collect(P);
    for (auto &kind : P->graphs)
        for (auto &g : kind) if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    if (P->piv_log) { cudaFree(P->piv_log); P->piv_log = nullptr; }
    if (P->mip) { glpb_mip_free(P->mip); P->mip = nullptr; }
    if (P->bnb) { glpb_bnb_free(P->bnb); P->bnb = nullptr; }
    if (P->h_ctrl) cudaFreeHost(P->h_ctrl);
    if (P->h_stage) cudaFreeHost(P->h_stage);
    if (P->stream) cudaStreamDestroy(P->stream);
    delete P;
}

static int create_device(glpb_prob *P)
{
    const int m = P->m, n = P->n, nnz = P->nnz;
    CK(cudaSetDevice(P->device));
    CK(cudaStreamCreateWithFlags(&P->stream, cudaStreamNonBlocking));
    int rc;
#define DA(ptr, cnt) if ((rc = dalloc(&P->ptr, (size_t)(cnt))) != 0) return rc
    DA(a_ptr, n + 1); DA(a_ind, nnz); DA(a_val, nnz);
    DA(at_ptr, m + 1); DA(at_ind, nnz); DA(at_val, nnz);
    DA(type, m + n); DA(orig_type, m + n); DA(stat, n); DA(refsp, m + n);
    DA(lb, m + n); DA(ub, m + n); DA(coef, m + n); DA(orig_lb, m + n); DA(orig_ub, m + n);
    DA(obj, n + 1); DA(head, m + n); DA(bind, m + n);
    DA(bbar, m); DA(cbar, n); DA(gamma, std::max(m, n));
    DA(tcol, m); DA(trow, n); DA(rho, m); DA(svec, n);
    DA(w1, m); DA(w2, m); DA(w3, m); DA(w4, m); DA(w5, m);
    P->ldt = (m + 7) & ~7;
    DA(yk, P->ldt); DA(wk, P->ldt); DA(yk2, P->ldt); DA(zn, P->ldt);
    DA(eng_slots, ENG_RING * ENG_MAXG + 1); DA(eng_cols, 3 * (size_t)n + m); DA(eng_fr, 2 * (size_t)ENG_DB * P->ldt + ENG_DB); DA(eng_cyc, 32); DA(eng_bytes, 24);
    DA(T, (size_t)P->ldt * P->ldt);
    DA(T2, (size_t)P->ldt * P->ldt);
    DA(ref_slots, ENG_RING * ENG_MAXG); DA(ref_flags, ENG_RING * ENG_MAXG * 32); DA(ref_xp, (size_t)REF_NB * P->ldt);
    P->partial_rows = cdiv(P->ldt, GEMV_TILE);
    DA(partial, (size_t)P->partial_rows * P->ldt);
    DA(rslot, m); DA(slot_pos, P->ldt); DA(cslot, m); DA(slot_row, P->ldt);
    DA(gj_piv, P->ldt); DA(gj_row, P->ldt);
    DA(scratch, 4096);
    DA(ctrl, 1);
    DA(sort_list, std::max(m, n)); DA(sort_tmp, 2 * (size_t)std::max(m, n) + 2);
#undef DA
    CK(cudaMallocHost((void **)&P->h_ctrl, sizeof(Ctrl)));
    /* pinned staging: bounds/costs [6 x (m+n) doubles + n+1], types [2(m+n)], header [2(m+n) ints + n],
       and the read-back of head/stat/bbar/cbar; all copies of a solve are asynchronous */
    P->stage_bytes = (size_t)(6 * (m + n) + n + 1 + m + n) * sizeof(double) + (size_t)4 * (m + n) * sizeof(int) + 4 * (size_t)(m + n) + 256;
    CK(cudaMallocHost((void **)&P->h_stage, P->stage_bytes));
    CK(cudaMemsetAsync(P->ctrl, 0, sizeof(Ctrl), P->stream));
    CK(cudaMemsetAsync(P->eng_cyc, 0, 32 * sizeof(long long), P->stream));
    CK(cudaMemsetAsync(P->eng_bytes, 0, 24 * sizeof(double), P->stream));
    /* persistent engine: one CTA per SM, dynamic shared memory for staging */
    {
        int coop = 0, smem_max = 0;
        CK(cudaDeviceGetAttribute(&P->sm_count, cudaDevAttrMultiProcessorCount, P->device));
        CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, P->device));
        CK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, P->device));
        /* dynamic shared memory of the engine: [dcap doubles: list / vector / flush staging]
           [ENG_LCAP ints][32 x 65 doubles][optional private copy of the basis header] */
        const int want_tma = getenv("GLPB_TMA") ? atoi(getenv("GLPB_TMA")) : 0;
        const int want_hdr = getenv("GLPB_HDR") ? atoi(getenv("GLPB_HDR")) : 1;
        const size_t total = (size_t)std::min(smem_max, 200 * 1024) - 18 * 1024;      /* minus the static part */
        const size_t fixed = (size_t)ENG_LCAP * 4 + 32 * 65 * 8;
        const size_t hdr = ((size_t)2 * (m + n) + 2 * (size_t)m + 2 * (size_t)P->ldt) * 4 + n + 16;
        const size_t dmin = std::max({(size_t)ENG_LCAP, (size_t)ENG_FLUSH_SMEM, want_tma ? (size_t)ENG_TMA_SMEM : (size_t)0});
        P->eng_tma = want_tma;
        P->eng_hdr = (want_hdr && dmin * 8 + fixed + hdr <= total) ? 1 : 0;
        const size_t room = total - fixed - (P->eng_hdr ? hdr : 0);
        int dcap = (int)std::max(dmin, std::min((size_t)P->ldt, room / 8));
        P->eng_dcap = dcap;
        P->eng_smem = (int)((size_t)dcap * 8 + fixed + (P->eng_hdr ? hdr : 0));
        if (!coop || P->sm_count > ENG_MAXG) { glpb_set_error("device lacks cooperative launch"); return GLPB_ENODEV; }
        /* two instantiations of each engine: basis header in per-CTA shared memory (small LPs) or read straight
           from global memory through the kernel parameters (no header pointer held in registers) */
        CK(cudaFuncSetAttribute(k_engine_primal<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, P->eng_smem));
        CK(cudaFuncSetAttribute(k_engine_dual<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, P->eng_smem));
        CK(cudaFuncSetAttribute(k_engine_primal<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, P->eng_smem));
        CK(cudaFuncSetAttribute(k_engine_dual<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, P->eng_smem));
        CK(cudaFuncSetAttribute(k_refactor, cudaFuncAttributeMaxDynamicSharedMemorySize, REF_SMEM_MAX));
        CK(cudaFuncSetAttribute(k_rho_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, RHO_SMEM_MAX));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_engine_dual<true>, ENG_NT, P->eng_smem));
        if (occ < 1) { glpb_set_error("engine does not fit an SM"); return GLPB_ENODEV; }
        P->eng_ready = 1;
    }
    {
        /* replayed launch sequences (Loop::graphed) for LPs small enough that launch cost,
           not arithmetic, is the time of a recomputation; GLPB_GRAPH=0 turns them off */
        const int want = getenv("GLPB_GRAPH") ? atoi(getenv("GLPB_GRAPH")) : 1;
        P->graph_ok = (want && m <= 512 && n <= 8192) ? 1 : 0;
    }
    return 0;
}

/* scaled copies of A by columns and by rows; the row copy keeps columns in
   ascending order, which is the row-list order after glp_sort_matrix */
static int upload_matrix(glpb_prob *P)
{
    const int m = P->m, n = P->n, nnz = P->nnz;
    std::vector<double> sval(nnz);
    for (int j = 0; j < n; j++)
        for (int t = P->h_aptr[j]; t < P->h_aptr[j + 1]; t++)
            sval[t] = P->h_rii[P->h_aind[t]] * P->h_aval[t] * P->h_sjj[j];
    P->h_atptr.assign(m + 1, 0);
    for (int t = 0; t < nnz; t++) P->h_atptr[P->h_aind[t] + 1]++;
    for (int i = 0; i < m; i++) P->h_atptr[i + 1] += P->h_atptr[i];
    P->h_atind.assign(nnz, 0); P->h_atval.assign(nnz, 0.0);
    std::vector<double> satval(nnz);
    std::vector<int> fill(P->h_atptr.begin(), P->h_atptr.end() - 1);
    for (int j = 0; j < n; j++)
        for (int t = P->h_aptr[j]; t < P->h_aptr[j + 1]; t++) {
            int pos = fill[P->h_aind[t]]++;
            P->h_atind[pos] = j; P->h_atval[pos] = P->h_aval[t]; satval[pos] = sval[t];
        }
    int rc;
    if ((rc = h2d(P, P->a_ptr, P->h_aptr.data(), n + 1))) return rc;
    if ((rc = h2d(P, P->a_ind, P->h_aind.data(), nnz))) return rc;
    if ((rc = h2d(P, P->a_val, sval.data(), nnz))) return rc;
    if ((rc = h2d(P, P->at_ptr, P->h_atptr.data(), m + 1))) return rc;
    if ((rc = h2d(P, P->at_ind, P->h_atind.data(), nnz))) return rc;
    if ((rc = h2d(P, P->at_val, satval.data(), nnz))) return rc;
    CK(cudaStreamSynchronize(P->stream));
    return 0;
}

static int default_stat(int type, double lb, double ub, int stat)
{
    /* glp_set_row_bnds / glp_set_col_bnds status rule, lib/glpapi01.js:217-281 */
    if (stat == GLP_BS) return stat;
    switch (type) {
    case GLP_FR: return GLP_NF;
    case GLP_LO: return GLP_NL;
    case GLP_UP: return GLP_NU;
    case GLP_DB: return (stat == GLP_NL || stat == GLP_NU) ? stat : (fabs(lb) <= fabs(ub) ? GLP_NL : GLP_NU);
    default: return GLP_NS;
    }
}

extern "C" glpb_prob *glpb_create(int m, int n, int nnz, int dir, double c0, const int *type,
                                  const double *lb, const double *ub, const double *coef,
                                  const int *kind, const double *rii, const double *sjj,
                                  const int *A_ptr, const int *A_ind, const double *A_val, int device)
{
    if (m < 1 || n < 1 || nnz < 0 || !type || !lb || !ub || !coef || !A_ptr || (nnz && (!A_ind || !A_val)) ||
        !(dir == GLP_MIN || dir == GLP_MAX)) {
        glpb_set_error("glpb_create: invalid argument");
        return nullptr;
    }
    /* the column pointer must be a monotone 0-based prefix sum that ends at nnz, row indices in range */
    if (A_ptr[0] != 0 || A_ptr[n] != nnz) { glpb_set_error("glpb_create: A_ptr[0] must be 0 and A_ptr[n] must equal nnz"); return nullptr; }
    for (int j = 0; j < n; j++)
        if (A_ptr[j + 1] < A_ptr[j]) { glpb_set_error("glpb_create: A_ptr is not monotone at column %d", j + 1); return nullptr; }
    int ndev = glpb_device_count();
    if (ndev < 1 || device < 0 || device >= ndev) {
        glpb_set_error("glpb_create: no CUDA device %d (found %d); there is no CPU fallback", device, ndev);
        return nullptr;
    }
    glpb_prob *P = new glpb_prob();
    P->device = device; P->m = m; P->n = n; P->nnz = nnz; P->dir = dir; P->c0 = c0;
    P->h_type.assign(type, type + m + n);
    P->h_lb.assign(lb, lb + m + n); P->h_ub.assign(ub, ub + m + n);
    P->h_coef.assign(coef, coef + n);
    if (kind) P->h_kind.assign(kind, kind + n); else P->h_kind.assign(n, GLP_CV);
    if (rii) P->h_rii.assign(rii, rii + m); else P->h_rii.assign(m, 1.0);
    if (sjj) P->h_sjj.assign(sjj, sjj + n); else P->h_sjj.assign(n, 1.0);
    P->h_aptr.assign(A_ptr, A_ptr + n + 1);
    P->h_aind.assign(A_ind, A_ind + nnz);
    P->h_aval.assign(A_val, A_val + nnz);
    for (int t = 0; t < nnz; t++)
        if (A_ind[t] < 0 || A_ind[t] >= m) { glpb_set_error("glpb_create: row index out of range"); delete P; return nullptr; }
    /* new rows are basic, new columns non-basic (lib/glpapi01.js:118,166) */
    P->h_stat.assign(m + n, GLP_BS);
    for (int j = 0; j < n; j++) P->h_stat[m + j] = default_stat(type[m + j], lb[m + j], ub[m + j], GLP_NS);
    for (int k = 0; k < m + n; k++) {
        /* normalise unused bound fields like glp_set_*_bnds does */
        switch (P->h_type[k]) {
        case GLP_FR: P->h_lb[k] = P->h_ub[k] = 0.0; break;
        case GLP_LO: P->h_ub[k] = 0.0; break;
        case GLP_UP: P->h_lb[k] = 0.0; break;
        case GLP_FX: P->h_ub[k] = P->h_lb[k]; break;
        default: break;
        }
    }
    P->h_head.assign(m, 0);
    P->h_prim.assign(m + n, 0.0); P->h_dual.assign(m + n, 0.0); P->h_mipx.assign(m + n, 0.0);
    P->trace = getenv("GLPB_TRACE") ? atoi(getenv("GLPB_TRACE")) : 0;
    if (create_device(P) != 0 || upload_matrix(P) != 0) { glpb_destroy(P); return nullptr; }
    return P;
}

extern "C" int glpb_set_bounds(glpb_prob *P, int count, const int *k, const int *type,
                               const double *lb, const double *ub)
{
    if (!P || count < 0 || (count > 0 && (!k || !type || !lb || !ub))) return GLPB_EINVAL;
    for (int t = 0; t < count; t++) {
        int kk = k[t] - 1;
        if (kk < 0 || kk >= P->m + P->n || type[t] < GLP_FR || type[t] > GLP_FX) return GLPB_EINVAL;
        double l = lb[t], u = ub[t];
        switch (type[t]) {
        case GLP_FR: l = u = 0.0; break;
        case GLP_LO: u = 0.0; break;
        case GLP_UP: l = 0.0; break;
        case GLP_FX: u = l; break;
        default: break;
        }
        P->h_type[kk] = type[t]; P->h_lb[kk] = l; P->h_ub[kk] = u;
        P->h_stat[kk] = default_stat(type[t], l, u, P->h_stat[kk]);
    }
    return 0;
}

extern "C" int glpb_set_basis(glpb_prob *P, const int *stat)
{
    if (!P || !stat) return GLPB_EINVAL;
    for (int k = 0; k < P->m + P->n; k++) {
        int s = stat[k];
        if (s < GLP_BS || s > GLP_NS) return GLPB_EINVAL;
        if (s != GLP_BS) { /* lib/glpapi05.js:9-17 */
            switch (P->h_type[k]) {
            case GLP_FR: s = GLP_NF; break;
            case GLP_LO: s = GLP_NL; break;
            case GLP_UP: s = GLP_NU; break;
            case GLP_DB: if (s != GLP_NU) s = GLP_NL; break;
            default: s = GLP_NS;
            }
        }
        if ((P->h_stat[k] == GLP_BS) != (s == GLP_BS)) P->valid = 0;
        P->h_stat[k] = s;
    }
    return 0;
}

extern "C" int glpb_std_basis(glpb_prob *P)
{
    if (!P) return GLPB_EINVAL;
    std::vector<int> st(P->m + P->n, GLP_BS);
    for (int j = 0; j < P->n; j++) {
        int k = P->m + j;
        st[k] = (P->h_type[k] == GLP_DB && fabs(P->h_lb[k]) > fabs(P->h_ub[k])) ? GLP_NU : GLP_NL;
    }
    return glpb_set_basis(P, st.data());
}

extern "C" int glpb_set_bfcp(glpb_prob *P, const glpb_bfcp *parm)
{
    if (!P) return GLPB_EINVAL;
    if (!parm) { P->bfcp = {100, 0.10, 1e-6}; return 0; }
    if (parm->nfs_max < 1 || parm->nfs_max > 32767) return GLPB_EINVAL;
    P->bfcp = *parm;
    return 0;
}

extern "C" int glpb_set_it_cnt(glpb_prob *P, int it_cnt)
{
    if (!P) return GLPB_EINVAL;
    P->it_cnt = it_cnt;
    return 0;
}

/* ------------------------------------------------------------------ */
/* basis solves                                                       */
/* ------------------------------------------------------------------ */

struct Dev { /* launch geometry derived from the handle */
    glpb_prob *P;
    int m, n, gc, gr, k;
    int kpad = 0;   /* head-room for grids when several iterations are enqueued (k grows by <= 1 each) */
    explicit Dev(glpb_prob *P_) : P(P_), m(P_->m), n(P_->n)
    {
        gc = pick_group((double)P->nnz / P->n);
        gr = pick_group((double)P->nnz / P->m);
        k = 0;
    }
};

/* rho = row ctrl->p of inv(B); k is the host's view of the kernel size */
static void launch_rho(glpb_prob *P, int k)
{
    const int m = P->m;
    size_t smem = (size_t)std::max(k, 1) * sizeof(double);
    if (smem <= RHO_SMEM_MAX) {
        P->next_bytes = 8.0 * k * (double)k + 12.0 * m;
        LAUNCH(P, k_rho_fast, cdiv(std::max((long)k * 32, (long)m), 256), 256, smem, P->ctrl, m, P->T, P->ldt,
               P->at_ptr, P->at_ind, P->at_val, P->head, P->bind, P->rslot, P->cslot, P->slot_row, P->rho);
    } else
        LAUNCH(P, k_rho, cdiv(m, 128), 128, 0, P->ctrl, m, P->T, P->ldt, P->at_ptr, P->at_ind, P->at_val,
               P->head, P->bind, P->rslot, P->cslot, P->rho);
}

/* x = inv(B) h   (bfd_ftran, lib/glpbfd.js:148-157); h and x must differ.
   cond != 0: the kernels run only while PSE weights are live (device flag) */
static void dev_ftran(Dev &D, const double *h, double *x, int cond = 0)
{
    glpb_prob *P = D.P;
    const int kg = D.k + D.kpad;
    if (kg > 0) {
        int tl = cdiv(kg, GEMV_TILE);
        if (kg <= 2048) {
            P->next_bytes = 8.0 * D.k * (double)D.k;
            LAUNCH(P, k_gemvN_part, dim3(tl, 1), GEMV_TILE, 0, P->ctrl, P->T, P->ldt, h, P->slot_row, P->yk, 1, cond);
        } else {
            P->next_bytes = 8.0 * D.k * (double)D.k;
            LAUNCH(P, k_gemvN_part, dim3(tl, tl), GEMV_TILE, 0, P->ctrl, P->T, P->ldt, h, P->slot_row, P->partial, 0, cond);
            LAUNCH(P, k_gemvN_fin, cdiv(kg, 256), 256, 0, P->ctrl, P->ldt, P->partial, P->yk, cond);
        }
    }
    P->next_bytes = 12.0 * P->nnz * (1.0 - (double)D.k / D.m) + 16.0 * D.m;
    GROUP_DISPATCH(D.gr, LAUNCH(P, k_ftran_tail<GG>, cdiv((long)D.m * GG, 256), 256, 0, P->ctrl, D.m,
                                P->at_ptr, P->at_ind, P->at_val, P->head, P->bind, P->rslot, h, P->yk, x, cond));
}

/* z = inv(B') c  (bfd_btran, lib/glpbfd.js:159-168); c and z must differ */
static void dev_btran(Dev &D, const double *c, double *z, int cond = 0)
{
    glpb_prob *P = D.P;
    const int kg = D.k + D.kpad;
    if (kg > 0) {
        GROUP_DISPATCH(D.gc, LAUNCH(P, k_btran_head<GG>, cdiv((long)kg * GG, 256), 256, 0, P->ctrl, D.m,
                                    P->a_ptr, P->a_ind, P->a_val, P->head, P->bind, P->slot_pos, c, P->wk, cond));
        P->next_bytes = 8.0 * D.k * (double)D.k;
        LAUNCH(P, k_gemvT, cdiv((long)kg * 32, 256), 256, 0, P->ctrl, P->T, P->ldt, P->wk, P->yk, cond);
    }
    LAUNCH(P, k_btran_tail, cdiv(D.m, 256), 256, 0, P->ctrl, D.m, P->cslot, P->bind, c, P->yk, z, cond);
}

/* refactorisation (invert_B -> bfd_factorize -> luf_factorize in the
   reference, lib/glpspx01.js:177-181): rebuild T for the current basis header
   by Gauss-Jordan with partial pivoting on the k x k structural kernel */
static int dev_refactor(Dev &D)
{
    glpb_prob *P = D.P;
    LAUNCH(P, k_build_slots, 1, 1024, 0, P->ctrl, D.m, P->head, P->bind, P->rslot, P->slot_pos, P->cslot, P->slot_row);
    int rc = sync_ctrl(P);
    if (rc) return rc;
    D.k = P->h_ctrl->k;
    if (P->h_ctrl->sing) return GLP_EBADB;
    const int k = D.k;
    P->n_refac++;
    if (k > 0) {
        /* one cooperative launch: rows of the panel are spread over G CTAs */
        static const int envg = getenv("GLPB_REF_GRID") ? atoi(getenv("GLPB_REF_GRID")) : 0;
        int G = std::min(P->sm_count, std::max(1, cdiv(k, 16)));
        if (envg > 0) G = std::min(envg, P->sm_count);
        /* panel: whole in CTA 0's shared memory while k * nbr doubles fit, else its rows
           are spread over all CTAs (one grid-wide arg-reduction per pivot) */
        static const int envs = getenv("GLPB_REF_SINGLE") ? atoi(getenv("GLPB_REF_SINGLE")) : -1;
        int nbr = 0, single = 0;
        for (int nb = REF_NB; nb >= 8 && !nbr; nb >>= 1)
            if ((size_t)k * (nb + 2) * sizeof(double) + 64 <= REF_SMEM_MAX) { nbr = nb; single = 1; }
        if (envs == 0 || !nbr) { nbr = REF_NB; single = 0; }
        static const int envpg = getenv("GLPB_REF_PG") ? atoi(getenv("GLPB_REF_PG")) : 32;
        const int pg = std::max(1, std::min(G, std::max(envpg, cdiv(k, 700))));     /* panel CTAs; <= 700 rows each */
        const int R = single ? k : cdiv(k, pg);
        const size_t smem = std::max({(size_t)R * (nbr + 2) * sizeof(double) + 64, (size_t)2 * k * sizeof(int), (size_t)REF_UPD_SMEM});
        if (G > P->sm_count || smem > REF_SMEM_MAX) {
            glpb_set_error("refactorisation: kernel of size %d does not fit the device", k);
            return GLPB_ENOMEM;
        }
        CK(cudaMemsetAsync(P->ref_slots, 0, ENG_RING * ENG_MAXG * sizeof(RefSlot), P->stream));
        CK(cudaMemsetAsync(P->ref_flags, 0, ENG_RING * ENG_MAXG * 32 * sizeof(unsigned int), P->stream));
        RefArgs A;
        A.ctrl = P->ctrl; A.m = D.m; A.ldt = P->ldt; A.X = P->T; A.T2 = P->T2;
        A.a_ptr = P->a_ptr; A.a_ind = P->a_ind; A.a_val = P->a_val;
        A.head = P->head; A.slot_pos = P->slot_pos; A.cslot = P->cslot;
        A.piv = P->gj_piv; A.slots = P->ref_slots; A.flags = P->ref_flags;
        A.prof_cyc = P->prof ? P->eng_cyc + 24 : nullptr;
        A.nbr = nbr; A.single = single; A.pg = pg; A.xp = P->ref_xp; A.plan = (int *)P->gj_row;   /* gj_row: ldt doubles, used as scratch ints */
        void *args[] = {&A};
        P->next_bytes = 16.0 * k * (double)k * cdiv(k, REF_NB);
        prof_begin(P, "k_refactor");
        cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_refactor, dim3(G), dim3(REF_NT), args, smem, P->stream);
        prof_end(P);
        if (e != cudaSuccess) { glpb_set_error("refactor launch: %s", cudaGetErrorString(e)); return GLPB_ENODEV; }
        P->n_launch++;
        std::swap(P->T, P->T2);
        rc = sync_ctrl(P);
        if (rc) return rc;
        if (P->h_ctrl->sing) return GLP_ESING;
    }
    return 0;
}

/* upload the basis header in init_csa order (lib/glpspx01.js:107-129) */
/* carve typed arrays out of the pinned staging buffer */
struct Stage {
    unsigned char *p;
    explicit Stage(glpb_prob *P, size_t off) : p(P->h_stage + off) {}
    template <class T> T *take(size_t count)
    {
        size_t a = (size_t)p & 7;
        if (a) p += 8 - a;
        T *r = (T *)p;
        p += count * sizeof(T);
        return r;
    }
};
/* staging layout: [header | bounds]; the two halves are filled by different calls */
static size_t stage_off_bounds(const glpb_prob *P) { return (size_t)2 * (P->m + P->n) * sizeof(int) + P->n + 64; }

static int upload_basis(glpb_prob *P)
{
    const int m = P->m, n = P->n;
    /* every entry point leaves the stream idle, and the two staging regions are filled
       at most once per device synchronisation: no copy is in flight from this memory */
    Stage S(P, 0);
    int *head = S.take<int>(m + n), *bind = S.take<int>(m + n);
    signed char *stat = S.take<signed char>(n);
    for (int i = 0; i < m; i++) head[i] = P->h_head[i] - 1;
    int kk = 0;
    for (int k = 0; k < m + n; k++)
        if (P->h_stat[k] != GLP_BS) {
            if (kk >= n) return GLP_EBADB;
            head[m + kk] = k; stat[kk] = (signed char)P->h_stat[k]; kk++;
        }
    if (kk != n) return GLP_EBADB;
    for (int t = 0; t < m + n; t++) bind[head[t]] = t;
    int rc;
    if ((rc = h2d(P, P->head, head, m + n))) return rc;
    if ((rc = h2d(P, P->bind, bind, m + n))) return rc;
    if ((rc = h2d(P, P->stat, stat, n))) return rc;
    return 0;
}

/* glp_factorize, lib/glpapi12.js:5-100 */
extern "C" int glpb_factorize(glpb_prob *P)
{
    if (!P) return GLPB_EINVAL;
    CK(cudaSetDevice(P->device));
    const int m = P->m, n = P->n;
    P->valid = 0;
    P->t_ok = 0;
    int j = 0;
    for (int k = 0; k < m + n; k++)
        if (P->h_stat[k] == GLP_BS) {
            j++;
            if (j > m) return GLP_EBADB;
            P->h_head[j - 1] = k + 1;
        }
    if (j < m) return GLP_EBADB;
    int rc = upload_basis(P);
    if (rc) return rc;
    Dev D(P);
    rc = dev_refactor(D);
    if (rc) return rc;
    P->valid = 1;
    P->t_ok = 1;
    return 0;
}

/* scaled working copies of bounds and costs (init_csa, lib/glpspx01.js:61-94) */
static int upload_bounds(glpb_prob *P, bool dual)
{
    const int m = P->m, n = P->n;
    Stage S(P, stage_off_bounds(P));
    double *lb = S.take<double>(m + n), *ub = S.take<double>(m + n), *coef = S.take<double>(m + n);
    double *obj = S.take<double>(n + 1);
    signed char *type = S.take<signed char>(m + n);
    for (int i = 0; i < m; i++) {
        type[i] = (signed char)P->h_type[i];
        lb[i] = P->h_lb[i] * P->h_rii[i]; ub[i] = P->h_ub[i] * P->h_rii[i];
        coef[i] = 0.0;
    }
    double cmax = 0.0;
    obj[0] = P->c0;
    for (int j = 0; j < n; j++) {
        type[m + j] = (signed char)P->h_type[m + j];
        lb[m + j] = P->h_lb[m + j] / P->h_sjj[j]; ub[m + j] = P->h_ub[m + j] / P->h_sjj[j];
        coef[m + j] = P->h_coef[j] * P->h_sjj[j];
        obj[1 + j] = coef[m + j];
        cmax = std::max(cmax, fabs(obj[1 + j]));
    }
    if (cmax == 0.0) cmax = 1.0;
    P->zeta = (P->dir == GLP_MIN ? +1.0 : -1.0) / cmax;
    if (fabs(P->zeta) < 1.0) P->zeta *= 1000.0;
    if (dual) for (int j = 0; j < n; j++) coef[m + j] *= P->zeta; /* lib/glpspx02.js:148 */
    int rc;
    if ((rc = h2d(P, P->type, type, m + n))) return rc;
    if ((rc = h2d(P, P->lb, lb, m + n))) return rc;
    if ((rc = h2d(P, P->ub, ub, m + n))) return rc;
    if ((rc = h2d(P, P->orig_type, type, m + n))) return rc;
    if ((rc = h2d(P, P->orig_lb, lb, m + n))) return rc;
    if ((rc = h2d(P, P->orig_ub, ub, m + n))) return rc;
    if ((rc = h2d(P, P->coef, coef, m + n))) return rc;
    if ((rc = h2d(P, P->obj, obj, n + 1))) return rc;
    CK(cudaMemsetAsync(P->refsp, 0, m + n, P->stream));
    return 0;
}

/* store_sol, lib/glpspx01.js:1591-1681: bring head/stat/bbar/cbar back and
   un-scale into the glp_prob fields the getters read */
/* read-back of store_sol enqueued ahead of a synchronisation the caller is about to make
   anyway (speculative: the caller may find that the solve goes on and drop it) */
static int store_sol_prefetch(glpb_prob *P)
{
    const int m = P->m, n = P->n;
    Stage S(P, 0);                /* same layout as store_sol */
    int *head = S.take<int>(m + n);
    double *bbar = S.take<double>(m), *cbar = S.take<double>(n);
    signed char *stat = S.take<signed char>(n);
    CK(cudaMemcpyAsync(head, P->head, (m + n) * sizeof(int), cudaMemcpyDeviceToHost, P->stream));
    CK(cudaMemcpyAsync(stat, P->stat, n, cudaMemcpyDeviceToHost, P->stream));
    CK(cudaMemcpyAsync(bbar, P->bbar, m * sizeof(double), cudaMemcpyDeviceToHost, P->stream));
    CK(cudaMemcpyAsync(cbar, P->cbar, n * sizeof(double), cudaMemcpyDeviceToHost, P->stream));
    return 0;
}

static int store_sol(glpb_prob *P, int p_stat, int d_stat, int ray, int it_cnt, bool prefetched = false)
{
    const int m = P->m, n = P->n;
    Stage S(P, 0);                /* the uploads of this solve completed long ago */
    int *head = S.take<int>(m + n);
    double *bbar = S.take<double>(m), *cbar = S.take<double>(n);
    signed char *stat = S.take<signed char>(n);
    if (!prefetched) {
        CK(cudaMemcpyAsync(head, P->head, (m + n) * sizeof(int), cudaMemcpyDeviceToHost, P->stream));
        CK(cudaMemcpyAsync(stat, P->stat, n, cudaMemcpyDeviceToHost, P->stream));
        CK(cudaMemcpyAsync(bbar, P->bbar, m * sizeof(double), cudaMemcpyDeviceToHost, P->stream));
        CK(cudaMemcpyAsync(cbar, P->cbar, n * sizeof(double), cudaMemcpyDeviceToHost, P->stream));
        CK(cudaStreamSynchronize(P->stream));
        P->n_sync++;
    }
    P->valid = 1;
    P->t_ok = 1;
    P->pbs_stat = p_stat; P->dbs_stat = d_stat;
    P->it_cnt = it_cnt;
    P->some = ray;
    auto nbval = [&](int k, int st) {
        switch (st) {
        case GLP_NL: return P->h_lb[k];
        case GLP_NU: return P->h_ub[k];
        case GLP_NF: return 0.0;
        default: return P->h_lb[k];
        }
    };
    for (int i = 0; i < m; i++) {
        int k = head[i];
        P->h_head[i] = k + 1;
        P->h_stat[k] = GLP_BS;
        P->h_prim[k] = (k < m) ? bbar[i] / P->h_rii[k] : bbar[i] * P->h_sjj[k - m];
        P->h_dual[k] = 0.0;
    }
    for (int j = 0; j < n; j++) {
        int k = head[m + j];
        P->h_stat[k] = stat[j];
        P->h_prim[k] = nbval(k, stat[j]);
        P->h_dual[k] = (k < m) ? (cbar[j] * P->h_rii[k]) / P->zeta : (cbar[j] / P->h_sjj[k - m]) / P->zeta;
    }
    /* eval_obj, lib/glpspx01.js:1524-1548, in the reference's summation order */
    double sum = P->c0;
    for (int i = 0; i < m; i++) {
        int k = head[i];
        if (k >= m) sum += (P->h_coef[k - m] * P->h_sjj[k - m]) * bbar[i];
    }
    for (int j = 0; j < n; j++) {
        int k = head[m + j];
        if (k >= m) {
            double xs;
            switch (stat[j]) {
            case GLP_NL: xs = P->h_lb[k] / P->h_sjj[k - m]; break;
            case GLP_NU: xs = P->h_ub[k] / P->h_sjj[k - m]; break;
            case GLP_NF: xs = 0.0; break;
            default: xs = P->h_lb[k] / P->h_sjj[k - m];
            }
            sum += (P->h_coef[k - m] * P->h_sjj[k - m]) * xs;
        }
    }
    P->obj_val = sum;
    return 0;
}

static int spx_fail(glpb_prob *P, int it_cnt)
{
    P->pbs_stat = P->dbs_stat = GLP_UNDEF;
    P->obj_val = 0.0;
    P->it_cnt = it_cnt;
    P->some = 0;
    P->valid = 0;
    P->t_ok = 0;
    return GLP_EFAIL;
}

/* ------------------------------------------------------------------ */
/* shared pieces of both loops                                        */
/* ------------------------------------------------------------------ */

struct Loop : Dev {
    const glpb_smcp &parm;
    int phase = 0, binv_st = 2, bbar_st = 0, cbar_st = 0, rigorous = 0;
    int it_cnt = 0, it_beg = 0, refct = 0, upd_cnt = 0;
    double tm_beg = 0.0;
    Loop(glpb_prob *P_, const glpb_smcp &parm_) : Dev(P_), parm(parm_) {}

    int grid1(int len) const { return std::max(1, std::min(cdiv(len, 256), 1184)); }

    void clear_ctrl() { LAUNCH(P, k_clear_ctrl, 1, 1, 0, P->ctrl, phase); }

    /* Run `body` (kernel launches only, no host decision inside) through a CUDA graph
       captured on first use.  Grids are sized for the largest kernel (k = m): every kernel
       takes the live k from the device control block, so one capture serves all bases.
       The graph is keyed by the T pointer (the refactorisation flips T/T2) and one scalar. */
    template <class F> void graphed(int kind, double key, F body)
    {
        if (!P->graph_ok || P->prof || P->capturing || P->trace) { body(); return; }
        glpb_prob::GraphSlot *s = nullptr;
        for (auto &g : P->graphs[kind]) if (g.exec && g.T == (const void *)P->T && g.key == key) s = &g;
        if (!s) {
            for (auto &g : P->graphs[kind]) if (!g.exec) { s = &g; break; }
            if (!s) { s = &P->graphs[kind][0]; cudaGraphExecDestroy(s->exec); s->exec = nullptr; }
            const int k_save = k, kpad_save = kpad;
            const long l0 = P->n_launch;
            cudaGraph_t g = nullptr;
            k = m; kpad = 0;
            P->capturing = 1;
            cudaError_t e = cudaStreamBeginCapture(P->stream, cudaStreamCaptureModeThreadLocal);
            if (e == cudaSuccess) {
                body();
                e = cudaStreamEndCapture(P->stream, &g);
            }
            P->capturing = 0;
            k = k_save; kpad = kpad_save;
            s->launches = (int)(P->n_launch - l0);
            P->n_launch = l0;
            if (e == cudaSuccess) e = cudaGraphInstantiate(&s->exec, g, 0);
            if (g) cudaGraphDestroy(g);
            if (e != cudaSuccess) {         /* no graphs on this handle: plain launches from now on */
                cudaGetLastError();
                s->exec = nullptr;
                P->graph_ok = 0;
                body();
                return;
            }
            s->T = (const void *)P->T;
            s->key = key;
        }
        if (cudaGraphLaunch(s->exec, P->stream) != cudaSuccess) { cudaGetLastError(); P->graph_ok = 0; body(); return; }
        P->n_launch += s->launches;
        P->n_graph++;
    }

    /* vectors up to 64k entries are reduced by one block of 1024 threads (no
       second stage); longer ones by a multi-block two-stage reduction */
    int red_blocks(int len) const { return len <= 65536 ? 1 : grid1(len); }
    int red_threads(int len) const { return len <= 65536 ? 1024 : 256; }

    int batch_size() const
    {
        static const int env = getenv("GLPB_BATCH") ? std::max(1, atoi(getenv("GLPB_BATCH"))) : 16;
        return (rigorous > 0 || P->trace || tie_resume) ? 1 : env;
    }

    /* Exact ties of a ratio test are settled in the order of the reference's sort_tcol / sort_trow
       list (k_sort_list).  The per-kernel path always scans that list; the engines scan the dense
       vector, flag an exact tie (ST_TIE) and hand that one iteration over.  GLPB_TIES=0: ties go to
       the lowest index everywhere (no list, no hand-over). */
    static bool list_order()
    {
        static const bool off = getenv("GLPB_TIES") && atoi(getenv("GLPB_TIES")) == 0;
        return !off;
    }
    bool tie_resume = false;      /* the next iteration repeats the one the engine stopped in */

    /* seed the device-resident loop state; the dense rhs of eval_tcol is
       cleared here because an aborted iteration may have left a column in it */
    void batch_begin(int B, double obj_ll, double obj_ul)
    {
        kpad = B;
        const int it_max = parm.it_lim < INT_MAX ? it_beg + parm.it_lim : INT_MAX;
        cudaMemsetAsync(P->w5, 0, (size_t)m * sizeof(double), P->stream);
        LAUNCH(P, k_batch_begin, 1, 1, 0, P->ctrl, phase, it_cnt, it_max, refct, upd_cnt, refac_period(),
               rigorous, bbar_st == 1, cbar_st == 1, binv_st == 1, parm.pricing == GLP_PT_PSE,
               obj_ll, obj_ul, P->zeta);
    }

    bool engine_ok() const
    {
        static const bool off = getenv("GLPB_ENGINE") && atoi(getenv("GLPB_ENGINE")) == 0;
        return !off && P->eng_ready && rigorous == 0 && !P->trace && !tie_resume;
    }

    /* grid of the persistent engine: every SM once the problem is big enough to
       feed them, a single CTA (barriers degrade to __syncthreads) for tiny LPs */
    int engine_grid() const
    {
        static const int env = getenv("GLPB_GRID") ? atoi(getenv("GLPB_GRID")) : 0;
        if (env > 0) return std::min(env, P->sm_count);
        double work = (double)P->nnz + (double)k * k + 4.0 * (m + n);
        if (work <= 32768.0) return 1;
        return (int)std::min<double>(P->sm_count, std::max(2.0, work / 8192.0));
    }

    /* run iterations on the device until the engine meets one of the
       reference's exceptional branches (or max_iters); one host sync */
    int run_engine(int dual, double rtol, double obj_ll, double obj_ul)
    {
        static const int iters = getenv("GLPB_ENG_ITERS") ? std::max(1, atoi(getenv("GLPB_ENG_ITERS"))) : 2000;
        batch_begin(0, obj_ll, obj_ul);
        EngArgs A;
        A.ctrl = P->ctrl; A.m = m; A.n = n; A.ldt = P->ldt; A.max_iters = iters;
        /* smcp.tm_lim is tested by the host between launches: with a time limit set, a launch is kept short
           (64 iterations: ~10 ms at the C3 size) so that the overshoot stays bounded */
        if (parm.tm_lim < INT_MAX) A.max_iters = std::min(iters, 64);
        A.dcap = P->eng_dcap;
        A.avg_col = (double)P->nnz / n; A.avg_row = (double)P->nnz / m;
        A.tol_bnd = parm.tol_bnd; A.tol_dj = parm.tol_dj; A.tol_piv = parm.tol_piv; A.rtol = rtol;
        A.a_ptr = P->a_ptr; A.a_ind = P->a_ind; A.a_val = P->a_val;
        A.at_ptr = P->at_ptr; A.at_ind = P->at_ind; A.at_val = P->at_val;
        A.type = P->type; A.stat = P->stat; A.refsp = P->refsp; A.orig_type = P->orig_type;
        A.lb = P->lb; A.ub = P->ub; A.coef = P->coef;
        A.head = P->head; A.bind = P->bind;
        A.bbar = P->bbar; A.cbar = P->cbar; A.gamma = P->gamma; A.tcol = P->tcol; A.trow = P->trow;
        A.rho = P->rho; A.svec = P->svec;
        A.hz = P->w5; A.v = P->w3; A.u = P->w2;
        A.yk = P->yk; A.yk2 = P->yk2; A.wk = P->wk; A.zn = P->zn;
        A.T = P->T;
        A.rslot = P->rslot; A.slot_pos = P->slot_pos; A.cslot = P->cslot; A.slot_row = P->slot_row;
        A.ycol = P->eng_cols; A.ycol2 = P->eng_cols + n; A.trowcol = P->eng_cols + 2 * (size_t)n;
        A.vrow = P->eng_cols + 3 * (size_t)n;
        A.slots = P->eng_slots;
        static const int env_defer = getenv("GLPB_DEFER") ? atoi(getenv("GLPB_DEFER")) : 1;
        static const int env_local = getenv("GLPB_LOCAL_MAX") ? atoi(getenv("GLPB_LOCAL_MAX")) : ENG_LOCAL_MAX;
        A.Fd = P->eng_fr; A.Rd = P->eng_fr + (size_t)ENG_DB * P->ldt; A.zbuf = P->eng_fr + 2 * (size_t)ENG_DB * P->ldt;
        A.defer = (env_defer && dual && P->eng_dcap >= ENG_FLUSH_SMEM) ? 1 : 0;
        /* measured on B200 (C3): per-column bulk copies of ~350 bytes are issue-bound (2.2 TB/s) and lose
           to the register-staged 16-byte loads (3.6 TB/s), so the TMA-staged stream is opt-in (GLPB_TMA=1) */
        A.use_tma = P->eng_tma;
        A.hdr_smem = P->eng_hdr;
        A.local_max = env_local;
        A.tie_stop = list_order() ? 1 : 0;
        static const int env_pf = getenv("GLPB_PF") ? std::max(0, atoi(getenv("GLPB_PF"))) : 2;
        A.pf_dist = env_pf;
        static const int env_pff = getenv("GLPB_PF_FIRST") ? std::max(0, atoi(getenv("GLPB_PF_FIRST"))) : 0;
        A.pf_first = env_pff;
        static const int env_async = getenv("GLPB_ASYNC") ? atoi(getenv("GLPB_ASYNC")) : 0;
        A.use_async = env_async;
        A.prof_cyc = P->prof ? P->eng_cyc + (dual ? 12 : 0) : nullptr;
        A.prof_bytes = P->prof ? P->eng_bytes + (dual ? 12 : 0) : nullptr;
        const int G = engine_grid();
        CK(cudaMemsetAsync(P->eng_slots, 0, (ENG_RING * ENG_MAXG + 1) * sizeof(EngSlot), P->stream));
        CK(cudaMemsetAsync(P->eng_cols, 0, (3 * (size_t)n + m) * sizeof(double), P->stream));
        void *args[] = {&A};
        prof_begin(P, dual ? "k_engine_dual" : "k_engine_primal");
        const void *kfn = P->eng_hdr ? (dual ? (const void *)k_engine_dual<true> : (const void *)k_engine_primal<true>)
                                     : (dual ? (const void *)k_engine_dual<false> : (const void *)k_engine_primal<false>);
        cudaError_t e = cudaLaunchCooperativeKernel(kfn,
                                                    dim3(G), dim3(ENG_NT), args, (size_t)P->eng_smem, P->stream);
        prof_end(P);
        if (e == cudaErrorCooperativeLaunchTooLarge) {
            /* the SMs are shared with somebody else right now: this handle goes on with the
               per-kernel path (same kernels' arithmetic, host-enqueued) */
            cudaGetLastError();
            P->eng_ready = 0;
            return sync_ctrl(P);        /* status is ST_OK with n_done = 0: the caller loops */
        }
        if (e != cudaSuccess) { glpb_set_error("engine launch: %s", cudaGetErrorString(e)); return GLPB_ENODEV; }
        P->n_launch++; P->n_eng_launch++;
        int rc = sync_ctrl(P);
        if (rc == 0 && P->prof) P->n_eng_prof_iter[dual ? 1 : 0] += P->h_ctrl->n_done;
        /* GLPB_ENG_LOG=file: one line per engine launch (ordinal, iterations, kernel size) so that an ncu
           capture of launch number i can be put per iteration */
        static const char *logf = getenv("GLPB_ENG_LOG");
        if (rc == 0 && logf) {
            if (FILE *f = fopen(logf, "a")) {
                fprintf(f, "%ld %s iters %d k %d status %d\n", (long)P->n_eng_launch, dual ? "dual" : "primal",
                        P->h_ctrl->n_done, P->h_ctrl->k, P->h_ctrl->status);
                fclose(f);
            }
        }
        return rc;
    }

    /* take the loop state back after the batch */
    void batch_end()
    {
        const Ctrl &c = *P->h_ctrl;
        k = c.k;
        kpad = 0;
        P->n_iter += c.n_done;
        P->n_update += c.upd_cnt - upd_cnt;
        it_cnt = c.it_cnt; upd_cnt = c.upd_cnt; refct = c.refct; rigorous = c.rigorous;
        bbar_st = c.bbar_fresh ? 1 : 2;
        cbar_st = c.cbar_fresh ? 1 : 2;
        binv_st = c.binv_fresh ? 1 : 2;
    }

    /* eval_bbar, lib/glpspx01.js:473-512,560-563 */
    void eval_bbar() { graphed(glpb_prob::GK_BBAR, 0.0, [&] { eval_bbar_launches(); }); }
    void eval_bbar_launches()
    {
        GROUP_DISPATCH(gr, LAUNCH(P, k_beta_rhs<GG>, cdiv((long)m * GG, 256), 256, 0, P->ctrl, m, n, P->at_ptr,
                                  P->at_ind, P->at_val, P->head, P->bind, P->stat, P->lb, P->ub, P->w2));
        dev_ftran(*this, P->w2, P->bbar);
        GROUP_DISPATCH(gr, LAUNCH(P, k_resid_ftran<GG>, cdiv((long)m * GG, 256), 256, 0, P->ctrl, m, P->at_ptr,
                                  P->at_ind, P->at_val, P->bind, P->w2, P->bbar, P->w1));
        dev_ftran(*this, P->w1, P->w3);
        LAUNCH(P, k_axpy1, cdiv(m, 256), 256, 0, m, P->bbar, P->w3);
    }

    /* eval_cbar, lib/glpspx01.js:514-584 */
    void eval_cbar() { graphed(glpb_prob::GK_CBAR, 0.0, [&] { eval_cbar_launches(); }); }
    void eval_cbar_launches()
    {
        LAUNCH(P, k_gather_cB, cdiv(m, 256), 256, 0, m, P->head, P->coef, P->w2);
        dev_btran(*this, P->w2, P->w3);
        GROUP_DISPATCH(gc, LAUNCH(P, k_resid_btran<GG>, cdiv((long)m * GG, 256), 256, 0, P->ctrl, m, P->a_ptr,
                                  P->a_ind, P->a_val, P->head, P->w2, P->w3, P->w1));
        dev_btran(*this, P->w1, P->w4);
        LAUNCH(P, k_axpy1, cdiv(m, 256), 256, 0, m, P->w3, P->w4);
        GROUP_DISPATCH(gc, LAUNCH(P, k_cbar<GG>, cdiv((long)n * GG, 256), 256, 0, P->ctrl, m, n, P->a_ptr,
                                  P->a_ind, P->a_val, P->head, P->coef, P->w3, P->cbar));
    }

    /* eval_tcol (+ refine_tcol), lib/glpspx01.js:690-771 */
    void eval_tcol()
    {
        LAUNCH(P, k_col_rhs, 4, 256, 0, P->ctrl, m, P->a_ptr, P->a_ind, P->a_val, P->head, P->w5);
        dev_ftran(*this, P->w5, P->tcol);
        if (rigorous) {
            GROUP_DISPATCH(gr, LAUNCH(P, k_resid_ftran<GG>, cdiv((long)m * GG, 256), 256, 0, P->ctrl, m, P->at_ptr,
                                      P->at_ind, P->at_val, P->bind, P->w5, P->tcol, P->w1));
            dev_ftran(*this, P->w1, P->w4);
            LAUNCH(P, k_axpy1, cdiv(m, 256), 256, 0, m, P->tcol, P->w4);
        }
    }

    /* eval_rho (+ refine_rho), lib/glpspx01.js:1030-1056 */
    void eval_rho()
    {
        launch_rho(P, k + kpad);
        if (rigorous) {
            LAUNCH(P, k_unit, cdiv(m, 256), 256, 0, P->ctrl, m, P->w2);
            GROUP_DISPATCH(gc, LAUNCH(P, k_resid_btran<GG>, cdiv((long)m * GG, 256), 256, 0, P->ctrl, m, P->a_ptr,
                                      P->a_ind, P->a_val, P->head, P->w2, P->rho, P->w1));
            dev_btran(*this, P->w1, P->w4);
            LAUNCH(P, k_axpy1, cdiv(m, 256), 256, 0, m, P->rho, P->w4);
        }
    }

    void update_basis(int dual)
    {
        const int kg = k + kpad;
        if (kg > 0) {
            dim3 grid(cdiv(kg, UPD_TB), cdiv(kg, UPD_TC));
            P->next_bytes = 16.0 * k * (double)k;
            LAUNCH(P, k_update_rank1, grid, UPD_TB, 0, P->ctrl, P->T, P->ldt, P->tcol, P->rho, P->slot_pos, P->slot_row);
        }
        LAUNCH(P, k_update_fix, 1, 1024, 0, P->ctrl, m, P->T, P->ldt, P->tcol, P->rho, P->rslot, P->slot_pos,
               P->cslot, P->slot_row, P->head, P->bind, P->stat, P->type, P->refsp, dual);
    }

    int refactor()
    {
        int rc = dev_refactor(*this);
        upd_cnt = 0;
        return rc;
    }

    /* refactorisation period: bfcp.nfs_max (the reference's eta-file limit, default 100);
       the explicit inverse needs a fresh start only for accuracy, so for large
       kernels the period grows with k -- max(nfs_max, 4k) -- unless GLPB_REFAC_AUTO=0 */
    int refac_period() const
    {
        static const bool fixed = getenv("GLPB_REFAC_AUTO") && atoi(getenv("GLPB_REFAC_AUTO")) == 0;
        /* k updates between refactorisations (measured on C3: 27 instead of 53 refactorisations, -1.5 s of
           22.3 s, same 120552 iterations, objective to 5e-14 of the HiGHS pin, KKT 5e-16; the accuracy triggers
           of the reference -- piv1/piv2 at 1e-8, d1/d2, check_stab -- still force one when needed) */
        static const int div = getenv("GLPB_REFAC_DIV") ? std::max(1, atoi(getenv("GLPB_REFAC_DIV"))) : 1;
        /* round 2, same box (profiles/r03_ab_refac_mul.txt): periods k / 2k / 4k give 27 / 14 / 8 refactorisations
           and 18.64 / 17.94 / 17.62 s on C3, 415 / 392 / 381 ms on C2, the SAME 120552 resp. 6728 iterations,
           objective to 5e-14 of the pin, KKT 4e-16 */
        static const int mul = getenv("GLPB_REFAC_MUL") ? std::max(1, atoi(getenv("GLPB_REFAC_MUL"))) : 4;
        if (fixed || k / div <= P->bfcp.nfs_max) return P->bfcp.nfs_max;   /* small kernels: the reference's limit */
        return (int)std::min<long>((long)k * mul / div, 1 << 28);
    }
    bool it_limit() const { return parm.it_lim < INT_MAX && it_cnt - it_beg >= parm.it_lim; }
    bool tm_limit() const { return parm.tm_lim < INT_MAX && (now_ms() - tm_beg) >= parm.tm_lim; }

    void trace_iter(const char *tag)
    {
        if (!P->trace) return;
        const Ctrl &c = *P->h_ctrl;
        fprintf(stderr, "[glpb %s] it %d ph %d st %d q %d p %d k %d teta %.12g delta %.12g piv1 %.6g piv2 %.6g rig %d\n",
                tag, it_cnt, phase, c.status, c.q, c.p, c.k, c.teta, c.delta, c.piv1, c.piv2, rigorous);
    }
};

/* ------------------------------------------------------------------ */
/* primal loop controller: lib/glpspx01.js:1684-2056                  */
/* ------------------------------------------------------------------ */

struct Primal : Loop {
    using Loop::Loop;

    void chuzc(int set_status)
    {
        P->next_bytes = 17.0 * n;
        LAUNCH(P, k_chuzc_primal, red_blocks(n), red_threads(n), 0, P->ctrl, n, P->stat, P->cbar, P->gamma,
               parm.tol_dj, set_status, P->scratch);
    }

    int set_aux_obj(int *cnt)
    {
        clear_ctrl();
        LAUNCH(P, k_set_aux_obj, cdiv(m + n, 256), 256, 0, P->ctrl, m, n, P->head, P->type, P->lb, P->ub,
               P->bbar, P->coef, 0.90 * parm.tol_bnd);
        int rc = sync_ctrl(P);
        *cnt = P->h_ctrl->cnt;
        return rc;
    }

    void set_orig_obj()
    {
        LAUNCH(P, k_set_orig_obj, cdiv(m + n, 256), 256, 0, m, n, P->obj, P->coef, P->zeta);
    }

    int check(int mode, int *flag)
    {
        clear_ctrl();
        LAUNCH(P, k_primal_check, cdiv(m, 256), 256, 0, P->ctrl, m, mode, phase, P->head, P->type, P->lb,
               P->ub, P->coef, P->bbar, parm.tol_bnd);
        int rc = sync_ctrl(P);
        *flag = P->h_ctrl->flag;
        return rc;
    }

    /* returns q != none after the run-out checks of lib/glpspx01.js:1813-1832 */
    int stop_on_limit(int code)
    {
        int p_stat, d_stat;
        clear_ctrl();
        if (phase == 1) { p_stat = GLP_INFEAS; set_orig_obj(); eval_cbar(); }
        else p_stat = GLP_FEAS;
        chuzc(0);
        int rc = sync_ctrl(P);
        if (rc) return rc;
        d_stat = (P->h_ctrl->q == P_NONE ? GLP_FEAS : GLP_INFEAS);
        rc = store_sol(P, p_stat, d_stat, 0, it_cnt);
        return rc ? rc : code;
    }

    int run()
    {
        int rc, flag, cnt;
        const int pse = (parm.pricing == GLP_PT_PSE);
        if ((rc = upload_bounds(P, false))) return rc;
        if ((rc = upload_basis(P))) return rc;
        k = P->k_host;      /* T and its slot maps survive between calls */
        P->valid = 0;
        P->t_ok = 0;         /* the device state is in motion until store_sol */
        it_beg = it_cnt = P->it_cnt;
        tm_beg = now_ms();
        refct = 0;
        LAUNCH(P, k_reset_refsp, cdiv(m + n, 256), 256, 0, m, n, P->head, P->refsp, P->gamma, 0);
        cudaMemsetAsync(P->refsp, 0, m + n, P->stream);
        for (;;) {
            if (binv_st == 0) {
                rc = refactor();
                if (rc < 0) return rc;
                if (rc != 0) return spx_fail(P, it_cnt);
                binv_st = 1;
                bbar_st = cbar_st = 0;
            }
            if (bbar_st == 0) {
                clear_ctrl();
                eval_bbar();
                bbar_st = 1;
                if (phase == 0) {
                    if ((rc = set_aux_obj(&cnt))) return rc;
                    if (cnt > 0) phase = 1;
                    else { set_orig_obj(); phase = 2; }
                    cbar_st = 0;
                }
                if ((rc = check(0, &flag))) return rc;
                if (flag) { phase = 0; binv_st = 0; rigorous = 5; continue; }
            }
            if (phase == 1) {
                if ((rc = check(1, &flag))) return rc;
                if (!flag) { phase = 2; set_orig_obj(); cbar_st = 0; }
            }
            if (cbar_st == 0) { clear_ctrl(); eval_cbar(); cbar_st = 1; }
            if (pse && refct == 0) {
                LAUNCH(P, k_reset_refsp, cdiv(m + n, 256), 256, 0, m, n, P->head, P->refsp, P->gamma, 0);
                refct = 1000;
            }
            if (it_limit() || tm_limit()) {
                int code = it_limit() ? GLP_EITLIM : GLP_ETMLIM;
                if (bbar_st != 1 || (phase == 2 && cbar_st != 1)) {
                    if (bbar_st != 1) bbar_st = 0;
                    if (phase == 2 && cbar_st != 1) cbar_st = 0;
                    continue;
                }
                return stop_on_limit(code);
            }
            /* ---- a batch of iterations, enqueued as a whole; the device keeps the
                    loop state and turns the rest of the queue into no-ops as soon
                    as one of the reference's exceptional branches is met ---- */
            const double rtol = (parm.r_test == GLP_RT_STD ? 0.0 : 0.30 * parm.tol_bnd);
            const bool eng = engine_ok();
            if (eng) { if ((rc = run_engine(0, rtol, -DBL_MAX, +DBL_MAX))) return rc; }
            const int B = eng ? 0 : batch_size();
            if (!eng) batch_begin(B, -DBL_MAX, +DBL_MAX);
            for (int b = 0; b < B; b++) {
                /* an iteration taken over from the engine keeps the engine's q: the engine has already
                   replaced cbar[q] by the recomputed value (reeval_cost), pricing again could differ */
                if (!tie_resume) chuzc(1);
                eval_tcol();
                LAUNCH(P, k_primal_prep, red_blocks(m), red_threads(m), 0, P->ctrl, m, P->head, P->coef, P->tcol,
                       P->refsp, P->cbar, P->w3, parm.tol_piv, P->scratch);
                /* sort_tcol: the significant entries in the reference's list order, so that exact ties
                   of the ratio test fall the way they do there (num = -1: count left on the device) */
                const int *lst = list_order() ? P->sort_list : nullptr;
                if (lst) LAUNCH(P, k_sort_list, 1, 1024, 0, P->ctrl, P->tcol, m, P->sort_list, P->sort_tmp);
                if (red_blocks(m) == 1) {
                    P->next_bytes = 90.0 * m;
                    LAUNCH(P, k_ratio_primal, 1, 1024, 0, P->ctrl, 0, m, P->type, P->lb, P->ub, P->coef, P->head,
                           P->bbar, P->tcol, lst, lst ? -1 : m, rtol, P->scratch);
                } else
                    for (int pass = 1; pass <= 2; pass++) {
                        P->next_bytes = 45.0 * m;
                        LAUNCH(P, k_ratio_primal, grid1(m), 256, 0, P->ctrl, pass, m, P->type, P->lb, P->ub, P->coef,
                               P->head, P->bbar, P->tcol, lst, lst ? -1 : m, rtol, P->scratch);
                    }
                eval_rho();
                if (pse) dev_btran(*this, P->w3, P->w2, 1);
                P->next_bytes = 12.0 * P->nnz * (1.0 - (double)k / n) + 13.0 * n + 8.0 * m;
                GROUP_DISPATCH(gc, LAUNCH(P, k_trow<GG>, cdiv((long)n * GG, 256), 256, 0, P->ctrl, m, n, P->a_ptr,
                                          P->a_ind, P->a_val, P->head, P->stat, P->rho,
                                          pse ? P->w2 : (const double *)nullptr, P->trow, P->svec, 0));
                LAUNCH(P, k_primal_piv, 1, 1, 0, P->ctrl, m, P->head, P->trow, P->cbar, P->coef, P->tcol);
                LAUNCH(P, k_primal_update, cdiv(std::max(m, n), 256), 256, 0, P->ctrl, m, n, P->head, P->stat,
                       P->type, P->lb, P->ub, P->coef, P->bbar, P->cbar, P->gamma, P->refsp, P->tcol, P->trow,
                       P->svec, P->w5);
                update_basis(0);
                if (phase == 1)
                    LAUNCH(P, k_phase1_stop, 1, 1024, 0, P->ctrl, 0, m, n, P->head, P->orig_type, P->lb, P->ub, P->coef,
                           P->bbar, P->cbar, parm.tol_bnd);
            }
            if (!eng && (rc = sync_ctrl(P))) return rc;
            trace_iter("primal");
            const Ctrl &c = *P->h_ctrl;
            batch_end();
            tie_resume = false;
            switch (c.status) {
            case ST_OK: case ST_LIMIT: case ST_REFSP: case ST_PHASE:
                break;
            case ST_REFAC:
                binv_st = 0;
                break;
            case ST_TIE:
                tie_resume = true;      /* this iteration again, through the list-ordered ratio test */
                P->n_tie++;
                break;
            case ST_NONE1:
                if (bbar_st != 1 || cbar_st != 1) {
                    if (bbar_st != 1) bbar_st = 0;
                    if (cbar_st != 1) cbar_st = 0;
                    break;
                }
                {
                    int p_stat, d_stat;
                    if (phase == 1) {
                        p_stat = GLP_NOFEAS;
                        set_orig_obj();
                        clear_ctrl();
                        eval_cbar();
                        chuzc(0);
                        if ((rc = sync_ctrl(P))) return rc;
                        d_stat = (P->h_ctrl->q == P_NONE ? GLP_FEAS : GLP_INFEAS);
                    } else
                        p_stat = d_stat = GLP_FEAS;
                    rc = store_sol(P, p_stat, d_stat, 0, it_cnt);
                    return rc;
                }
            case ST_D1D2:
                if (cbar_st != 1) cbar_st = 0;
                rigorous = 5;
                break;
            case ST_NONE2:
                if (bbar_st != 1 || cbar_st != 1 || !rigorous) {
                    if (bbar_st != 1) bbar_st = 0;
                    if (cbar_st != 1) cbar_st = 0;
                    rigorous = 1;
                    break;
                }
                if (phase == 1) return spx_fail(P, it_cnt);
                {
                    std::vector<int> hq(1);
                    CK(cudaMemcpy(hq.data(), P->head + m + c.q, sizeof(int), cudaMemcpyDeviceToHost));
                    rc = store_sol(P, GLP_FEAS, GLP_NOFEAS, hq[0] + 1, it_cnt);
                    return rc;
                }
            case ST_PIVSMALL:
                rigorous = 5;
                break;
            case ST_PIV12:
                if (binv_st != 1) binv_st = 0;
                rigorous = 5;
                break;
            default:
                glpb_set_error("primal: unexpected device status %d", c.status);
                return GLPB_ESTATE;
            }
        }
    }
};

/* ------------------------------------------------------------------ */
/* dual loop controller: lib/glpspx02.js:1592-1966                    */
/* ------------------------------------------------------------------ */

struct Dual : Loop {
    using Loop::Loop;

    int check(int mode, double tol, int *flag)
    {
        clear_ctrl();
        LAUNCH(P, k_dual_check, cdiv(n, 256), 256, 0, P->ctrl, m, n, mode, P->head, P->orig_type, P->stat,
               P->cbar, tol);
        int rc = sync_ctrl(P);
        *flag = P->h_ctrl->flag;
        return rc;
    }

    void set_bnds(int aux)
    {
        LAUNCH(P, k_dual_set_bnds, cdiv(m + n, 256), 256, 0, P->ctrl, m, n, aux, P->head, P->bind, P->orig_type,
               P->orig_lb, P->orig_ub, P->type, P->lb, P->ub, P->stat, P->cbar);
    }

    void eval_obj()
    {
        LAUNCH(P, k_eval_obj, grid1(m + n), 256, 0, P->ctrl, m, n, P->head, P->obj, P->bbar, P->stat, P->lb,
               P->ub, P->scratch);
    }

    int stop_on_limit(int code)
    {
        int d_stat;
        if (phase == 1) { d_stat = GLP_INFEAS; set_bnds(0); clear_ctrl(); eval_bbar(); }
        else d_stat = GLP_FEAS;
        int rc = store_sol(P, GLP_INFEAS, d_stat, 0, it_cnt);
        return rc ? rc : code;
    }

    int run()
    {
        int rc, flag;
        const int pse = (parm.pricing == GLP_PT_PSE);
        double obj_track = 0.0;
        /* GLPB_FINAL=0: the last round of a solve goes through the generic loop (two
           synchronisations and one more engine launch) */
        static const bool final_round = !(getenv("GLPB_FINAL") && atoi(getenv("GLPB_FINAL")) == 0);
        if ((rc = upload_bounds(P, true))) return rc;
        if ((rc = upload_basis(P))) return rc;
        k = P->k_host;
        P->valid = 0;
        P->t_ok = 0;         /* the device state is in motion until store_sol */
        it_beg = it_cnt = P->it_cnt;
        tm_beg = now_ms();
        refct = 0;
        LAUNCH(P, k_reset_refsp, cdiv(m + n, 256), 256, 0, m, n, P->head, P->refsp, P->gamma, 1);
        cudaMemsetAsync(P->refsp, 0, m + n, P->stream);
        for (;;) {
            if (binv_st == 0) {
                rc = refactor();
                if (rc < 0) return rc;
                if (rc != 0) return spx_fail(P, it_cnt);
                binv_st = 1;
                bbar_st = cbar_st = 0;
            }
            if (cbar_st == 0 && phase == 0) {
                /* start of a solve: phase decision, stability check, bbar and the objective
                   are enqueued as a whole (the device selects the bounds from its own
                   feasibility flag) and read back with ONE synchronisation */
                graphed(glpb_prob::GK_DSTART, parm.tol_dj, [&] {
                    clear_ctrl();
                    eval_cbar();
                    LAUNCH(P, k_dual_check, cdiv(n, 256), 256, 0, P->ctrl, m, n, 1, P->head, P->orig_type, P->stat,
                           P->cbar, 0.90 * parm.tol_dj, 0);
                    set_bnds(-1);
                    LAUNCH(P, k_dual_check, cdiv(n, 256), 256, 0, P->ctrl, m, n, 0, P->head, P->orig_type, P->stat,
                           P->cbar, parm.tol_dj, 1);
                    eval_bbar();
                    eval_obj();
                });
                cbar_st = 1;
                if ((rc = sync_ctrl(P))) return rc;
                phase = P->h_ctrl->flag ? 1 : 2;
                refct = 0;
                if (P->h_ctrl->cnt) {
                    if (parm.meth == GLP_DUALP) {
                        rc = store_sol(P, GLP_UNDEF, GLP_UNDEF, 0, it_cnt);
                        return rc ? rc : GLP_EFAIL;
                    }
                    phase = 0; binv_st = 0; rigorous = 5;
                    continue;
                }
                bbar_st = 1;
                if (phase == 2) obj_track = P->h_ctrl->obj;
            }
            if (cbar_st == 0) {
                clear_ctrl();
                eval_cbar();
                cbar_st = 1;
                if (phase == 0) {
                    if ((rc = check(1, 0.90 * parm.tol_dj, &flag))) return rc;
                    if (flag) { phase = 1; set_bnds(1); }
                    else { phase = 2; set_bnds(0); }
                    refct = 0;
                    bbar_st = 0;
                }
                if ((rc = check(0, parm.tol_dj, &flag))) return rc;
                if (flag) {
                    if (parm.meth == GLP_DUALP) {
                        rc = store_sol(P, GLP_UNDEF, GLP_UNDEF, 0, it_cnt);
                        return rc ? rc : GLP_EFAIL;
                    }
                    phase = 0; binv_st = 0; rigorous = 5;
                    continue;
                }
            }
            if (phase == 1) {
                if ((rc = check(1, parm.tol_dj, &flag))) return rc;
                if (!flag) {
                    phase = 2;
                    if (cbar_st != 1) { clear_ctrl(); eval_cbar(); cbar_st = 1; }
                    set_bnds(0);
                    refct = 0;
                    bbar_st = 0;
                }
            }
            if (bbar_st == 0) {
                clear_ctrl();
                eval_bbar();
                if (phase == 2) {
                    eval_obj();
                    if ((rc = sync_ctrl(P))) return rc;
                    obj_track = P->h_ctrl->obj;
                }
                bbar_st = 1;
            }
            if (pse && refct == 0) {
                LAUNCH(P, k_reset_refsp, cdiv(m + n, 256), 256, 0, m, n, P->head, P->refsp, P->gamma, 1);
                refct = 1000;
            }
            /* objective cut-offs used by branch-and-bound, lib/glpspx02.js:1727-1760 */
            if (phase == 2 && ((P->zeta < 0.0 && parm.obj_ll > -DBL_MAX && obj_track <= parm.obj_ll) ||
                               (P->zeta > 0.0 && parm.obj_ul < +DBL_MAX && obj_track >= parm.obj_ul))) {
                if (bbar_st != 1 || cbar_st != 1) {
                    if (bbar_st != 1) bbar_st = 0;
                    if (cbar_st != 1) cbar_st = 0;
                    continue;
                }
                rc = store_sol(P, GLP_INFEAS, GLP_FEAS, 0, it_cnt);
                return rc ? rc : (P->zeta < 0.0 ? GLP_EOBJLL : GLP_EOBJUL);
            }
            if (it_limit() || tm_limit()) {
                int code = it_limit() ? GLP_EITLIM : GLP_ETMLIM;
                if ((phase == 2 && bbar_st != 1) || cbar_st != 1) {
                    if (phase == 2 && bbar_st != 1) bbar_st = 0;
                    if (cbar_st != 1) cbar_st = 0;
                    continue;
                }
                return stop_on_limit(code);
            }
            /* ---- a batch of iterations (see the primal loop) ---- */
            const double rtol = (parm.r_test == GLP_RT_STD ? 0.0 : 0.30 * parm.tol_dj);
            const bool eng = engine_ok();
            if (eng) { if ((rc = run_engine(1, rtol, parm.obj_ll, parm.obj_ul))) return rc; }
            const int B = eng ? 0 : batch_size();
            if (!eng) batch_begin(B, parm.obj_ll, parm.obj_ul);
            for (int b = 0; b < B; b++) {
                P->next_bytes = 37.0 * m;
                LAUNCH(P, k_chuzr_dual, red_blocks(m), red_threads(m), 0, P->ctrl, m, P->type, P->lb, P->ub, P->head,
                       P->bbar, P->gamma, parm.tol_bnd, 1, P->scratch);
                eval_rho();
                P->next_bytes = 12.0 * P->nnz * (1.0 - (double)k / n) + 13.0 * n + 8.0 * m;
                GROUP_DISPATCH(gc, LAUNCH(P, k_trow<GG>, cdiv((long)n * GG, 256), 256, 0, P->ctrl, m, n, P->a_ptr,
                                          P->a_ind, P->a_val, P->head, P->stat, P->rho, (const double *)nullptr,
                                          P->trow, P->svec, 1));
                LAUNCH(P, k_dual_rowmax, 1, 1, 0, P->ctrl, parm.tol_bnd); /* sic: tol_bnd, lib/glpspx02.js:1851 */
                /* sort_trow: see the primal loop */
                const int *lst = list_order() ? P->sort_list : nullptr;
                if (lst) LAUNCH(P, k_sort_list, 1, 1024, 0, P->ctrl, P->trow, n, P->sort_list, P->sort_tmp);
                if (red_blocks(n) == 1) {
                    P->next_bytes = 34.0 * n;
                    LAUNCH(P, k_ratio_dual, 1, 1024, 0, P->ctrl, 0, P->stat, P->cbar, P->trow, lst,
                           lst ? -1 : n, rtol, P->scratch);
                } else
                    for (int pass = 1; pass <= 2; pass++) {
                        P->next_bytes = 17.0 * n;
                        LAUNCH(P, k_ratio_dual, grid1(n), 256, 0, P->ctrl, pass, P->stat, P->cbar, P->trow,
                               lst, lst ? -1 : n, rtol, P->scratch);
                    }
                eval_tcol();
                LAUNCH(P, k_dual_prep, red_blocks(n), red_threads(n), 0, P->ctrl, m, n, P->head, P->refsp, P->trow,
                       P->tcol, P->cbar, P->stat, P->zeta, P->scratch);
                if (pse) {
                    GROUP_DISPATCH(gr, LAUNCH(P, k_dual_gamma_rhs<GG>, cdiv((long)m * GG, 256), 256, 0, P->ctrl, m,
                                              P->at_ptr, P->at_ind, P->at_val, P->bind, P->refsp, P->trow, P->w3, 1));
                    dev_ftran(*this, P->w3, P->w2, 1);
                }
                LAUNCH(P, k_dual_update, cdiv(std::max(m, n), 256), 256, 0, P->ctrl, m, n, P->head, P->stat, P->type,
                       P->lb, P->ub, P->bbar, P->cbar, P->gamma, P->refsp, P->tcol, P->trow, P->w2, P->w5);
                update_basis(1);
                if (phase == 1)
                    LAUNCH(P, k_phase1_stop, 1, 1024, 0, P->ctrl, 1, m, n, P->head, P->orig_type, P->lb, P->ub, P->coef,
                           P->bbar, P->cbar, parm.tol_dj);
            }
            if (!eng && (rc = sync_ctrl(P))) return rc;
            trace_iter("dual");
            const Ctrl &c = *P->h_ctrl;
            batch_end();
            tie_resume = false;
            obj_track = c.obj;
            switch (c.status) {
            case ST_OK: case ST_LIMIT: case ST_REFSP: case ST_OBJLIM: case ST_PHASE:
                break;
            case ST_REFAC:
                binv_st = 0;
                break;
            case ST_TIE:
                tie_resume = true;      /* this iteration again, through the list-ordered ratio test */
                P->n_tie++;
                break;
            case ST_NONE1:
                if (phase == 2 && bbar_st != 1 && cbar_st != 1 && final_round) {
                    /* Nothing to price, but bbar and cbar are updated, not freshly computed
                       vectors: the reference recomputes both and goes round its loop once more
                       (lib/glpspx02.js:1660-1760,1802-1812).  That round -- eval_cbar, stability
                       check, eval_bbar, objective, pricing -- is enqueued as a whole and read back
                       with ONE synchronisation; the decisions below are taken in the loop's order. */
                    graphed(glpb_prob::GK_DFINAL, parm.tol_dj, [&] {
                        clear_ctrl();
                        eval_cbar();
                        LAUNCH(P, k_dual_check, cdiv(n, 256), 256, 0, P->ctrl, m, n, 0, P->head, P->orig_type,
                               P->stat, P->cbar, parm.tol_dj, 0);
                        eval_bbar();
                        eval_obj();
                    });
                    P->next_bytes = 37.0 * m;
                    LAUNCH(P, k_chuzr_dual, red_blocks(m), red_threads(m), 0, P->ctrl, m, P->type, P->lb, P->ub,
                           P->head, P->bbar, P->gamma, parm.tol_bnd, 1, P->scratch);
                    if ((rc = store_sol_prefetch(P))) return rc;   /* the solution rides on the same sync */
                    if ((rc = sync_ctrl(P))) return rc;
                    bbar_st = cbar_st = 1;
                    obj_track = P->h_ctrl->obj;
                    if (P->h_ctrl->flag) {              /* dual feasibility lost: as at the top of the loop */
                        if (parm.meth == GLP_DUALP) {
                            rc = store_sol(P, GLP_UNDEF, GLP_UNDEF, 0, it_cnt);
                            return rc ? rc : GLP_EFAIL;
                        }
                        phase = 0; binv_st = 0; rigorous = 5;
                        break;
                    }
                    if (P->h_ctrl->status != ST_NONE1) break;   /* something to price after all */
                    if ((P->zeta < 0.0 && parm.obj_ll > -DBL_MAX && obj_track <= parm.obj_ll) ||
                        (P->zeta > 0.0 && parm.obj_ul < +DBL_MAX && obj_track >= parm.obj_ul)) {
                        rc = store_sol(P, GLP_INFEAS, GLP_FEAS, 0, it_cnt, true);
                        return rc ? rc : (P->zeta < 0.0 ? GLP_EOBJLL : GLP_EOBJUL);
                    }
                    if (it_limit() || tm_limit()) return stop_on_limit(it_limit() ? GLP_EITLIM : GLP_ETMLIM);
                    return store_sol(P, GLP_FEAS, GLP_FEAS, 0, it_cnt, true);
                }
                if (bbar_st != 1 || cbar_st != 1) {
                    if (bbar_st != 1) bbar_st = 0;
                    if (cbar_st != 1) cbar_st = 0;
                    break;
                }
                {
                    int p_stat, d_stat;
                    if (phase == 1) {
                        set_bnds(0);
                        clear_ctrl();
                        eval_bbar();
                        p_stat = GLP_INFEAS; d_stat = GLP_NOFEAS;
                    } else
                        p_stat = d_stat = GLP_FEAS;
                    return store_sol(P, p_stat, d_stat, 0, it_cnt);
                }
            case ST_NONE2:
                if (bbar_st != 1 || cbar_st != 1 || !rigorous) {
                    if (bbar_st != 1) bbar_st = 0;
                    if (cbar_st != 1) cbar_st = 0;
                    rigorous = 1;
                    break;
                }
                if (phase == 1) return spx_fail(P, it_cnt);
                {
                    int hp;
                    CK(cudaMemcpy(&hp, P->head + c.p, sizeof(int), cudaMemcpyDeviceToHost));
                    return store_sol(P, GLP_NOFEAS, GLP_FEAS, hp + 1, it_cnt);
                }
            case ST_PIVSMALL:
                rigorous = 5;
                break;
            case ST_PIV12:
                if (binv_st != 1) binv_st = 0;
                rigorous = 5;
                break;
            default:
                glpb_set_error("dual: unexpected device status %d", c.status);
                return GLPB_ESTATE;
            }
        }
    }
};

/* ------------------------------------------------------------------ */
/* drivers                                                            */
/* ------------------------------------------------------------------ */

/* trivial_lp, lib/glpapi06.js:149-258 (host: there is no matrix to stream) */
static void trivial_lp(glpb_prob *P, const glpb_smcp &parm)
{
    const int m = P->m, n = P->n;
    P->valid = 0;
    P->pbs_stat = P->dbs_stat = GLP_FEAS;
    P->obj_val = P->c0;
    P->some = 0;
    for (int i = 0; i < m; i++) {
        P->h_stat[i] = GLP_BS; P->h_prim[i] = P->h_dual[i] = 0.0;
        int t = P->h_type[i];
        if ((t == GLP_LO || t == GLP_DB || t == GLP_FX) && P->h_lb[i] > +parm.tol_bnd) {
            P->pbs_stat = GLP_NOFEAS;
            if (P->some == 0 && parm.meth != GLP_PRIMAL) P->some = i + 1;
        }
        if ((t == GLP_UP || t == GLP_DB || t == GLP_FX) && P->h_ub[i] < -parm.tol_bnd) {
            P->pbs_stat = GLP_NOFEAS;
            if (P->some == 0 && parm.meth != GLP_PRIMAL) P->some = i + 1;
        }
    }
    double zeta = 1.0;
    for (int j = 0; j < n; j++) zeta = std::max(zeta, fabs(P->h_coef[j]));
    zeta = (P->dir == GLP_MIN ? +1.0 : -1.0) / zeta;
    for (int j = 0; j < n; j++) {
        int k = m + j, t = P->h_type[k];
        double c = P->h_coef[j];
        bool lo;
        if (t == GLP_FR) { P->h_stat[k] = GLP_NF; P->h_prim[k] = 0.0; }
        else if (t == GLP_FX) { P->h_stat[k] = GLP_NS; P->h_prim[k] = P->h_lb[k]; }
        else {
            if (t == GLP_LO) lo = true;
            else if (t == GLP_UP) lo = false;
            else if (zeta * c > 0.0) lo = true;
            else if (zeta * c < 0.0) lo = false;
            else lo = fabs(P->h_lb[k]) <= fabs(P->h_ub[k]);
            P->h_stat[k] = lo ? GLP_NL : GLP_NU;
            P->h_prim[k] = lo ? P->h_lb[k] : P->h_ub[k];
        }
        P->h_dual[k] = c;
        P->obj_val += c * P->h_prim[k];
        if ((t == GLP_FR || t == GLP_LO) && zeta * c < -parm.tol_dj) {
            P->dbs_stat = GLP_NOFEAS;
            if (P->some == 0 && parm.meth == GLP_PRIMAL) P->some = k + 1;
        }
        if ((t == GLP_FR || t == GLP_UP) && zeta * c > +parm.tol_dj) {
            P->dbs_stat = GLP_NOFEAS;
            if (P->some == 0 && parm.meth == GLP_PRIMAL) P->some = k + 1;
        }
    }
}

static int check_smcp(const glpb_smcp &p)
{
    /* lib/glpapi06.js:271-300 */
    if (p.msg_lev < 0 || p.msg_lev > 4) return GLPB_EINVAL;
    if (!(p.meth == GLP_PRIMAL || p.meth == GLP_DUALP || p.meth == GLP_DUAL)) return GLPB_EINVAL;
    if (!(p.pricing == GLP_PT_STD || p.pricing == GLP_PT_PSE)) return GLPB_EINVAL;
    if (!(p.r_test == GLP_RT_STD || p.r_test == GLP_RT_HAR)) return GLPB_EINVAL;
    if (!(0.0 < p.tol_bnd && p.tol_bnd < 1.0)) return GLPB_EINVAL;
    if (!(0.0 < p.tol_dj && p.tol_dj < 1.0)) return GLPB_EINVAL;
    if (!(0.0 < p.tol_piv && p.tol_piv < 1.0)) return GLPB_EINVAL;
    if (p.it_lim < 0 || p.tm_lim < 0 || p.out_frq < 1 || p.out_dly < 0) return GLPB_EINVAL;
    if (!(p.presolve == 0 || p.presolve == 1)) return GLPB_EINVAL;
    return 0;
}

extern "C" int glpb_simplex(glpb_prob *P, const glpb_smcp *parm_)
{
    if (!P) return GLPB_EINVAL;
    glpb_smcp parm;
    if (parm_) parm = *parm_; else glpb_init_smcp(&parm);
    int rc = check_smcp(parm);
    if (rc) { glpb_set_error("glp_simplex: invalid control parameter"); return rc; }
    if (parm.presolve) { glpb_set_error("glp_simplex: presolve stays in the host binding (SURVEY 8f)"); return GLPB_EINVAL; }
    CK(cudaSetDevice(P->device));
    const int m = P->m, n = P->n;
    P->pbs_stat = P->dbs_stat = GLP_UNDEF;
    P->obj_val = 0.0;
    P->some = 0;
    for (int k = 0; k < m + n; k++)
        if (P->h_type[k] == GLP_DB && P->h_lb[k] >= P->h_ub[k]) return GLP_EBOUND;
    if (P->nnz == 0) { trivial_lp(P, parm); return 0; }
    /* solve_lp, lib/glpapi06.js:3-39.  glp_set_*_stat clears `valid` on any basic <-> non-basic
       flip, also when a sequence of them (branch-and-bound: restore the root statuses, then
       apply the node's) ends in the very set of basic variables the device inverse was left
       with; that inverse is then still the inverse of this basis */
    if (!P->valid && P->t_ok) {
        int nbs = 0;
        for (int k = 0; k < m + n; k++) nbs += (P->h_stat[k] == GLP_BS);
        bool same = (nbs == m);
        for (int i = 0; i < m && same; i++) same = (P->h_stat[P->h_head[i] - 1] == GLP_BS);
        if (same) P->valid = 1;
    }
    if (!P->valid) {
        rc = glpb_factorize(P);
        if (rc != 0) return rc;
    }
    /* device time of the whole solve: CUDA events on the solve stream */
    cudaEvent_t ev0, ev1;
    CK(cudaEventCreate(&ev0));
    CK(cudaEventCreate(&ev1));
    CK(cudaEventRecord(ev0, P->stream));
    int ret;
    if (parm.meth == GLP_PRIMAL) { Primal s(P, parm); ret = s.run(); }
    else {
        { Dual s(P, parm); ret = s.run(); }
        if (parm.meth == GLP_DUALP && ret == GLP_EFAIL && P->valid) { Primal s(P, parm); ret = s.run(); }
    }
    float ms = 0.f;
    cudaEventRecord(ev1, P->stream);
    cudaEventSynchronize(ev1);
    cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    P->last_solve_us = ms * 1000.0;
    return ret;
}

/* Pivot log for parity tests: the (q, p) pair of every iteration whose number
   (the handle's it_cnt before the iteration) is below `cap`; q = 1..n position
   in the non-basic list, p = 1..m position in the basis header, p = -1 for a
   bound flip (the reference's csa.q / csa.p, lib/glpspx01.js:30-32). */
extern "C" int glpb_set_pivot_log(glpb_prob *P, int cap)
{
    if (!P || cap < 0) return GLPB_EINVAL;
    CK(cudaSetDevice(P->device));
    CK(cudaStreamSynchronize(P->stream));
    if (P->piv_log) { cudaFree(P->piv_log); P->piv_log = nullptr; }
    P->piv_cap = cap;
    if (cap > 0) {
        CK(cudaMalloc((void **)&P->piv_log, (size_t)2 * cap * sizeof(int)));
        CK(cudaMemset(P->piv_log, 0xFF, (size_t)2 * cap * sizeof(int)));
    }
    int *ptr = P->piv_log;
    int capv[2] = {cap, 0};
    CK(cudaMemcpy(&P->ctrl->piv_log, &ptr, sizeof ptr, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(&P->ctrl->piv_cap, capv, sizeof capv, cudaMemcpyHostToDevice));
    return 0;
}

/* Live device vectors for parity tests, CSA layout of the reference (lib/glpspx01.js:13-38):
   "gamma" (primal: [n] by non-basic position, dual: [m] by basic position), "cbar" [n], "bbar" [m],
   "head" [m+n] as 1-based variable numbers, "stat" [n].  out receives `count` doubles. */
extern "C" int glpb_debug_get(glpb_prob *P, const char *name, double *out, int count)
{
    if (!P || !name || !out || count < 0) return GLPB_EINVAL;
    CK(cudaSetDevice(P->device));
    CK(cudaStreamSynchronize(P->stream));
    const int m = P->m, n = P->n;
    std::string s(name);
    if (s == "gamma" || s == "cbar" || s == "bbar") {
        const double *src = (s == "gamma") ? P->gamma : (s == "cbar" ? P->cbar : P->bbar);
        const int len = (s == "gamma") ? std::max(m, n) : (s == "cbar" ? n : m);
        if (count > len) return GLPB_EINVAL;
        CK(cudaMemcpy(out, src, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost));
        return 0;
    }
    if (s == "head") {
        if (count > m + n) return GLPB_EINVAL;
        std::vector<int> h(count);
        CK(cudaMemcpy(h.data(), P->head, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost));
        for (int i = 0; i < count; i++) out[i] = h[i] + 1;
        return 0;
    }
    if (s == "stat") {
        if (count > n) return GLPB_EINVAL;
        std::vector<signed char> h(count);
        CK(cudaMemcpy(h.data(), P->stat, (size_t)count, cudaMemcpyDeviceToHost));
        for (int i = 0; i < count; i++) out[i] = h[i];
        return 0;
    }
    return GLPB_EINVAL;
}

extern "C" int glpb_get_pivot_log(glpb_prob *P, int *qp, int cap, int *count)
{
    if (!P || !qp || !count) return GLPB_EINVAL;
    CK(cudaSetDevice(P->device));
    CK(cudaStreamSynchronize(P->stream));
    int n = std::min(std::min(cap, P->piv_cap), P->it_cnt);
    if (n < 0) n = 0;
    if (n > 0) CK(cudaMemcpy(qp, P->piv_log, (size_t)2 * n * sizeof(int), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) {
        qp[2 * i] += 1;
        qp[2 * i + 1] = (qp[2 * i + 1] == P_FLIP) ? -1 : qp[2 * i + 1] + 1;
    }
    *count = n;
    return 0;
}

extern "C" int glpb_get_status(glpb_prob *P)
{
    if (!P) return GLPB_EINVAL;
    int status = P->pbs_stat;
    if (status == GLP_FEAS) {
        if (P->dbs_stat == GLP_FEAS) status = GLP_OPT;
        else if (P->dbs_stat == GLP_NOFEAS) status = GLP_UNBND;
    }
    return status;
}

extern "C" int glpb_get_solution(glpb_prob *P, int *stat, double *prim, double *dual, int *head,
                                 int *pbs_stat, int *dbs_stat, double *obj_val, int *it_cnt, int *some)
{
    if (!P) return GLPB_EINVAL;
    const int m = P->m, n = P->n;
    if (stat) memcpy(stat, P->h_stat.data(), (m + n) * sizeof(int));
    if (prim) memcpy(prim, P->h_prim.data(), (m + n) * sizeof(double));
    if (dual) memcpy(dual, P->h_dual.data(), (m + n) * sizeof(double));
    if (head) memcpy(head, P->h_head.data(), m * sizeof(int));
    if (pbs_stat) *pbs_stat = P->pbs_stat;
    if (dbs_stat) *dbs_stat = P->dbs_stat;
    if (obj_val) *obj_val = P->obj_val;
    if (it_cnt) *it_cnt = P->it_cnt;
    if (some) *some = P->some;
    return 0;
}

extern "C" int glpb_get_counters(glpb_prob *P, long *out, int count)
{
    if (!P || !out) return GLPB_EINVAL;
    long v[9] = {P->n_iter, P->n_refac, P->n_launch, P->n_sync, P->n_update,
                 (long)(P->h_ctrl ? P->h_ctrl->k : 0), (long)P->last_solve_us, P->n_graph, P->n_tie};
    for (int i = 0; i < count && i < 9; i++) out[i] = v[i];
    return 0;
}

/* per-kernel device time measured with CUDA events on the solve stream;
   the report is text: one line "name count total_ms algorithmic_bytes" */
extern "C" int glpb_set_profile(glpb_prob *P, int on)
{
    if (!P) return GLPB_EINVAL;
    prof_collect(P);
    P->prof = on;
    if (on) {
        P->prof_acc.clear();
        P->n_eng_prof_iter[0] = P->n_eng_prof_iter[1] = 0;
        cudaMemsetAsync(P->eng_cyc, 0, 32 * sizeof(long long), P->stream);
        cudaMemsetAsync(P->eng_bytes, 0, 24 * sizeof(double), P->stream);
    }
    return 0;
}

extern "C" const char *glpb_profile_report(glpb_prob *P)
{
    if (!P) return "";
    prof_collect(P);
    P->prof_text.clear();
    char line[256];
    for (auto &kv : P->prof_acc) {
        snprintf(line, sizeof line, "%s %ld %.6f %.0f\n", kv.first.c_str(), kv.second.count, kv.second.ms, kv.second.bytes);
        P->prof_text += line;
    }
    /* phases of the persistent engine: SM cycles of CTA 0 between barriers,
       scaled to the CUDA-event time of the engine launches; count = iterations */
    long long cyc[32];
    double byt[24];
    if (cudaMemcpy(cyc, P->eng_cyc, sizeof cyc, cudaMemcpyDeviceToHost) == cudaSuccess &&
        cudaMemcpy(byt, P->eng_bytes, sizeof byt, cudaMemcpyDeviceToHost) == cudaSuccess) {
        static const char *pn[24] = {"P0_chuzc_first", "PA_tcol_head", "PB5_barrier", "PB4_ratio_local", "PR2_ratio2",
                                     "PC_rho_gemvT", "PD_unused", "PE_trow_svec", "PF_update_T_chuzc", "PB1_tail", "PB2_allreduce", "PB3_btran_head",
                                     "D0_chuzr_first", "D1_rho", "D2_trow", "DR1_ratio1_gamma_rhs", "DR2_ratio2",
                                     "DX_ratio_local_gamma_rhs", "D3_gemvN_tcol_head", "D4_tcol_tail_utail",
                                     "D5_update_allreduce_header", "D9_flush", "D1a_header_handoff", "D5a_update_own_work"};
        for (int half = 0; half < 2; half++) {
            const char *kn = half ? "k_engine_dual" : "k_engine_primal";
            auto itp = P->prof_acc.find(kn);
            if (itp == P->prof_acc.end()) continue;
            long long tot = 0;
            for (int i = 0; i < 12; i++) tot += cyc[half * 12 + i];
            if (tot <= 0) continue;
            for (int i = 0; i < 12; i++) {
                if (cyc[half * 12 + i] == 0) continue;
                snprintf(line, sizeof line, "eng_%s %ld %.6f %.0f\n", pn[half * 12 + i], P->n_eng_prof_iter[half],
                         itp->second.ms * (double)cyc[half * 12 + i] / (double)tot, byt[half * 12 + i]);
                P->prof_text += line;
            }
        }
        /* phases inside k_refactor (CTA 0's view) */
        auto itr = P->prof_acc.find("k_refactor");
        if (itr != P->prof_acc.end()) {
            static const char *rn[8] = {"build", "panel_pivots", "panel_io", "barrier_wait", "swap_update", "finish", "pivot_search", "pivot_bcast"};
            long long tot = 0;
            for (int i = 0; i < 8; i++) tot += cyc[24 + i];
            for (int i = 0; i < 8 && tot > 0; i++) {
                snprintf(line, sizeof line, "ref_%s %ld %.6f 0\n", rn[i], itr->second.count,
                         itr->second.ms * (double)cyc[24 + i] / (double)tot);
                P->prof_text += line;
            }
        }
    }
    return P->prof_text.c_str();
}

/* bfd_ftran / bfd_btran on the current factorisation (scaled space) */
static int solve_host_vec(glpb_prob *P, double *x, bool tr)
{
    if (!P || !x) return GLPB_EINVAL;
    if (!P->valid) return GLPB_ESTATE;
    CK(cudaSetDevice(P->device));
    const int m = P->m;
    Dev D(P);
    int rc = sync_ctrl(P);
    if (rc) return rc;
    D.k = P->h_ctrl->k;
    if ((rc = h2d(P, P->w1, x, m))) return rc;
    LAUNCH(P, k_clear_ctrl, 1, 1, 0, P->ctrl, 0);
    if (tr) dev_btran(D, P->w1, P->w2); else dev_ftran(D, P->w1, P->w2);
    CK(cudaMemcpyAsync(x, P->w2, m * sizeof(double), cudaMemcpyDeviceToHost, P->stream));
    CK(cudaStreamSynchronize(P->stream));
    return 0;
}

extern "C" int glpb_ftran(glpb_prob *P, double *x) { return solve_host_vec(P, x, false); }
extern "C" int glpb_btran(glpb_prob *P, double *x) { return solve_host_vec(P, x, true); }
