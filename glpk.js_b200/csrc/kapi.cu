/* kapi.cu -- kernel-level entry points of the C ABI.
 *
 * Each function takes the reference's own CSA arrays (1-based, slot 0
 * unused; lib/glpspx01.js:5-40, lib/glpspx02.js:5-43), stages them on the
 * device, runs exactly the kernel the solver uses and returns the selected
 * index in the reference's convention.  This is what the parity tests call to
 * show bit-exact entering/leaving indices on the reference's basis and
 * weights.
 */
#include "kernels.cuh"
#include <cstring>
#include <vector>

namespace {

struct Tmp { /* scoped device buffers on the default stream */
    std::vector<void *> ptrs;
    bool ok = true;
    ~Tmp() { for (void *p : ptrs) cudaFree(p); }
    template <class T> T *up(const T *host, size_t count)
    {
        T *d = nullptr;
        if (cudaMalloc((void **)&d, (count ? count : 1) * sizeof(T)) != cudaSuccess) { ok = false; return nullptr; }
        ptrs.push_back(d);
        if (host && count)
            if (cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) ok = false;
        return d;
    }
    template <class T> T *zero(size_t count)
    {
        T *d = up<T>(nullptr, count);
        if (d) cudaMemset(d, 0, (count ? count : 1) * sizeof(T));
        return d;
    }
};

int need_device()
{
    if (glpb_device_count() < 1) { glpb_set_error("no CUDA device; there is no CPU fallback"); return GLPB_ENODEV; }
    return 0;
}

int finish(Tmp &t, Ctrl *dctrl, Ctrl *out)
{
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(out, dctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess || !t.ok) { glpb_set_error("kernel entry point: %s", cudaGetErrorString(e)); return GLPB_ENODEV; }
    return 0;
}

int grid1(int len) { int g = (len + 255) / 256; return g < 1 ? 1 : (g > 1184 ? 1184 : g); }

std::vector<int> shift(const int *a, size_t count)
{
    std::vector<int> v(count);
    for (size_t i = 0; i < count; i++) v[i] = a[i + 1] - 1;
    return v;
}

} /* namespace */

extern "C" int glpb_k_chuzc_primal(int n, const signed char *stat, const double *cbar,
                                   const double *gamma, double tol_dj, int *q)
{
    int rc = need_device();
    if (rc) return rc;
    if (n < 1 || !stat || !cbar || !gamma || !q) return GLPB_EINVAL;
    Tmp t;
    signed char *d_stat = t.up(stat + 1, n);
    double *d_cbar = t.up(cbar + 1, n), *d_gamma = t.up(gamma + 1, n);
    Ctrl *ctrl = t.zero<Ctrl>(1);
    Key *scr = t.zero<Key>(4096);
    if (!t.ok) return GLPB_ENOMEM;
    k_chuzc_primal<<<grid1(n), 256>>>(ctrl, n, d_stat, d_cbar, d_gamma, tol_dj, 0, scr);
    Ctrl h;
    if ((rc = finish(t, ctrl, &h))) return rc;
    *q = (h.q >= 0) ? h.q + 1 : 0;
    return 0;
}

extern "C" int glpb_k_chuzr_dual(int m, int n, const signed char *type, const double *lb,
                                 const double *ub, const int *head, const double *bbar,
                                 const double *gamma, double tol_bnd, int *p, double *delta)
{
    int rc = need_device();
    if (rc) return rc;
    if (m < 1 || n < 1 || !type || !lb || !ub || !head || !bbar || !gamma || !p || !delta) return GLPB_EINVAL;
    Tmp t;
    std::vector<int> head0 = shift(head, m + n);
    signed char *d_type = t.up(type + 1, m + n);
    double *d_lb = t.up(lb + 1, m + n), *d_ub = t.up(ub + 1, m + n);
    int *d_head = t.up(head0.data(), m + n);
    double *d_bbar = t.up(bbar + 1, m), *d_gamma = t.up(gamma + 1, m);
    Ctrl *ctrl = t.zero<Ctrl>(1);
    Key *scr = t.zero<Key>(4096);
    if (!t.ok) return GLPB_ENOMEM;
    k_chuzr_dual<<<grid1(m), 256>>>(ctrl, m, d_type, d_lb, d_ub, d_head, d_bbar, d_gamma, tol_bnd, 0, scr);
    Ctrl h;
    if ((rc = finish(t, ctrl, &h))) return rc;
    *p = (h.p >= 0) ? h.p + 1 : 0;
    *delta = h.delta;
    return 0;
}

extern "C" int glpb_k_ratio_primal(int m, int n, const signed char *type, const double *lb,
                                   const double *ub, const double *coef, const int *head, int phase,
                                   const double *bbar, double cbar_q, int q, const int *tcol_ind,
                                   const double *tcol_vec, int tcol_num, double rtol, int *p,
                                   int *p_stat, double *teta)
{
    int rc = need_device();
    if (rc) return rc;
    if (m < 1 || n < 1 || q < 1 || q > n || tcol_num < 0 || tcol_num > m || !p || !p_stat || !teta) return GLPB_EINVAL;
    Tmp t;
    std::vector<int> head0 = shift(head, m + n), ind0 = shift(tcol_ind, tcol_num);
    signed char *d_type = t.up(type + 1, m + n);
    double *d_lb = t.up(lb + 1, m + n), *d_ub = t.up(ub + 1, m + n), *d_coef = t.up(coef + 1, m + n);
    int *d_head = t.up(head0.data(), m + n), *d_ind = t.up(ind0.data(), tcol_num);
    double *d_bbar = t.up(bbar + 1, m), *d_tcol = t.up(tcol_vec + 1, m);
    Ctrl hc;
    memset(&hc, 0, sizeof hc);
    hc.q = q - 1; hc.phase = phase; hc.d1 = cbar_q; hc.rigorous = 1;
    Ctrl *ctrl = t.up(&hc, 1);
    Key *scr = t.zero<Key>(4096);
    if (!t.ok) return GLPB_ENOMEM;
    for (int pass = 1; pass <= 2; pass++)
        k_ratio_primal<<<grid1(tcol_num), 256>>>(ctrl, pass, m, d_type, d_lb, d_ub, d_coef, d_head, d_bbar,
                                                 d_tcol, d_ind, tcol_num, rtol, scr);
    Ctrl h;
    if ((rc = finish(t, ctrl, &h))) return rc;
    *p = (h.p >= 0) ? h.p + 1 : (h.p == P_FLIP ? -1 : 0);
    *p_stat = h.p_stat;
    *teta = h.teta;
    return 0;
}

extern "C" int glpb_k_ratio_dual(int n, const signed char *stat, const double *cbar, double delta,
                                 const int *trow_ind, const double *trow_vec, int trow_num,
                                 double rtol, int *q, double *new_dq)
{
    int rc = need_device();
    if (rc) return rc;
    if (n < 1 || trow_num < 0 || trow_num > n || !q || !new_dq) return GLPB_EINVAL;
    Tmp t;
    std::vector<int> ind0 = shift(trow_ind, trow_num);
    signed char *d_stat = t.up(stat + 1, n);
    double *d_cbar = t.up(cbar + 1, n), *d_trow = t.up(trow_vec + 1, n);
    int *d_ind = t.up(ind0.data(), trow_num);
    Ctrl hc;
    memset(&hc, 0, sizeof hc);
    hc.delta = delta; hc.rigorous = 1;
    Ctrl *ctrl = t.up(&hc, 1);
    Key *scr = t.zero<Key>(4096);
    if (!t.ok) return GLPB_ENOMEM;
    for (int pass = 1; pass <= 2; pass++)
        k_ratio_dual<<<grid1(trow_num), 256>>>(ctrl, pass, d_stat, d_cbar, d_trow, d_ind, trow_num, rtol, scr);
    Ctrl h;
    if ((rc = finish(t, ctrl, &h))) return rc;
    *q = (h.q >= 0) ? h.q + 1 : 0;
    *new_dq = h.new_dq;
    return 0;
}

/* sort_tcol / sort_trow: the significant entries of vec[1..n] (|v| >= eps) in the order the reference's
   swap loop leaves them (lib/glpspx01.js:795-804, lib/glpspx02.js:780-789); list[1..*num], 1-based */
extern "C" int glpb_k_sort_list(int n, const double *vec, double eps, int *list, int *num)
{
    int rc = need_device();
    if (rc) return rc;
    if (n < 1 || !vec || !list || !num) return GLPB_EINVAL;
    Tmp t;
    double *d_vec = t.up(vec + 1, n);
    int *d_list = t.zero<int>(n), *d_tmp = t.zero<int>(2 * (size_t)n + 2);
    Ctrl hc;
    memset(&hc, 0, sizeof hc);
    hc.eps = eps;
    Ctrl *ctrl = t.up(&hc, 1);
    if (!t.ok) return GLPB_ENOMEM;
    k_sort_list<<<1, 1024>>>(ctrl, d_vec, n, d_list, d_tmp);
    Ctrl h;
    if ((rc = finish(t, ctrl, &h))) return rc;
    *num = h.list_num;
    std::vector<int> out(n);
    if (cudaMemcpy(out.data(), d_list, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return GLPB_ENODEV;
    for (int i = 0; i < h.list_num; i++) list[i + 1] = out[i] + 1;
    return 0;
}

extern "C" int glpb_k_trow(int m, int n, const int *A_ptr, const int *A_ind, const double *A_val,
                           const int *head, const signed char *stat, const double *rho,
                           double *trow_vec)
{
    int rc = need_device();
    if (rc) return rc;
    if (m < 1 || n < 1 || !A_ptr || !head || !stat || !rho || !trow_vec) return GLPB_EINVAL;
    Tmp t;
    const int nnz = A_ptr[n + 1] - 1;
    std::vector<int> ptr0 = shift(A_ptr, n + 1), ind0 = shift(A_ind, nnz), head0 = shift(head, m + n);
    int *d_ptr = t.up(ptr0.data(), n + 1), *d_ind = t.up(ind0.data(), nnz), *d_head = t.up(head0.data(), m + n);
    double *d_val = t.up(A_val + 1, nnz), *d_rho = t.up(rho + 1, m);
    signed char *d_stat = t.up(stat + 1, n);
    double *d_trow = t.zero<double>(n), *d_s = t.zero<double>(n);
    Ctrl *ctrl = t.zero<Ctrl>(1);
    if (!t.ok) return GLPB_ENOMEM;
    const double avg = (double)nnz / n;
    if (avg >= 48.0) k_trow<32><<<(int)(((long)n * 32 + 255) / 256), 256>>>(ctrl, m, n, d_ptr, d_ind, d_val, d_head, d_stat, d_rho, nullptr, d_trow, d_s, 0);
    else if (avg >= 12.0) k_trow<8><<<(int)(((long)n * 8 + 255) / 256), 256>>>(ctrl, m, n, d_ptr, d_ind, d_val, d_head, d_stat, d_rho, nullptr, d_trow, d_s, 0);
    else k_trow<4><<<(int)(((long)n * 4 + 255) / 256), 256>>>(ctrl, m, n, d_ptr, d_ind, d_val, d_head, d_stat, d_rho, nullptr, d_trow, d_s, 0);
    Ctrl h;
    if ((rc = finish(t, ctrl, &h))) return rc;
    if (cudaMemcpy(trow_vec + 1, d_trow, n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return GLPB_ENODEV;
    return 0;
}

/* Stand-alone timing of one streaming kernel over resident synthetic data.
 * name: "chuzc_primal" (17 B/col), "chuzr_dual" (37 B/row),
 *       "update_rank1" (16 k^2 B; m is taken as k), "trow" (12 B/nnz, 16 nnz/col).
 * Buffers are made larger than L2 where the kernel's working set allows it by
 * rotating over several independent copies. */
extern "C" int glpb_bench_kernel(const char *name, int m, int n, int reps, double *usec, double *bytes)
{
    int rc = need_device();
    if (rc) return rc;
    if (!name || reps < 1 || !usec || !bytes) return GLPB_EINVAL;
    Tmp t;
    Ctrl *ctrl = t.zero<Ctrl>(1);
    Key *scr = t.zero<Key>(4096);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0.f;
    std::string nm(name);
    if (nm == "chuzc_primal") {
        const int copies = 8;
        std::vector<signed char> st(n);
        std::vector<double> cb(n), ga(n);
        for (int j = 0; j < n; j++) { st[j] = (signed char)(2 + (j % 3 == 0)); cb[j] = ((j * 2654435761u) % 2001) / 1000.0 - 1.0; ga[j] = 1.0 + (j % 7); }
        std::vector<signed char *> ds(copies); std::vector<double *> dc(copies), dg(copies);
        for (int c = 0; c < copies; c++) { ds[c] = t.up(st.data(), n); dc[c] = t.up(cb.data(), n); dg[c] = t.up(ga.data(), n); }
        if (!t.ok) return GLPB_ENOMEM;
        for (int r = 0; r < 3; r++) k_chuzc_primal<<<grid1(n), 256>>>(ctrl, n, ds[0], dc[0], dg[0], 1e-7, 0, scr);
        cudaEventRecord(e0);
        for (int r = 0; r < reps; r++) { int c = r % copies; k_chuzc_primal<<<grid1(n), 256>>>(ctrl, n, ds[c], dc[c], dg[c], 1e-7, 0, scr); }
        cudaEventRecord(e1);
        *bytes = 17.0 * n;
    } else if (nm == "update_rank1") {
        const int k = m, ld = (k + 7) & ~7;
        double *T = t.zero<double>((size_t)ld * ld);
        std::vector<double> v(ld, 0.5);
        std::vector<int> id(ld);
        for (int i = 0; i < ld; i++) id[i] = i;
        double *tc = t.up(v.data(), ld), *rh = t.up(v.data(), ld);
        int *sp = t.up(id.data(), ld), *sr = t.up(id.data(), ld);
        Ctrl hc; memset(&hc, 0, sizeof hc); hc.k = k; hc.p = 0; hc.q = 0;
        cudaMemcpy(ctrl, &hc, sizeof hc, cudaMemcpyHostToDevice);
        if (!t.ok) return GLPB_ENOMEM;
        dim3 grid((k + UPD_TB - 1) / UPD_TB, (k + UPD_TC - 1) / UPD_TC);
        for (int r = 0; r < 3; r++) k_update_rank1<<<grid, UPD_TB>>>(ctrl, T, ld, tc, rh, sp, sr);
        cudaEventRecord(e0);
        for (int r = 0; r < reps; r++) k_update_rank1<<<grid, UPD_TB>>>(ctrl, T, ld, tc, rh, sp, sr);
        cudaEventRecord(e1);
        *bytes = 16.0 * k * (double)k;
    } else if (nm == "chuzr_dual" || nm == "chuzr_dual_seq") {
        const bool seq = (nm == "chuzr_dual_seq");     /* header in ascending order, as glp_factorize leaves it */
        /* 37 B per row: head, then type/lb/ub gathered through it, bbar, gamma; m rows,
           every basic variable a structural one of an n = m column problem */
        const int copies = 4;
        std::vector<signed char> ty(2 * (size_t)m);
        std::vector<double> lo(2 * (size_t)m), hi(2 * (size_t)m), bb(m), ga(m);
        std::vector<int> hd(m);
        for (int i = 0; i < m; i++) {
            hd[i] = seq ? m + i : m + (int)(((unsigned)i * 2654435761u) % (unsigned)m);      /* scattered gather */
            bb[i] = ((i * 40503u) % 2001) / 1000.0 - 1.0; ga[i] = 1.0 + (i % 5);
        }
        for (size_t k = 0; k < 2 * (size_t)m; k++) { ty[k] = (signed char)(k % 3 == 0 ? GLP_DB : GLP_LO); lo[k] = -0.5; hi[k] = 0.5; }
        std::vector<signed char *> dt(copies); std::vector<double *> dl(copies), du(copies), db(copies), dg(copies);
        std::vector<int *> dh(copies);
        for (int c = 0; c < copies; c++) {
            dt[c] = t.up(ty.data(), ty.size()); dl[c] = t.up(lo.data(), lo.size()); du[c] = t.up(hi.data(), hi.size());
            db[c] = t.up(bb.data(), m); dg[c] = t.up(ga.data(), m); dh[c] = t.up(hd.data(), m);
        }
        if (!t.ok) return GLPB_ENOMEM;
        const int g = m <= 65536 ? 1 : grid1(m), th = m <= 65536 ? 1024 : 256;
        for (int r = 0; r < 3; r++) k_chuzr_dual<<<g, th>>>(ctrl, m, dt[0], dl[0], du[0], dh[0], db[0], dg[0], 1e-7, 0, scr);
        cudaEventRecord(e0);
        for (int r = 0; r < reps; r++) { int c = r % copies; k_chuzr_dual<<<g, th>>>(ctrl, m, dt[c], dl[c], du[c], dh[c], db[c], dg[c], 1e-7, 0, scr); }
        cudaEventRecord(e1);
        *bytes = 37.0 * m;
    } else if (nm == "trow") {
        /* pivot-row SpMV over n non-basic structural columns of 16 non-zeros each
           (the C3 shape), rows scattered; 12 B per non-zero + 13 n + 8 m */
        const int per = 16, copies = 2;
        const size_t nnz = (size_t)n * per;
        if (m < per || nnz > 0x7fffffffu) { cudaEventDestroy(e0); cudaEventDestroy(e1); return GLPB_EINVAL; }
        std::vector<int> ptr(n + 1), ind(nnz), hd((size_t)m + n);
        std::vector<double> val(nnz), rho(m);
        std::vector<signed char> st(n, (signed char)GLP_NL);
        for (int j = 0; j <= n; j++) ptr[j] = j * per;
        for (size_t e = 0; e < nnz; e++) { ind[e] = (int)((e * 2654435761ull + (e >> 4)) % (unsigned)m); val[e] = 0.1 + (e % 9) * 0.1; }
        for (int i = 0; i < m; i++) { hd[i] = i; rho[i] = (i % 5 == 0) ? 0.0 : 1.0 / (1 + i % 7); }
        for (int j = 0; j < n; j++) hd[(size_t)m + j] = m + j;
        std::vector<int *> dp(copies), di(copies); std::vector<double *> dv(copies);
        for (int c = 0; c < copies; c++) { dp[c] = t.up(ptr.data(), ptr.size()); di[c] = t.up(ind.data(), nnz); dv[c] = t.up(val.data(), nnz); }
        int *dh = t.up(hd.data(), hd.size());
        signed char *ds = t.up(st.data(), n);
        double *dr = t.up(rho.data(), m), *dtr = t.zero<double>(n), *dsv = t.zero<double>(n);
        if (!t.ok) return GLPB_ENOMEM;
        /* rho staged in shared memory when it fits (the C3 row count does): one CTA of 1024 threads per SM,
           walking the columns with a grid stride, 4 columns per group and step */
        const bool stage = (size_t)m * 8 <= 200 * 1024;
        int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
        const int th = stage ? 1024 : 256;
        const int g = stage ? sms : (int)std::min<long>((((long)n * 8 / 4) + 255) / 256, 148L * 64);
        const size_t sh = stage ? (size_t)m * 8 : 0;
        if (stage) cudaFuncSetAttribute(k_trow<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
        const int wm = 1 | (stage ? 2 : 0);
        for (int r = 0; r < 3; r++) k_trow<8><<<g, th, sh>>>(ctrl, m, n, dp[0], di[0], dv[0], dh, ds, dr, nullptr, dtr, dsv, wm);
        cudaEventRecord(e0);
        for (int r = 0; r < reps; r++) { int c = r % copies; k_trow<8><<<g, th, sh>>>(ctrl, m, n, dp[c], di[c], dv[c], dh, ds, dr, nullptr, dtr, dsv, wm); }
        cudaEventRecord(e1);
        *bytes = 12.0 * (double)nnz + 13.0 * n + 8.0 * m;
    } else {
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return GLPB_EINVAL;
    }
    cudaError_t e = cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e != cudaSuccess) { glpb_set_error("bench: %s", cudaGetErrorString(e)); return GLPB_ENODEV; }
    *usec = 1000.0 * ms / reps;
    return 0;
}
