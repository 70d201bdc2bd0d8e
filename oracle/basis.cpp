/* basis.cpp -- CPU ORACLE (test infrastructure only; see glpo.h).
 *
 * Restatement of the reference's basis factorisation stack:
 *   sparse LU with Markowitz/threshold pivoting   lib/glpluf.js
 *   Forrest-Tomlin update  B = F*H*V              lib/glpfhv.js
 *   driver / error mapping                        lib/glpbfd.js
 * Only the FT variant (GLP_BF_FT, the default) is restated.
 */
#include "glpo.h"
#include <cassert>
#include <cmath>
#include <cstring>

namespace glpo {

/* ------------------------------------------------------------------ */
/* Sparse Vector Area helpers                                         */
/* ------------------------------------------------------------------ */

/* lib/glpluf.js:40-94 luf_defrag_sva: squeeze rows/cols of V to the left in
   the order of the addressing list so that all free space is contiguous */
static void defrag_sva(LUF &L)
{
    const int n = L.n;
    int pos = 1, k;
    /* leading nodes that already sit where they should are only trimmed */
    for (k = L.sv_head; k != 0; k = L.sv_next[k]) {
        if (k <= n) {
            if (L.vr_ptr[k] != pos) break;
            L.vr_cap[k] = L.vr_len[k];
            pos += L.vr_cap[k];
        } else {
            int j = k - n;
            if (L.vc_ptr[j] != pos) break;
            L.vc_cap[j] = L.vc_len[j];
            pos += L.vc_cap[j];
        }
    }
    /* the rest are moved */
    for (; k != 0; k = L.sv_next[k]) {
        if (k <= n) {
            int len = L.vr_len[k];
            memmove(&L.sv_ind[pos], &L.sv_ind[L.vr_ptr[k]], len * sizeof(int));
            memmove(&L.sv_val[pos], &L.sv_val[L.vr_ptr[k]], len * sizeof(double));
            L.vr_ptr[k] = pos;
            L.vr_cap[k] = len;
            pos += len;
        } else {
            int j = k - n, len = L.vc_len[j];
            memmove(&L.sv_ind[pos], &L.sv_ind[L.vc_ptr[j]], len * sizeof(int));
            memmove(&L.sv_val[pos], &L.sv_val[L.vc_ptr[j]], len * sizeof(double));
            L.vc_ptr[j] = pos;
            L.vc_cap[j] = len;
            pos += len;
        }
    }
    L.sv_beg = pos;
}

/* move node k of the addressing list to its tail; the room it leaves is
   donated to its predecessor (lib/glpluf.js:131-153, 192-214) */
static void relink_to_tail(LUF &L, int k, int cur)
{
    const int n = L.n;
    if (L.sv_prev[k] == 0)
        L.sv_head = L.sv_next[k];
    else {
        int kk = L.sv_prev[k];
        if (kk <= n) L.vr_cap[kk] += cur; else L.vc_cap[kk - n] += cur;
        L.sv_next[L.sv_prev[k]] = L.sv_next[k];
    }
    if (L.sv_next[k] == 0)
        L.sv_tail = L.sv_prev[k];
    else
        L.sv_prev[L.sv_next[k]] = L.sv_prev[k];
    L.sv_prev[k] = L.sv_tail;
    L.sv_next[k] = 0;
    if (L.sv_prev[k] == 0)
        L.sv_head = k;
    else
        L.sv_next[L.sv_prev[k]] = k;
    L.sv_tail = k;
}

/* lib/glpluf.js:96-155 luf_enlarge_row; returns non-zero on SVA overflow */
static int enlarge_row(LUF &L, int i, int cap)
{
    assert(L.vr_cap[i] < cap);
    if (L.sv_end - L.sv_beg < cap) {
        defrag_sva(L);
        if (L.sv_end - L.sv_beg < cap) return 1;
    }
    int cur = L.vr_cap[i];
    memmove(&L.sv_ind[L.sv_beg], &L.sv_ind[L.vr_ptr[i]], L.vr_len[i] * sizeof(int));
    memmove(&L.sv_val[L.sv_beg], &L.sv_val[L.vr_ptr[i]], L.vr_len[i] * sizeof(double));
    L.vr_ptr[i] = L.sv_beg;
    L.vr_cap[i] = cap;
    L.sv_beg += cap;
    relink_to_tail(L, i, cur);
    return 0;
}

/* lib/glpluf.js:157-216 luf_enlarge_col */
static int enlarge_col(LUF &L, int j, int cap)
{
    assert(L.vc_cap[j] < cap);
    if (L.sv_end - L.sv_beg < cap) {
        defrag_sva(L);
        if (L.sv_end - L.sv_beg < cap) return 1;
    }
    int cur = L.vc_cap[j];
    memmove(&L.sv_ind[L.sv_beg], &L.sv_ind[L.vc_ptr[j]], L.vc_len[j] * sizeof(int));
    memmove(&L.sv_val[L.sv_beg], &L.sv_val[L.vc_ptr[j]], L.vc_len[j] * sizeof(double));
    L.vc_ptr[j] = L.sv_beg;
    L.vc_cap[j] = cap;
    L.sv_beg += cap;
    relink_to_tail(L, L.n + j, cur);
    return 0;
}

/* lib/glpluf.js:218-249 reallocate */
static void reallocate(LUF &L, int n)
{
    L.n = n;
    if ((int)L.fr_ptr.size() >= 1 + n) return;
    int cap = 1 + n + 100;
    L.fr_ptr.assign(cap, 0); L.fr_len.assign(cap, 0);
    L.fc_ptr.assign(cap, 0); L.fc_len.assign(cap, 0);
    L.vr_ptr.assign(cap, 0); L.vr_len.assign(cap, 0); L.vr_cap.assign(cap, 0);
    L.vr_piv.assign(cap, 0.0);
    L.vc_ptr.assign(cap, 0); L.vc_len.assign(cap, 0); L.vc_cap.assign(cap, 0);
    L.pp_row.assign(cap, 0); L.pp_col.assign(cap, 0);
    L.qq_row.assign(cap, 0); L.qq_col.assign(cap, 0);
    L.sv_prev.assign(2 * cap, 0); L.sv_next.assign(2 * cap, 0);
    L.vr_max.assign(cap, 0.0);
    L.rs_head.assign(cap, 0); L.rs_prev.assign(cap, 0); L.rs_next.assign(cap, 0);
    L.cs_head.assign(cap, 0); L.cs_prev.assign(cap, 0); L.cs_next.assign(cap, 0);
    L.flag.assign(cap, 0);
    L.work.assign(cap, 0.0);
}

/* active-set list primitives (rows keyed by vr_len, columns by vc_len) */
static inline void rs_remove(LUF &L, int i)
{
    if (L.rs_prev[i] == 0) L.rs_head[L.vr_len[i]] = L.rs_next[i];
    else L.rs_next[L.rs_prev[i]] = L.rs_next[i];
    if (L.rs_next[i] != 0) L.rs_prev[L.rs_next[i]] = L.rs_prev[i];
}
static inline void rs_insert(LUF &L, int i)
{
    int len = L.vr_len[i];
    L.rs_prev[i] = 0;
    L.rs_next[i] = L.rs_head[len];
    if (L.rs_next[i] != 0) L.rs_prev[L.rs_next[i]] = i;
    L.rs_head[len] = i;
}
static inline void cs_remove(LUF &L, int j)
{
    if (L.cs_prev[j] == 0) L.cs_head[L.vc_len[j]] = L.cs_next[j];
    else L.cs_next[L.cs_prev[j]] = L.cs_next[j];
    if (L.cs_next[j] != 0) L.cs_prev[L.cs_next[j]] = L.cs_prev[j];
}
static inline void cs_insert(LUF &L, int j)
{
    int len = L.vc_len[j];
    L.cs_prev[j] = 0;
    L.cs_next[j] = L.cs_head[len];
    if (L.cs_next[j] != 0) L.cs_prev[L.cs_next[j]] = j;
    L.cs_head[len] = j;
}

/* lib/glpluf.js:251-435 initialize: V := A (col-wise then row-wise), F := I,
   P = Q = I, active lists by count.  Returns 1 on SVA overflow. */
static int luf_initialize(LUF &L, col_fn col, void *info)
{
    const int n = L.n;
    int sv_beg = 1, sv_end = L.sv_size + 1;
    for (int j = 1; j <= n; j++) { L.fc_ptr[j] = sv_end; L.fc_len[j] = 0; }
    for (int i = 1; i <= n; i++) { L.vr_len[i] = L.vr_cap[i] = 0; L.flag[i] = 0; }
    int nnz = 0;
    double big = 0.0;
    for (int j = 1; j <= n; j++) {
        int *rn = L.pp_row.data();
        double *aj = L.work.data();
        int len = col(info, j, rn, aj);
        assert(0 <= len && len <= n);
        if (sv_end - sv_beg < len) return 1;
        L.vc_ptr[j] = sv_beg;
        L.vc_len[j] = L.vc_cap[j] = len;
        nnz += len;
        for (int t = 1; t <= len; t++) {
            int i = rn[t];
            double v = aj[t];
            assert(1 <= i && i <= n && !L.flag[i] && v != 0.0);
            L.sv_ind[sv_beg] = i;
            L.sv_val[sv_beg] = v;
            sv_beg++;
            if (v < 0.0) v = -v;
            if (big < v) big = v;
            L.flag[i] = 1;
            L.vr_cap[i]++;
        }
        for (int t = 1; t <= len; t++) L.flag[rn[t]] = 0;
    }
    for (int i = 1; i <= n; i++) {
        int len = L.vr_cap[i];
        if (sv_end - sv_beg < len) return 1;
        L.vr_ptr[i] = sv_beg;
        sv_beg += len;
    }
    for (int j = 1; j <= n; j++) {
        int jb = L.vc_ptr[j], je = jb + L.vc_len[j] - 1;
        for (int k = jb; k <= je; k++) {
            int i = L.sv_ind[k];
            int ip = L.vr_ptr[i] + L.vr_len[i];
            L.sv_ind[ip] = j;
            L.sv_val[ip] = L.sv_val[k];
            L.vr_len[i]++;
        }
    }
    for (int k = 1; k <= n; k++)
        L.pp_row[k] = L.pp_col[k] = L.qq_row[k] = L.qq_col[k] = k;
    L.sv_beg = sv_beg;
    L.sv_end = sv_end;
    /* physical order: columns n+1..n+n first, then rows 1..n */
    L.sv_head = n + 1;
    L.sv_tail = n;
    for (int i = 1; i <= n; i++) { L.sv_prev[i] = i - 1; L.sv_next[i] = i + 1; }
    L.sv_prev[1] = n + n;
    L.sv_next[n] = 0;
    for (int j = 1; j <= n; j++) { L.sv_prev[n + j] = n + j - 1; L.sv_next[n + j] = n + j + 1; }
    L.sv_prev[n + 1] = 0;
    L.sv_next[n + n] = 1;
    for (int k = 1; k <= n; k++) { L.flag[k] = 0; L.work[k] = 0.0; }
    L.nnz_a = nnz; L.nnz_f = 0; L.nnz_v = nnz;
    L.max_a = big; L.big_v = big;
    L.rank = -1;
    for (int i = 1; i <= n; i++) L.vr_max[i] = -1.0;
    for (int len = 0; len <= n; len++) L.rs_head[len] = 0;
    for (int i = 1; i <= n; i++) rs_insert(L, i);
    for (int len = 0; len <= n; len++) L.cs_head[len] = 0;
    for (int j = 1; j <= n; j++) cs_insert(L, j);
    return 0;
}

static inline double row_max(LUF &L, int i)
{
    double big = L.vr_max[i];
    if (big < 0.0) {
        int ib = L.vr_ptr[i], ie = ib + L.vr_len[i] - 1;
        for (int t = ib; t <= ie; t++) {
            double v = L.sv_val[t];
            if (v < 0.0) v = -v;
            if (big < v) big = v;
        }
        L.vr_max[i] = big;
    }
    return big;
}

/* lib/glpluf.js:437-635 find_pivot: Markowitz search with threshold piv_tol,
   singleton shortcuts, at most piv_lim candidates, Uwe Suhl's column
   exclusion.  Returns non-zero when the active submatrix is empty. */
static int find_pivot(LUF &L, int &p, int &q)
{
    const int n = L.n;
    p = q = 0;
    double best = DBL_MAX;
    int ncand = 0;
    int j = L.cs_head[1];
    if (j != 0) { p = L.sv_ind[L.vc_ptr[j]]; q = j; return 0; }
    int i = L.rs_head[1];
    if (i != 0) { p = i; q = L.sv_ind[L.vr_ptr[i]]; return 0; }
    for (int len = 2; len <= n; len++) {
        int next_j;
        for (j = L.cs_head[len]; j != 0; j = next_j) {
            int jb = L.vc_ptr[j], je = jb + L.vc_len[j] - 1;
            next_j = L.cs_next[j];
            int min_p = 0, min_q = 0, min_len = INT_MAX;
            for (int jp = jb; jp <= je; jp++) {
                i = L.sv_ind[jp];
                if (L.vr_len[i] >= min_len) continue;
                double big = row_max(L, i);
                int ip;
                for (ip = L.vr_ptr[i]; L.sv_ind[ip] != j; ip++) {}
                double temp = L.sv_val[ip];
                if (temp < 0.0) temp = -temp;
                if (temp < L.piv_tol * big) continue;
                min_p = i; min_q = j; min_len = L.vr_len[i];
                if (min_len <= len) { p = min_p; q = min_q; return 0; }
            }
            if (min_p != 0) {
                ncand++;
                double cost = (double)(min_len - 1) * (double)(len - 1);
                if (cost < best) { p = min_p; q = min_q; best = cost; }
                if (ncand == L.piv_lim) return 0;
            } else if (L.suhl) {
                /* exclude the column until it becomes a singleton */
                cs_remove(L, j);
                L.cs_prev[j] = L.cs_next[j] = j;
            }
        }
        for (i = L.rs_head[len]; i != 0; i = L.rs_next[i]) {
            int ib = L.vr_ptr[i], ie = ib + L.vr_len[i] - 1;
            double big = row_max(L, i);
            int min_p = 0, min_q = 0, min_len = INT_MAX;
            for (int ip = ib; ip <= ie; ip++) {
                j = L.sv_ind[ip];
                if (L.vc_len[j] >= min_len) continue;
                double temp = L.sv_val[ip];
                if (temp < 0.0) temp = -temp;
                if (temp < L.piv_tol * big) continue;
                min_p = i; min_q = j; min_len = L.vc_len[j];
                if (min_len <= len) { p = min_p; q = min_q; return 0; }
            }
            if (min_p != 0) {
                ncand++;
                double cost = (double)(len - 1) * (double)(min_len - 1);
                if (cost < best) { p = min_p; q = min_q; best = cost; }
                if (ncand == L.piv_lim) return 0;
            } else
                assert(min_p != 0);
        }
    }
    return p == 0;
}

/* lib/glpluf.js:637-967 eliminate: one step of right-looking elimination with
   pivot v[p,q]; multipliers go to column p of F in the SVA tail */
static int eliminate(LUF &L, int p, int q)
{
    const int n = L.n;
    int *ndx = L.fr_len.data(); /* free during factorisation */
    rs_remove(L, p);
    cs_remove(L, q);
    int p_beg = L.vr_ptr[p], p_end = p_beg + L.vr_len[p] - 1, p_ptr;
    for (p_ptr = p_beg; L.sv_ind[p_ptr] != q; p_ptr++) {}
    assert(p_ptr <= p_end);
    double vpq = (L.vr_piv[p] = L.sv_val[p_ptr]);
    L.sv_ind[p_ptr] = L.sv_ind[p_end];
    L.sv_val[p_ptr] = L.sv_val[p_end];
    L.vr_len[p]--; p_end--;
    int q_beg = L.vc_ptr[q], q_end = q_beg + L.vc_len[q] - 1, q_ptr;
    for (q_ptr = q_beg; L.sv_ind[q_ptr] != p; q_ptr++) {}
    assert(q_ptr <= q_end);
    L.sv_ind[q_ptr] = L.sv_ind[q_end];
    L.vc_len[q]--; q_end--;
    /* scatter the pivot row, detach its columns from the active set */
    for (p_ptr = p_beg; p_ptr <= p_end; p_ptr++) {
        int j = L.sv_ind[p_ptr];
        L.flag[j] = 1;
        L.work[j] = L.sv_val[p_ptr];
        cs_remove(L, j);
        int jb = L.vc_ptr[j], je = jb + L.vc_len[j] - 1, jp;
        for (jp = jb; L.sv_ind[jp] != p; jp++) {}
        assert(jp <= je);
        L.sv_ind[jp] = L.sv_ind[je];
        L.vc_len[j]--;
    }
    while (q_beg <= q_end) {
        int i = L.sv_ind[q_beg];
        rs_remove(L, i);
        int i_beg = L.vr_ptr[i], i_end = i_beg + L.vr_len[i] - 1, i_ptr;
        for (i_ptr = i_beg; L.sv_ind[i_ptr] != q; i_ptr++) {}
        assert(i_ptr <= i_end);
        double fip = L.sv_val[i_ptr] / vpq;
        L.sv_ind[i_ptr] = L.sv_ind[i_end];
        L.sv_val[i_ptr] = L.sv_val[i_end];
        L.vr_len[i]--; i_end--;
        L.sv_ind[q_beg] = L.sv_ind[q_end];
        L.vc_len[q]--; q_end--;
        int fill = L.vr_len[p];
        for (i_ptr = i_beg; i_ptr <= i_end; i_ptr++) {
            int j = L.sv_ind[i_ptr];
            if (L.flag[j]) {
                double temp = (L.sv_val[i_ptr] -= fip * L.work[j]);
                if (temp < 0.0) temp = -temp;
                L.flag[j] = 0;
                fill--;
                if (temp == 0.0 || temp < L.eps_tol) {
                    L.sv_ind[i_ptr] = L.sv_ind[i_end];
                    L.sv_val[i_ptr] = L.sv_val[i_end];
                    L.vr_len[i]--; i_ptr--; i_end--;
                    int jb = L.vc_ptr[j], je = jb + L.vc_len[j] - 1, jp;
                    for (jp = jb; L.sv_ind[jp] != i; jp++) {}
                    assert(jp <= je);
                    L.sv_ind[jp] = L.sv_ind[je];
                    L.vc_len[j]--;
                } else if (L.big_v < temp)
                    L.big_v = temp;
            }
        }
        if (L.vr_len[i] + fill > L.vr_cap[i]) {
            if (enlarge_row(L, i, L.vr_len[i] + fill)) return 1;
            p_beg = L.vr_ptr[p]; p_end = p_beg + L.vr_len[p] - 1;
            q_beg = L.vc_ptr[q]; q_end = q_beg + L.vc_len[q] - 1;
        }
        int len = 0;
        for (p_ptr = p_beg; p_ptr <= p_end; p_ptr++) {
            int j = L.sv_ind[p_ptr];
            if (L.flag[j]) {
                double val = -fip * L.work[j];
                double temp = val < 0.0 ? -val : val;
                if (!(temp == 0.0 || temp < L.eps_tol)) {
                    int ip = L.vr_ptr[i] + L.vr_len[i];
                    L.sv_ind[ip] = j;
                    L.sv_val[ip] = val;
                    L.vr_len[i]++;
                    ndx[++len] = j;
                    if (L.big_v < temp) L.big_v = temp;
                }
            } else
                L.flag[j] = 1;
        }
        for (int k = 1; k <= len; k++) {
            int j = ndx[k];
            if (L.vc_len[j] + 1 > L.vc_cap[j]) {
                if (enlarge_col(L, j, L.vc_len[j] + 10)) return 1;
                p_beg = L.vr_ptr[p]; p_end = p_beg + L.vr_len[p] - 1;
                q_beg = L.vc_ptr[q]; q_end = q_beg + L.vc_len[q] - 1;
            }
            L.sv_ind[L.vc_ptr[j] + L.vc_len[j]] = i;
            L.vc_len[j]++;
        }
        rs_insert(L, i);
        L.vr_max[i] = -1.0;
        if (L.sv_end - L.sv_beg < 1) {
            defrag_sva(L);
            if (L.sv_end - L.sv_beg < 1) return 1;
            p_beg = L.vr_ptr[p]; p_end = p_beg + L.vr_len[p] - 1;
            q_beg = L.vc_ptr[q]; q_end = q_beg + L.vc_len[q] - 1;
        }
        L.sv_end--;
        L.sv_ind[L.sv_end] = i;
        L.sv_val[L.sv_end] = fip;
        L.fc_len[p]++;
    }
    assert(L.vc_len[q] == 0);
    L.vc_cap[q] = 0;
    int k = n + q;
    if (L.sv_prev[k] == 0) L.sv_head = L.sv_next[k];
    else L.sv_next[L.sv_prev[k]] = L.sv_next[k];
    if (L.sv_next[k] == 0) L.sv_tail = L.sv_prev[k];
    else L.sv_prev[L.sv_next[k]] = L.sv_prev[k];
    L.fc_ptr[p] = L.sv_end;
    for (p_ptr = p_beg; p_ptr <= p_end; p_ptr++) {
        int j = L.sv_ind[p_ptr];
        L.flag[j] = 0;
        L.work[j] = 0.0;
        if (!(L.vc_len[j] != 1 && L.cs_prev[j] == j && L.cs_next[j] == j))
            cs_insert(L, j);
    }
    return 0;
}

/* lib/glpluf.js:969-1043 build_v_cols */
static int build_v_cols(LUF &L)
{
    const int n = L.n;
    int nnz = 0;
    for (int i = 1; i <= n; i++) {
        int ib = L.vr_ptr[i], ie = ib + L.vr_len[i] - 1;
        for (int t = ib; t <= ie; t++) L.vc_cap[L.sv_ind[t]]++;
        nnz += L.vr_len[i];
    }
    L.nnz_v = nnz;
    if (L.sv_end - L.sv_beg < nnz) return 1;
    for (int j = 1; j <= n; j++) { L.vc_ptr[j] = L.sv_beg; L.sv_beg += L.vc_cap[j]; }
    for (int i = 1; i <= n; i++) {
        int ib = L.vr_ptr[i], ie = ib + L.vr_len[i] - 1;
        for (int t = ib; t <= ie; t++) {
            int j = L.sv_ind[t];
            int jp = L.vc_ptr[j] + L.vc_len[j];
            L.sv_ind[jp] = i;
            L.sv_val[jp] = L.sv_val[t];
            L.vc_len[j]++;
        }
    }
    for (int k = n + 1; k <= n + n; k++) { L.sv_prev[k] = k - 1; L.sv_next[k] = k + 1; }
    L.sv_prev[n + 1] = L.sv_tail;
    L.sv_next[L.sv_tail] = n + 1;
    L.sv_next[n + n] = 0;
    L.sv_tail = n + n;
    return 0;
}

/* lib/glpluf.js:1045-1103 build_f_rows */
static int build_f_rows(LUF &L)
{
    const int n = L.n;
    for (int i = 1; i <= n; i++) L.fr_len[i] = 0;
    int nnz = 0;
    for (int j = 1; j <= n; j++) {
        int jb = L.fc_ptr[j], je = jb + L.fc_len[j] - 1;
        for (int t = jb; t <= je; t++) L.fr_len[L.sv_ind[t]]++;
        nnz += L.fc_len[j];
    }
    L.nnz_f = nnz;
    if (L.sv_end - L.sv_beg < nnz) return 1;
    for (int i = 1; i <= n; i++) { L.fr_ptr[i] = L.sv_end; L.sv_end -= L.fr_len[i]; }
    for (int j = 1; j <= n; j++) {
        int jb = L.fc_ptr[j], je = jb + L.fc_len[j] - 1;
        for (int t = jb; t <= je; t++) {
            int i = L.sv_ind[t];
            int ptr = --L.fr_ptr[i];
            L.sv_ind[ptr] = j;
            L.sv_val[ptr] = L.sv_val[t];
        }
    }
    return 0;
}

/* lib/glpluf.js:1105-1225 luf_factorize: retry loop doubling the SVA */
int luf_factorize(LUF &L, int n, col_fn col, void *info)
{
    assert(n >= 1);
    L.valid = 0;
    reallocate(L, n);
    if (L.sv_size == 0 && L.new_sva == 0) L.new_sva = 5 * (n + 10);
    for (;;) {
        if (L.new_sva > 0) {
            L.sv_size = L.new_sva;
            L.sv_ind.assign(1 + L.sv_size, 0);
            L.sv_val.assign(1 + L.sv_size, 0.0);
            L.new_sva = 0;
        }
        bool again = false;
        if (luf_initialize(L, col, info)) {
            L.new_sva = L.sv_size + L.sv_size;
            continue;
        }
        for (int k = 1; k <= n; k++) {
            int p, q;
            if (find_pivot(L, p, q)) { L.rank = k - 1; return 1 /*LUF_ESING*/; }
            int i = L.pp_col[p], j = L.qq_row[q];
            assert(k <= i && i <= n && k <= j && j <= n);
            int t = L.pp_row[k];
            L.pp_row[i] = t; L.pp_col[t] = i;
            L.pp_row[k] = p; L.pp_col[p] = k;
            t = L.qq_col[k];
            L.qq_col[j] = t; L.qq_row[t] = j;
            L.qq_col[k] = q; L.qq_row[q] = k;
            if (eliminate(L, p, q)) {
                L.new_sva = L.sv_size + L.sv_size;
                again = true;
                break;
            }
            if (L.big_v > L.max_gro * L.max_a) { L.rank = k - 1; return 2 /*LUF_ECOND*/; }
        }
        if (again) continue;
        defrag_sva(L);
        if (build_v_cols(L)) { L.new_sva = L.sv_size + L.sv_size; continue; }
        if (build_f_rows(L)) { L.new_sva = L.sv_size + L.sv_size; continue; }
        break;
    }
    L.valid = 1;
    L.rank = n;
    int t = 3 * (n + L.nnz_v) + 2 * L.nnz_f;
    if (L.sv_size < t) {
        L.new_sva = L.sv_size;
        while (L.new_sva < t) L.new_sva += L.new_sva;
    }
    return 0;
}

/* lib/glpluf.js:1227-1266 luf_f_solve; pp_row is passed explicitly because
   the FT driver substitutes P0 (lib/glpfhv.js:122-126) */
void luf_f_solve(LUF &L, int tr, double *x, const int *pp_row)
{
    const int n = L.n;
    if (!tr) {
        for (int j = 1; j <= n; j++) {
            int k = pp_row[j];
            double xk = x[k];
            if (xk != 0.0) {
                int beg = L.fc_ptr[k], end = beg + L.fc_len[k] - 1;
                for (int t = beg; t <= end; t++) x[L.sv_ind[t]] -= L.sv_val[t] * xk;
            }
        }
    } else {
        for (int i = n; i >= 1; i--) {
            int k = pp_row[i];
            double xk = x[k];
            if (xk != 0.0) {
                int beg = L.fr_ptr[k], end = beg + L.fr_len[k] - 1;
                for (int t = beg; t <= end; t++) x[L.sv_ind[t]] -= L.sv_val[t] * xk;
            }
        }
    }
}

/* lib/glpluf.js:1268-1313 luf_v_solve */
void luf_v_solve(LUF &L, int tr, double *x)
{
    const int n = L.n;
    double *b = L.work.data();
    for (int k = 1; k <= n; k++) { b[k] = x[k]; x[k] = 0.0; }
    if (!tr) {
        for (int k = n; k >= 1; k--) {
            int i = L.pp_row[k], j = L.qq_col[k];
            double temp = b[i];
            if (temp != 0.0) {
                x[j] = (temp /= L.vr_piv[i]);
                int beg = L.vc_ptr[j], end = beg + L.vc_len[j] - 1;
                for (int t = beg; t <= end; t++) b[L.sv_ind[t]] -= L.sv_val[t] * temp;
            }
        }
    } else {
        for (int k = 1; k <= n; k++) {
            int i = L.pp_row[k], j = L.qq_col[k];
            double temp = b[j];
            if (temp != 0.0) {
                x[i] = (temp /= L.vr_piv[i]);
                int beg = L.vr_ptr[i], end = beg + L.vr_len[i] - 1;
                for (int t = beg; t <= end; t++) b[L.sv_ind[t]] -= L.sv_val[t] * temp;
            }
        }
    }
}

/* ------------------------------------------------------------------ */
/* Forrest-Tomlin                                                     */
/* ------------------------------------------------------------------ */

/* lib/glpfhv.js:77-112 fhv_h_solve */
static void h_solve(BFD &B, int tr, double *x)
{
    LUF &L = B.luf;
    if (!tr) {
        for (int k = 1; k <= B.hh_nfs; k++) {
            int i = B.hh_ind[k];
            double temp = x[i];
            int beg = B.hh_ptr[k], end = beg + B.hh_len[k] - 1;
            for (int t = beg; t <= end; t++) temp -= L.sv_val[t] * x[L.sv_ind[t]];
            x[i] = temp;
        }
    } else {
        for (int k = B.hh_nfs; k >= 1; k--) {
            int i = B.hh_ind[k];
            double temp = x[i];
            if (temp == 0.0) continue;
            int beg = B.hh_ptr[k], end = beg + B.hh_len[k] - 1;
            for (int t = beg; t <= end; t++) x[L.sv_ind[t]] -= L.sv_val[t] * temp;
        }
    }
}

/* lib/glpbfd.js:47-146 bfd_factorize + lib/glpfhv.js:26-75 fhv_factorize */
int bfd_factorize(BFD &B, int m, col_fn col, void *info)
{
    B.valid = 0;
    LUF &L = B.luf;
    if (B.fresh) { L.new_sva = B.lu_size; B.hh_max = B.nfs_max; B.fresh = false; }
    L.piv_tol = B.piv_tol; L.piv_lim = B.piv_lim; L.suhl = B.suhl;
    L.eps_tol = B.eps_tol; L.max_gro = B.max_gro;
    B.m = m;
    if ((int)B.hh_ind.size() < 1 + B.hh_max) {
        B.hh_ind.assign(1 + B.hh_max, 0);
        B.hh_ptr.assign(1 + B.hh_max, 0);
        B.hh_len.assign(1 + B.hh_max, 0);
    }
    if ((int)B.p0_row.size() < 1 + m) {
        B.p0_row.assign(1 + m + 100, 0); B.p0_col.assign(1 + m + 100, 0);
        B.cc_ind.assign(1 + m + 100, 0); B.cc_val.assign(1 + m + 100, 0.0);
    }
    B.n_factorize++;
    int ret = luf_factorize(L, m, col, info);
    if (ret == 1) return BFD_ESING;
    if (ret == 2) return BFD_ECOND;
    B.hh_nfs = 0;
    memcpy(&B.p0_row[1], &L.pp_row[1], m * sizeof(int));
    memcpy(&B.p0_col[1], &L.pp_col[1], m * sizeof(int));
    B.nnz_h = 0;
    B.valid = 1;
    B.upd_cnt = 0;
    return 0;
}

/* lib/glpfhv.js:114-129 fhv_ftran: x := inv(V) inv(H) inv(F) x, F with P0 */
void bfd_ftran(BFD &B, double *x)
{
    assert(B.valid);
    B.n_ftran++;
    luf_f_solve(B.luf, 0, x, B.p0_row.data());
    h_solve(B, 0, x);
    luf_v_solve(B.luf, 0, x);
}

/* lib/glpfhv.js:131-146 fhv_btran */
void bfd_btran(BFD &B, double *x)
{
    assert(B.valid);
    B.n_btran++;
    luf_v_solve(B.luf, 1, x);
    h_solve(B, 1, x);
    luf_f_solve(B.luf, 1, x, B.p0_row.data());
}

/* lib/glpfhv.js:148-447 fhv_update_it (+ error mapping lib/glpbfd.js:170-223):
   replace column j of B by the sparse column (ind[idx+1..idx+len], val[1..len]) */
int bfd_update_it(BFD &B, int j, int len, const int *ind, int idx, const double *val)
{
    assert(B.valid);
    const int m = B.m;
    LUF &L = B.luf;
    int *cc_ind = B.cc_ind.data();
    double *cc_val = B.cc_val.data();
    double *work = L.work.data();
    B.n_update++;
    if (B.hh_nfs == B.hh_max) { B.valid = 0; return BFD_ELIMIT; }
    for (int i = 1; i <= m; i++) cc_val[i] = 0.0;
    for (int k = 1; k <= len; k++) {
        int i = ind[idx + k];
        assert(1 <= i && i <= m && cc_val[i] == 0.0 && val[k] != 0.0);
        cc_val[i] = val[k];
    }
    /* new column of V := inv(F*H) * (new column of B) */
    luf_f_solve(L, 0, cc_val, B.p0_row.data());
    h_solve(B, 0, cc_val);
    len = 0;
    for (int i = 1; i <= m; i++) {
        double temp = cc_val[i];
        if (temp == 0.0 || fabs(temp) < L.eps_tol) continue;
        len++; cc_ind[len] = i; cc_val[len] = temp;
    }
    /* clear old column j of V from the row lists */
    {
        int jb = L.vc_ptr[j], je = jb + L.vc_len[j] - 1;
        for (int jp = jb; jp <= je; jp++) {
            int i = L.sv_ind[jp];
            int ib = L.vr_ptr[i], ie = ib + L.vr_len[i] - 1, ip;
            for (ip = ib; L.sv_ind[ip] != j; ip++) {}
            assert(ip <= ie);
            L.sv_ind[ip] = L.sv_ind[ie];
            L.sv_val[ip] = L.sv_val[ie];
            L.vr_len[i]--;
        }
    }
    L.nnz_v -= L.vc_len[j];
    L.vc_len[j] = 0;
    int k1 = L.qq_row[j], k2 = 0;
    for (int t = 1; t <= len; t++) {
        int i = cc_ind[t];
        if (L.vr_len[i] + 1 > L.vr_cap[i]) {
            if (enlarge_row(L, i, L.vr_len[i] + 10)) {
                B.valid = 0; L.new_sva = L.sv_size + L.sv_size; return BFD_EROOM;
            }
        }
        int ip = L.vr_ptr[i] + L.vr_len[i];
        L.sv_ind[ip] = j;
        L.sv_val[ip] = cc_val[t];
        L.vr_len[i]++;
        if (k2 < L.pp_col[i]) k2 = L.pp_col[i];
    }
    if (L.vc_cap[j] < len) {
        if (enlarge_col(L, j, len)) {
            B.valid = 0; L.new_sva = L.sv_size + L.sv_size; return BFD_EROOM;
        }
    }
    memcpy(&L.sv_ind[L.vc_ptr[j]], &cc_ind[1], len * sizeof(int));
    memcpy(&L.sv_val[L.vc_ptr[j]], &cc_val[1], len * sizeof(double));
    L.vc_len[j] = len;
    L.nnz_v += len;
    if (k1 > k2) { B.valid = 0; return BFD_ESING; }
    /* cyclic shift of rows/cols k1..k2 of U */
    int i = L.pp_row[k1];
    j = L.qq_col[k1];
    for (int k = k1; k < k2; k++) {
        L.pp_row[k] = L.pp_row[k + 1]; L.pp_col[L.pp_row[k]] = k;
        L.qq_col[k] = L.qq_col[k + 1]; L.qq_row[L.qq_col[k]] = k;
    }
    L.pp_row[k2] = i; L.pp_col[i] = k2;
    L.qq_col[k2] = j; L.qq_row[j] = k2;
    /* row i of V (= row k2 of U) goes to the dense work array */
    for (int jj = 1; jj <= m; jj++) work[jj] = 0.0;
    {
        int ib = L.vr_ptr[i], ie = ib + L.vr_len[i] - 1;
        for (int ip = ib; ip <= ie; ip++) {
            int jj = L.sv_ind[ip];
            work[jj] = L.sv_val[ip];
            int jb = L.vc_ptr[jj], je = jb + L.vc_len[jj] - 1, jp;
            for (jp = jb; L.sv_ind[jp] != i; jp++) {}
            assert(jp <= je);
            L.sv_ind[jp] = L.sv_ind[je];
            L.sv_val[jp] = L.sv_val[je];
            L.vc_len[jj]--;
        }
    }
    L.nnz_v -= L.vr_len[i];
    L.vr_len[i] = 0;
    B.hh_nfs++;
    B.hh_ind[B.hh_nfs] = i;
    B.hh_len[B.hh_nfs] = 0;
    if (L.sv_end - L.sv_beg < k2 - k1) {
        defrag_sva(L);
        if (L.sv_end - L.sv_beg < k2 - k1) {
            B.valid = L.valid = 0;
            L.new_sva = L.sv_size + L.sv_size;
            return BFD_EROOM;
        }
    }
    /* eliminate the row spike */
    for (int k = k1; k < k2; k++) {
        int p = L.pp_row[k], q = L.qq_col[k];
        if (work[q] == 0.0) continue;
        double f = work[q] / L.vr_piv[p];
        int pb = L.vr_ptr[p], pe = pb + L.vr_len[p] - 1;
        for (int pp = pb; pp <= pe; pp++) work[L.sv_ind[pp]] -= f * L.sv_val[pp];
        L.sv_end--;
        L.sv_ind[L.sv_end] = p;
        L.sv_val[L.sv_end] = f;
        B.hh_len[B.hh_nfs]++;
    }
    if (B.hh_len[B.hh_nfs] == 0)
        B.hh_nfs--;
    else {
        B.hh_ptr[B.hh_nfs] = L.sv_end;
        B.nnz_h += B.hh_len[B.hh_nfs];
    }
    L.vr_piv[i] = work[L.qq_col[k2]];
    len = 0;
    for (int k = k2 + 1; k <= m; k++) {
        int jj = L.qq_col[k];
        double temp = work[jj];
        if (fabs(temp) < L.eps_tol) continue;
        if (L.vc_len[jj] + 1 > L.vc_cap[jj]) {
            if (enlarge_col(L, jj, L.vc_len[jj] + 10)) {
                B.valid = 0; L.new_sva = L.sv_size + L.sv_size; return BFD_EROOM;
            }
        }
        int jp = L.vc_ptr[jj] + L.vc_len[jj];
        L.sv_ind[jp] = i;
        L.sv_val[jp] = temp;
        L.vc_len[jj]++;
        len++; cc_ind[len] = jj; cc_val[len] = temp;
    }
    if (L.vr_cap[i] < len) {
        if (enlarge_row(L, i, len)) {
            B.valid = 0; L.new_sva = L.sv_size + L.sv_size; return BFD_EROOM;
        }
    }
    memcpy(&L.sv_ind[L.vr_ptr[i]], &cc_ind[1], len * sizeof(int));
    memcpy(&L.sv_val[L.vr_ptr[i]], &cc_val[1], len * sizeof(double));
    L.vr_len[i] = len;
    L.nnz_v += len;
    /* stability check of the new diagonal element */
    double temp = 0.0;
    i = L.pp_row[k2];
    {
        int ib = L.vr_ptr[i], ie = ib + L.vr_len[i] - 1;
        for (int ip = ib; ip <= ie; ip++)
            if (temp < fabs(L.sv_val[ip])) temp = fabs(L.sv_val[ip]);
    }
    j = L.qq_col[k2];
    {
        int jb = L.vc_ptr[j], je = jb + L.vc_len[j] - 1;
        for (int jp = jb; jp <= je; jp++)
            if (temp < fabs(L.sv_val[jp])) temp = fabs(L.sv_val[jp]);
    }
    if (fabs(L.vr_piv[i]) < B.upd_tol * temp) { B.valid = 0; return BFD_ECHECK; }
    B.upd_cnt++;
    return 0;
}

} /* namespace glpo */
