/* csa.h -- CPU ORACLE (test infrastructure only; see glpo.h).
 * Common storage area shared by the primal and dual restatements
 * (lib/glpspx01.js:5-40 alloc_csa, lib/glpspx02.js:5-43 alloc_csa).
 * All arrays 1-based.
 */
#ifndef GLPO_CSA_H
#define GLPO_CSA_H
#include "glpo.h"

namespace glpo {

struct CSA {
    int m, n;
    std::vector<signed char> type, orig_type;
    std::vector<double> lb, ub, coef, orig_lb, orig_ub, obj;
    std::vector<int> A_ptr, A_ind;
    std::vector<double> A_val;
    std::vector<int> AT_ptr, AT_ind;        /* dual only */
    std::vector<double> AT_val;
    std::vector<int> head, bind;            /* bind: dual only */
    std::vector<signed char> stat;
    std::vector<int> N_ptr, N_len, N_ind;   /* primal only: row-wise N */
    std::vector<double> N_val;
    std::vector<double> bbar, cbar;
    std::vector<signed char> refsp;
    std::vector<double> gamma;
    std::vector<int> tcol_ind, trow_ind;
    std::vector<double> tcol_vec, trow_vec;
    std::vector<double> work1, work2, work3, work4;
    double zeta;
    int valid;
    BFD *bfd;
    int phase;
    double tm_beg;
    int it_beg, it_cnt, it_dpy;
    int refct;
    int p, p_stat, q;
    double teta, delta, new_dq;
    int tcol_nnz, tcol_num, trow_nnz, trow_num;
    double tcol_max, trow_max;
    /* scratch for hooks: arguments of the selection call being reported */
    double hook_tol;
};

} /* namespace glpo */
#endif
