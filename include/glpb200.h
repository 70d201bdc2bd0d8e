/* glpb200.h -- C ABI of libglpb200.so, the B200-native (sm_100a) replacement
 * for the simplex hot path of glpk.js (JavaScript port of GLPK 4.49).
 *
 * This is the drop-in boundary.  A host binding (N-API addon,
 * ctypes, cgo ...) keeps one handle per glp_prob; the handle owns all
 * device-resident problem data: CSC and CSR copies of the scaled constraint
 * matrix, bounds, costs, the basis header, steepest-edge weights and the
 * basis inverse.  Every entry point is blocking, never throws, and retains no
 * caller pointer after it returns.  There is NO CPU fallback: without a CUDA
 * device every compute call fails with GLPB_ENODEV.
 *
 * Index conventions at this boundary are the reference's: variables are
 * numbered k = 1..m+n (auxiliary rows first), basis header entries are such
 * k, statuses/types/return codes are the GLP_* values of lib/glpk.js.
 * Arrays are plain 0-based C arrays (element [k-1] belongs to variable k)
 * unless a parameter is documented as "CSA layout" (1-based, slot 0 unused,
 * exactly as lib/glpspx01.js / lib/glpspx02.js keep them).
 */
#ifndef GLPB200_H
#define GLPB200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- API misuse / environment errors (negative; GLP_* codes are >= 0) ---- */
#define GLPB_EINVAL  (-1)   /* invalid argument (the JS facade throws)        */
#define GLPB_ENODEV  (-2)   /* no CUDA device / CUDA runtime error            */
#define GLPB_ENOMEM  (-3)   /* device allocation failed                      */
#define GLPB_ESTATE  (-4)   /* call not valid in the handle's current state   */

typedef struct glpb_prob glpb_prob;

/* glp_smcp, field for field: lib/glpapi06.js:359-375 (SMCP constructor) */
typedef struct glpb_smcp {
    int msg_lev, meth, pricing, r_test;
    double tol_bnd, tol_dj, tol_piv, obj_ll, obj_ul;
    int it_lim, tm_lim, out_frq, out_dly, presolve;
} glpb_smcp;

/* glp_iocp: lib/glpapi09.js:392-414 (IOCP constructor); node_lim is an
 * extension used only by throughput benchmarks (-1 = off) */
typedef struct glpb_iocp {
    int msg_lev, br_tech, bt_tech;
    double tol_int, tol_obj;
    int tm_lim, out_frq, out_dly, pp_tech;
    double mip_gap;
    int presolve;
    long node_lim;
} glpb_iocp;

/* glp_bfcp subset: lib/glpapi12.js:108-122.  The device keeps an explicit
 * inverse of the structural kernel of B instead of F*H*V, so only the update
 * count (nfs_max -> refactorisation period) and the pivot tolerances apply. */
typedef struct glpb_bfcp {
    int nfs_max;        /* refactorise after max(nfs_max, 4k) updates, k = basic structurals (default 100;
                           the explicit inverse has no eta file to outgrow: the period only bounds
                           rounding, the reference's accuracy triggers still force a fresh start;
                           GLPB_REFAC_AUTO=0 in the environment pins the period to nfs_max) */
    double piv_tol;     /* relative pivot threshold in the dense inverse      */
    double upd_tol;     /* reserved                                           */
} glpb_bfcp;

void glpb_init_smcp(glpb_smcp *parm);   /* replaces: new SMCP()  api06:359 */
void glpb_init_iocp(glpb_iocp *parm);   /* replaces: new IOCP()  api09:392 */

/* Library / device probes. */
int glpb_device_count(void);
const char *glpb_last_error(void);
const char *glpb_version(void);

/* Create a device-resident problem.  Replaces what alloc_csa/init_csa copy
 * out of the glp_prob (lib/glpspx01.js:5-145, lib/glpspx02.js:5-190):
 *   dir GLP_MIN/GLP_MAX, c0 constant term,
 *   type/lb/ub [m+n] (rows then columns, UNSCALED), coef[n] objective,
 *   kind[n] GLP_CV/GLP_IV (may be NULL), rii[m], sjj[n] scale factors (NULL=1),
 *   A in CSC: A_ptr[n+1], A_ind[nnz] 0-based rows, A_val[nnz] UNSCALED, each
 *   column in the reference's list order (col.ptr -> c_next).
 * Returns NULL on failure (see glpb_last_error). */
glpb_prob *glpb_create(int m, int n, int nnz, int dir, double c0,
                       const int *type, const double *lb, const double *ub,
                       const double *coef, const int *kind,
                       const double *rii, const double *sjj,
                       const int *A_ptr, const int *A_ind, const double *A_val,
                       int device);
void glpb_destroy(glpb_prob *P);                     /* idempotent on NULL */

/* glp_set_row_bnds / glp_set_col_bnds for a list of variables k (1..m+n),
 * api01:217-281: non-basic statuses are re-derived, the basis stays valid. */
int glpb_set_bounds(glpb_prob *P, int count, const int *k, const int *type,
                    const double *lb, const double *ub);
/* glp_set_row_stat / glp_set_col_stat for all variables, api05:1-47 */
int glpb_set_basis(glpb_prob *P, const int *stat /* [m+n] */);
int glpb_std_basis(glpb_prob *P);                    /* api05:49-63 */
int glpb_set_bfcp(glpb_prob *P, const glpb_bfcp *parm);
int glpb_set_it_cnt(glpb_prob *P, int it_cnt);

/* glp_factorize, api12:5-100: 0 / GLP_EBADB / GLP_ESING / GLP_ECOND */
int glpb_factorize(glpb_prob *P);
/* glp_simplex with presolve OFF, api06:261-339 -> solve_lp api06:3-39 ->
 * spx_primal / spx_dual.  Returns 0 or GLP_EBADB/ESING/ECOND/EBOUND/EFAIL/
 * EOBJLL/EOBJUL/EITLIM/ETMLIM. */
int glpb_simplex(glpb_prob *P, const glpb_smcp *parm);
/* glp_intopt with presolve OFF, api09:61-114 -> ios_driver (ios03:507-951):
 * 0 / GLP_EROOT / GLP_EFAIL / GLP_EMIPGAP / GLP_ETMLIM / GLP_ESTOP */
int glpb_intopt(glpb_prob *P, const glpb_iocp *parm);

/* What store_sol writes back (lib/glpspx01.js:1591-1681).  Any pointer may be
 * NULL.  stat/prim/dual are [m+n]; head is [m] (values k = 1..m+n). */
int glpb_get_solution(glpb_prob *P, int *stat, double *prim, double *dual,
                      int *head, int *pbs_stat, int *dbs_stat, double *obj_val,
                      int *it_cnt, int *some);
int glpb_get_status(glpb_prob *P);                   /* api06:398-427 */
/* record_solution (ios03:118-139): mipx is [m+n] */
int glpb_get_mip(glpb_prob *P, int *mip_stat, double *mip_obj, double *mipx,
                 long *nodes);

/* Resumable branch-and-bound for multi-GPU node sharding.
 * glpb_intopt == begin + run(-1) + end.  Between slices a host exchanges the
 * incumbent objective and migrates open nodes between ranks; a node record is
 * self-contained (glpb_mip_record_bytes bytes: bound, lp_obj, level and the
 * complete type/stat/lb/ub vectors).
 *   run:  0 = local pool exhausted, 1 = slice limit reached, GLP_E* = stopped */
int glpb_mip_begin(glpb_prob *P, const glpb_iocp *parm);
int glpb_mip_run(glpb_prob *P, long max_nodes, long *solved);
int glpb_mip_get_incumbent(glpb_prob *P, int *has_solution, double *obj);
int glpb_mip_set_cutoff(glpb_prob *P, double obj);
int glpb_mip_open_count(glpb_prob *P);
long glpb_mip_record_bytes(glpb_prob *P);
int glpb_mip_export_nodes(glpb_prob *P, int max_count, void *buf, long cap, int *count);
int glpb_mip_import_nodes(glpb_prob *P, const void *buf, int count);
int glpb_mip_end(glpb_prob *P, int ret);

/* Batched, device-resident branch-and-bound (replaces the serial
 * ios_driver loop, lib/glpios03.js:507-951, for problems whose node LP fits one
 * CTA: m <= 64, n <= 1024).  The open nodes' states live in a device slab; one
 * round = ONE kernel launch that takes up to `batch` open nodes, one CTA each,
 * through ios_preprocess_node, ios_solve_node (warm-started dual simplex out of
 * shared memory), ios_round_bound, check_integrality, fix_by_red_cost, the
 * Driebeck-Tomlin choice and branch_on; the host only keeps the tree.
 * glpb_intopt takes this path by itself for eligible problems (environment
 * GLPB_BNB=serial keeps the one-LP-at-a-time driver).
 *   begin:  batch <= 0 -> 16 x SM count (GLPB_BNB_BATCH), slab_nodes <= 0 ->
 *           262144 (GLPB_BNB_SLAB); the root LP must be solved (GLP_EROOT).
 *   round:  1 = a round was run (*done nodes), 0 = local pool empty,
 *           GLP_ETMLIM / GLP_ESTOP (node_lim) / GLP_EMIPGAP / GLP_EFAIL.
 *   export/import: self-contained node records in DEVICE memory of the
 *           handle's GPU (glpb_bnb_record_bytes each) for migration between
 *           ranks, e.g. through an NCCL all-gather; set_cutoff installs an
 *           incumbent objective found elsewhere.
 *   stats:  [0] node LPs solved (= ios_solve_node calls), [1] nodes processed,
 *           [2] rounds, [3] dual simplex iterations, [4] basis inversions,
 *           [5] open nodes, [6] shared-memory bytes per CTA, [7] matrix in smem.
 *   end:    solve_mip's status mapping (lib/glpapi09.js:82-112); returns ret. */
int glpb_bnb_begin(glpb_prob *P, const glpb_iocp *parm, int batch, int slab_nodes);
int glpb_bnb_round(glpb_prob *P, long max_tasks, long *done);
int glpb_bnb_open_count(glpb_prob *P);
int glpb_bnb_get_incumbent(glpb_prob *P, int *has_solution, double *obj);
int glpb_bnb_set_cutoff(glpb_prob *P, double obj);
int glpb_bnb_clear(glpb_prob *P);     /* drop all open nodes: a rank that starts empty and is fed by migration */
long glpb_bnb_record_bytes(glpb_prob *P);
int glpb_bnb_export_nodes(glpb_prob *P, int max_count, void *dev_buf, int *count);
int glpb_bnb_import_nodes(glpb_prob *P, const void *dev_buf, int count);
int glpb_bnb_stats(glpb_prob *P, long *out, int count);
int glpb_bnb_end(glpb_prob *P, int ret);

/* Counters for measurement: out[0] iterations, [1] refactorisations,
 * [2] kernel launches, [3] host<->device syncs, [4] basis updates,
 * [5] current kernel size k, [6] device microseconds in the last solve,
 * [7] CUDA-graph replays (fixed launch sequences of small LPs; their kernels
 * are counted in [2]). */
int glpb_get_counters(glpb_prob *P, long *out, int count);

/* Per-kernel device time, measured with CUDA events on the handle's stream
 * (bench.py roofline leg).  The report is text, one line per kernel:
 * "name launches total_ms algorithmic_bytes". */
int glpb_set_profile(glpb_prob *P, int on);
const char *glpb_profile_report(glpb_prob *P);

/* Pivot log (parity tests of the iteration engine): after glpb_set_pivot_log(P, cap)
 * every simplex iteration whose number is below cap records the entering / leaving
 * pair the reference keeps in csa.q / csa.p (lib/glpspx01.js:30-32: q = 1..n in the
 * non-basic list, p = 1..m in the basis header, p = -1 for a bound flip).
 * glpb_get_pivot_log copies min(cap, iterations done) pairs: qp[2 it] = q, qp[2 it + 1] = p. */
int glpb_set_pivot_log(glpb_prob *P, int cap);
int glpb_get_pivot_log(glpb_prob *P, int *qp, int cap, int *count);
/* Live device vectors for parity tests (what update_gamma / update_cbar / update_bbar left,
 * lib/glpspx01.js:1100-1255, lib/glpspx02.js:1020-1188): name = "gamma" (primal [n] by non-basic
 * position, dual [m] by basic position), "cbar" [n], "bbar" [m], "head" [m+n] (1-based variable
 * numbers), "stat" [n]; 0-based arrays of `count` doubles. */
int glpb_debug_get(glpb_prob *P, const char *name, double *out, int count);

/* Basis solves with the current factorisation, scaled space:
 * bfd_ftran / bfd_btran (lib/glpbfd.js:148-168); x is [m], in place. */
int glpb_ftran(glpb_prob *P, double *x);
int glpb_btran(glpb_prob *P, double *x);

/* ---- kernel-level entry points (parity tests feed them the reference's
 *      arrays; all arrays CSA layout = 1-based, lengths as in alloc_csa) ---- */

/* chuzc, lib/glpspx01.js:646-688: returns q (0 = none) */
int glpb_k_chuzc_primal(int n, const signed char *stat, const double *cbar,
                        const double *gamma, double tol_dj, int *q);
/* chuzr, lib/glpspx02.js:572-625: returns p (0 = none) and delta */
int glpb_k_chuzr_dual(int m, int n, const signed char *type, const double *lb,
                      const double *ub, const int *head, const double *bbar,
                      const double *gamma, double tol_bnd, int *p,
                      double *delta);
/* chuzr, lib/glpspx01.js:808-1028, on the sorted list tcol_ind[1..tcol_num];
 * p = 0 none, -1 bound flip */
int glpb_k_ratio_primal(int m, int n, const signed char *type, const double *lb,
                        const double *ub, const double *coef, const int *head,
                        int phase, const double *bbar, double cbar_q, int q,
                        const int *tcol_ind, const double *tcol_vec,
                        int tcol_num, double rtol, int *p, int *p_stat,
                        double *teta);
/* chuzc, lib/glpspx02.js:793-935, on the sorted list trow_ind[1..trow_num] */
int glpb_k_ratio_dual(int n, const signed char *stat, const double *cbar,
                      double delta, const int *trow_ind, const double *trow_vec,
                      int trow_num, double rtol, int *q, double *new_dq);
/* sort_tcol, lib/glpspx01.js:773-806 / sort_trow, lib/glpspx02.js:754-791: the
 * entries of vec[1..n] with |v| >= eps, in the order the reference's swap loop
 * leaves them (the order its ratio tests examine, which settles exact ties);
 * out list[1..*num].  The solver's own per-kernel path runs the same kernel
 * before every ratio test; the persistent engines stop at an exact tie and hand
 * the iteration to that path (GLPB_TIES=0: lowest index instead). */
int glpb_k_sort_list(int n, const double *vec, double eps, int *list, int *num);
/* eval_trow1, lib/glpspx02.js:655-693 (pivot row as column dots).  A in CSA
 * layout; rho[1..m]; out trow_vec[1..n] */
int glpb_k_trow(int m, int n, const int *A_ptr, const int *A_ind,
                const double *A_val, const int *head, const signed char *stat,
                const double *rho, double *trow_vec);

/* Stand-alone streaming benchmark of the pricing kernel over resident data
 * (bench.py roofline leg): runs `reps` launches on n columns, returns the
 * average device time per launch in microseconds. */
int glpb_bench_kernel(const char *name, int m, int n, int reps, double *usec,
                      double *bytes);

/* Synthetic problem generators (RNG = lib/glprng01.js restated).
 * Each allocates with malloc; free with glpb_free_problem. */
typedef struct glpb_problem_data {
    int m, n, nnz, dir;
    double c0;
    int *type; double *lb, *ub;     /* [m+n] */
    double *coef; int *kind;        /* [n]   */
    int *A_ptr, *A_ind; double *A_val;
} glpb_problem_data;
int glpb_gen_packing(int m, int n, double density, int seed, glpb_problem_data *out);   /* C2 */
int glpb_gen_covering(int m, int n, int kmin, int kspan, int seed, glpb_problem_data *out); /* C3 */
int glpb_gen_mkp(int m, int n, int seed, glpb_problem_data *out);                       /* C5 */
void glpb_free_problem(glpb_problem_data *d);
void glpb_rng_fill(int seed, int count, int *out);

/* ---- Host-side preparation of the path's inputs --------
 * No device involved; O(nnz) work done once per solve on the caller thread.
 *
 * glpb_scale_prob replaces glp_scale_prob(lp, flags) (lib/glpscl.js:216-225,
 * scale_prob :167-214): cancels the current scaling, then geometric-mean
 * iterations (15 x, tau 0.90), equilibration and rounding to powers of two as
 * `flags` (GLP_SF_* of lib/glpk.js:30-34) ask.  A in CSC (0-based rows, any
 * order within a column); rii[m], sjj[n] are OUTPUT and are exactly what
 * glpb_create takes.  report (optional, 13 doubles): [0] = bit mask of the
 * stages reported (1 A, 2 GM, 4 EQ, 8 2N, +16 "well scaled, skipped"), then
 * (min|aij|, max|aij|, ratio) per stage -- the numbers the reference prints. */
#define GLPB_SF_GM   0x01
#define GLPB_SF_EQ   0x10
#define GLPB_SF_2N   0x20
#define GLPB_SF_SKIP 0x40
#define GLPB_SF_AUTO 0x80
int glpb_scale_prob(int m, int n, const int *A_ptr, const int *A_ind, const double *A_val,
                    int flags, double *rii, double *sjj, double *report);

/* glpb_adv_basis replaces glp_adv_basis(lp, 0) (lib/glpini01.js:281-363): the
 * maximal triangular part of (I | -A) without fixed columns becomes the
 * initial basis.  Columns (A_ptr/A_ind) and rows (R_ptr/R_ind, 0-based column
 * indices) must be given in the reference's LIST order (the order
 * glp_get_mat_col / glp_get_mat_row return): ties between rows are broken by
 * the order in which the column patterns are walked.  type/lb/ub [m+n] as for
 * glpb_create; stat[m+n] is OUTPUT (GLP_BS/NL/NU/NF/NS) and is what
 * glpb_set_basis takes; tri_size (optional) = size of the triangular part. */
int glpb_adv_basis(int m, int n, const int *A_ptr, const int *A_ind, const int *R_ptr,
                   const int *R_ind, const int *type, const double *lb, const double *ub,
                   int *stat, int *tri_size);

/* glpb_read_lp replaces glp_read_lp (lib/glpcpx.js:10-753) for a text held in
 * memory: CPLEX LP format -> the arrays glpb_create takes (columns with ascending
 * row indices, i.e. the state after the reference's final glp_sort_matrix; rows
 * and columns numbered in order of first appearance; c0 = 0).  names (optional)
 * receives one malloc'ed block of NUL-terminated strings: the objective name, the
 * m row names ("r.<i>" where the text gives none), the n column names; release it
 * with glpb_free_names and the arrays with glpb_free_problem.  Returns 0, 1 on a
 * syntax error (the reference's return value; text in glpb_last_error, with the
 * line number), GLPB_EINVAL / GLPB_ENOMEM. */
int glpb_read_lp(const char *text, long len, glpb_problem_data *out, char **names, long *names_len);
void glpb_free_names(char *names);

/* glpb_write_lp replaces glp_write_lp (lib/glpcpx.js:755-999): the problem as CPLEX
 * LP text, identical line for line to what the reference hands to its callback
 * (numbers as JavaScript's string concatenation shows them, lines broken before
 * column 73, r_<i> / x_<j> / obj for missing or invalid names -- the reference's
 * adjust_name assigns into an immutable string and changes nothing).  type/lb/ub
 * [m+n], coef/kind [n] as for glpb_create (unscaled); col_len[n] = number of
 * elements per column (an empty column is written into the objective with a zero
 * coefficient); rows R_ptr/R_ind/R_val (0-based column indices) in the reference's
 * LIST order, which is the order of the terms of every constraint.  prob_name and
 * names are optional; names = one block of 1+m+n NUL-terminated strings
 * (objective, rows, columns -- the layout glpb_read_lp returns), "" = no name.
 * *text receives a malloc'ed NUL-terminated buffer, lines separated by '\n'
 * (release it with glpb_free_names); *lines = the count the reference reports in
 * "<count> lines were written".  Returns 0, GLPB_EINVAL or GLPB_ENOMEM. */
int glpb_write_lp(int m, int n, int dir, double c0, const int *type, const double *lb,
                  const double *ub, const double *coef, const int *kind, const int *col_len,
                  const int *R_ptr, const int *R_ind, const double *R_val, const char *prob_name,
                  const char *names, char **text, long *text_len, int *lines);

/* ---- LP / MIP presolver (csrc/presolve.cpp) --------------
 * Host only, no device involved.  One workspace per `presolve: GLP_ON` solve,
 * used in the order of lib/glpapi06.js:41-146 (LP) / lib/glpapi09.js:116-256
 * (MIP):
 *
 *   glpb_npp_create                      npp_create_wksp   lib/glpnpp01.js:2-22
 *   glpb_npp_load_prob(.., sol)          npp_load_prob(npp, P, GLP_OFF, sol, GLP_OFF)   :262-394
 *   glpb_npp_simplex | glpb_npp_integer  npp_simplex | npp_integer   lib/glpnpp05.js:430-521
 *   glpb_npp_get_size, glpb_npp_build_prob   npp_build_prob          lib/glpnpp01.js:396-472
 *        ... scale, crash basis, glpb_create / glpb_simplex / glpb_intopt on the REDUCED problem ...
 *   glpb_npp_postprocess                 npp_postprocess             lib/glpnpp01.js:474-570
 *   glpb_npp_destroy
 *
 * load_prob: type/lb/ub [m+n] (rows first), coef/kind [n] and the matrix by
 * columns as for glpb_create, but UNSCALED and with the elements of every column
 * in the reference's LIST order (the order glp_get_mat_col returns): the order
 * of the reduced problem's rows, columns and elements follows from it.  sol =
 * GLP_SOL (1) or GLP_MIP (3); kind is read for GLP_MIP only.
 *
 * simplex / integer return 0, GLP_ENOPFS (0x0A) or GLP_ENODFS (0x0B) like the
 * reference; `binarize` = iocp.binarize.  get_counts (optional): what
 * npp_integer prints -- [0] hidden packing, [1] hidden covering inequalities,
 * [2] reduced coefficients, [3..6] binarization (variables replaced, binaries
 * created, rows added, failures), [7] entries on the recovery stack, [8..20] of
 * those per transformation (free row, fixed column, make equality, make fixed,
 * empty column, equality singleton, inequality singleton, implied slack, implied
 * free, forcing row, inactive bound, shifted lower bound, binarized column).
 *
 * build_prob emits the reduced problem (sizes from get_size) in the layout
 * glpb_create takes: c0 and coef in the ORIGINAL objective sense, bounds with
 * -/+DBL_MAX for absent ones plus the GLP_* type npp_build_prob derives, the
 * elements of every column in the order the reference hands them to
 * glp_set_mat_col (a binding that mirrors the reference's list state prepends
 * them one by one), row_ref/col_ref (optional) = reference numbers of the
 * reduced rows / columns in the workspace.  The workspace keeps only the
 * recovery stack afterwards.
 *
 * postprocess takes the solution of the reduced problem -- basic: r_stat,
 * r_dual [m'], c_stat, c_value (= prim) [n']; MIP: c_value (= mipx) only, the
 * other inputs and outputs may be NULL -- runs the recovery stack and returns
 * what npp_unload_sol then copies into the original problem (lib/glpnpp01.js:
 * 596-660): row statuses and duals [orig m] (duals already in the original
 * objective sense), column statuses and values [orig n].  Primal values of
 * non-basic variables, row activities and reduced costs are recomputed by the
 * binding from ITS coefficients exactly as npp_unload_sol does.  Returns 0,
 * GLPB_EINVAL, or GLPB_ESTATE where the reference's xassert(tse.func() == 0)
 * would throw.  Every call returns GLPB_ENOMEM instead of letting an allocation
 * failure cross the C ABI. */
typedef struct glpb_npp glpb_npp;
glpb_npp *glpb_npp_create(void);
void glpb_npp_destroy(glpb_npp *npp);                /* idempotent on NULL */
int glpb_npp_load_prob(glpb_npp *npp, int m, int n, int dir, double c0, const int *type,
                       const double *lb, const double *ub, const double *coef, const int *kind,
                       const int *A_ptr, const int *A_ind, const double *A_val, int sol);
int glpb_npp_simplex(glpb_npp *npp);
int glpb_npp_integer(glpb_npp *npp, int binarize);
int glpb_npp_get_counts(glpb_npp *npp, int *out, int count);
int glpb_npp_get_size(glpb_npp *npp, int *m, int *n, int *nnz);
int glpb_npp_build_prob(glpb_npp *npp, double *c0, int *type, double *lb, double *ub, double *coef,
                        int *kind, int *A_ptr, int *A_ind, double *A_val, int *row_ref, int *col_ref);
int glpb_npp_postprocess(glpb_npp *npp, const int *r_stat, const double *r_dual, const int *c_stat,
                         const double *c_value, int *out_r_stat, double *out_r_dual,
                         int *out_c_stat, double *out_c_value);

#ifdef __cplusplus
}
#endif
#endif /* GLPB200_H */
