/* glpo.h -- CPU ORACLE (test infrastructure, NOT the product path).
 *
 * A single-threaded C++ restatement of the simplex hot path of glpk.js
 * (a JavaScript port of GLPK 4.49).  It exists only so that tests/, the
 * smoke check and bench.py's cpu_baseline / --impl reference legs have
 * something to compare the CUDA path against.  Nothing under glpk.js_b200/
 * links, loads or calls this code.
 *
 * PARITY STATUS: "parity unpinned" by the reference itself -- the reference
 * ships no asserted results and cannot be executed in this image (no JS
 * engine).  The oracle is pinned instead by (i) independent HiGHS optima of
 * the reference's fixtures, (ii) the hand trace of test.lpt,
 * (iii) the reference's own err_in_bbar/cbar/gamma invariants.
 *
 * All arrays are 1-based like the reference (slot 0 unused or a scalar).
 * Every function cites the reference file:line it follows.
 */
#ifndef GLPO_H
#define GLPO_H

#include <cfloat>
#include <climits>
#include <cstdint>
#include <string>
#include <vector>

namespace glpo {

/* ---- constants: lib/glpk.js ---- */
enum { GLP_MIN = 1, GLP_MAX = 2 };
enum { GLP_CV = 1, GLP_IV = 2, GLP_BV = 3 };
enum { GLP_FR = 1, GLP_LO = 2, GLP_UP = 3, GLP_DB = 4, GLP_FX = 5 };
enum { GLP_BS = 1, GLP_NL = 2, GLP_NU = 3, GLP_NF = 4, GLP_NS = 5 };
enum { GLP_UNDEF = 1, GLP_FEAS = 2, GLP_INFEAS = 3, GLP_NOFEAS = 4,
       GLP_OPT = 5, GLP_UNBND = 6 };
enum { GLP_MSG_OFF = 0, GLP_MSG_ERR = 1, GLP_MSG_ON = 2, GLP_MSG_ALL = 3,
       GLP_MSG_DBG = 4 };
enum { GLP_PRIMAL = 1, GLP_DUALP = 2, GLP_DUAL = 3 };
enum { GLP_PT_STD = 0x11, GLP_PT_PSE = 0x22 };
enum { GLP_RT_STD = 0x11, GLP_RT_HAR = 0x22 };
enum { GLP_BR_FFV = 1, GLP_BR_LFV = 2, GLP_BR_MFV = 3, GLP_BR_DTH = 4,
       GLP_BR_PCH = 5 };
enum { GLP_BT_DFS = 1, GLP_BT_BFS = 2, GLP_BT_BLB = 3, GLP_BT_BPH = 4 };
enum { GLP_PP_NONE = 0, GLP_PP_ROOT = 1, GLP_PP_ALL = 2 };
enum { GLP_EBADB = 0x01, GLP_ESING = 0x02, GLP_ECOND = 0x03,
       GLP_EBOUND = 0x04, GLP_EFAIL = 0x05, GLP_EOBJLL = 0x06,
       GLP_EOBJUL = 0x07, GLP_EITLIM = 0x08, GLP_ETMLIM = 0x09,
       GLP_ENOPFS = 0x0A, GLP_ENODFS = 0x0B, GLP_EROOT = 0x0C,
       GLP_ESTOP = 0x0D, GLP_EMIPGAP = 0x0E };
enum { BFD_ESING = 1, BFD_ECOND = 2, BFD_ECHECK = 3, BFD_ELIMIT = 4,
       BFD_EROOM = 5 };

/* ---- sparse LU  (lib/glpluf.js) ---- */
struct LUF {
    int n = 0, valid = 0;
    std::vector<int> fr_ptr, fr_len, fc_ptr, fc_len;
    std::vector<int> vr_ptr, vr_len, vr_cap;
    std::vector<double> vr_piv;
    std::vector<int> vc_ptr, vc_len, vc_cap;
    std::vector<int> pp_row, pp_col, qq_row, qq_col;
    int sv_size = 0, sv_beg = 0, sv_end = 0;
    std::vector<int> sv_ind;
    std::vector<double> sv_val;
    int sv_head = 0, sv_tail = 0;
    std::vector<int> sv_prev, sv_next;
    std::vector<double> vr_max;
    std::vector<int> rs_head, rs_prev, rs_next, cs_head, cs_prev, cs_next;
    std::vector<int> flag;
    std::vector<double> work;
    int new_sva = 0;
    double piv_tol = 0.10;
    int piv_lim = 4, suhl = 1;
    double eps_tol = 1e-15, max_gro = 1e+10;
    int nnz_a = 0, nnz_f = 0, nnz_v = 0;
    double max_a = 0, big_v = 0;
    int rank = 0;
};

/* column callback: fills ind[1..len], val[1..len]; returns len */
typedef int (*col_fn)(void *info, int j, int *ind, double *val);

int luf_factorize(LUF &luf, int n, col_fn col, void *info);
void luf_f_solve(LUF &luf, int tr, double *x, const int *pp_row);
void luf_v_solve(LUF &luf, int tr, double *x);

/* ---- Forrest-Tomlin  (lib/glpfhv.js) + driver (lib/glpbfd.js) ---- */
struct BFD {
    int valid = 0;
    int m = 0;
    LUF luf;
    int hh_max = 100, hh_nfs = 0;
    std::vector<int> hh_ind, hh_ptr, hh_len;
    std::vector<int> p0_row, p0_col;
    std::vector<int> cc_ind;
    std::vector<double> cc_val;
    double upd_tol = 1e-6;
    int nnz_h = 0;
    int upd_cnt = 0;
    /* bfcp (lib/glpbfd.js:10-29) */
    int lu_size = 0;
    double piv_tol = 0.10;
    int piv_lim = 4, suhl = 1;
    double eps_tol = 1e-15, max_gro = 1e+10;
    int nfs_max = 100;
    bool fresh = true;
    /* statistics for the CPU baseline */
    long n_factorize = 0, n_update = 0, n_ftran = 0, n_btran = 0;
};

int bfd_factorize(BFD &bfd, int m, col_fn col, void *info);
void bfd_ftran(BFD &bfd, double *x);
void bfd_btran(BFD &bfd, double *x);
int bfd_update_it(BFD &bfd, int j, int len, const int *ind, int idx,
                  const double *val);

/* ---- problem object (the subset of lib/glpapi01.js the path reads) ---- */
struct Elem { int idx; double val; };

struct Prob {
    int m = 0, n = 0, nnz = 0;
    int dir = GLP_MIN;
    double c0 = 0.0;
    /* rows 1..m */
    std::vector<int> r_type, r_stat, r_bind;
    std::vector<double> r_lb, r_ub, r_rii, r_prim, r_dual, r_mipx;
    std::vector<std::string> r_name;
    /* cols 1..n */
    std::vector<int> c_type, c_stat, c_bind, c_kind;
    std::vector<double> c_lb, c_ub, c_coef, c_sjj, c_prim, c_dual, c_mipx;
    std::vector<std::string> c_name;
    /* matrix in *list order*  */
    std::vector<std::vector<Elem>> col_list; /* col j: (row i, val) */
    std::vector<std::vector<Elem>> row_list; /* row i: (col j, val) */
    /* basis */
    int valid = 0;
    std::vector<int> head;
    BFD *bfd = nullptr;
    /* solution */
    int pbs_stat = GLP_UNDEF, dbs_stat = GLP_UNDEF;
    double obj_val = 0.0;
    int it_cnt = 0, some = 0;
    int mip_stat = GLP_UNDEF;
    double mip_obj = 0.0;
    std::string name, obj_name;
    ~Prob();
};

struct SMCP { /* lib/glpapi06.js:359-375 */
    int msg_lev = GLP_MSG_ALL;
    int meth = GLP_PRIMAL;
    int pricing = GLP_PT_PSE;
    int r_test = GLP_RT_HAR;
    double tol_bnd = 1e-7, tol_dj = 1e-7, tol_piv = 1e-10;
    double obj_ll = -DBL_MAX, obj_ul = +DBL_MAX;
    int it_lim = INT_MAX, tm_lim = INT_MAX;
    int out_frq = 500, out_dly = 0;
    int presolve = 0;
};

struct IOCP { /* lib/glpapi09.js:392-414 */
    int msg_lev = GLP_MSG_ALL;
    int br_tech = GLP_BR_DTH;
    int bt_tech = GLP_BT_BLB;
    double tol_int = 1e-5, tol_obj = 1e-7;
    int tm_lim = INT_MAX;
    int out_frq = 5000, out_dly = 10000;
    int pp_tech = GLP_PP_ALL;
    double mip_gap = 0.0;
    int presolve = 0;
    long node_lim = -1; /* extension: stop after this many solved nodes */
};

/* hook for kernel-level parity tests: called at fixed points of the loops
   with a pointer to the live CSA (see primal.cpp / dual.cpp) */
typedef void (*hook_fn)(void *user, int event, void *csa);
enum { EV_P_CHUZC = 1, EV_P_CHUZR = 2, EV_P_TROW = 3, EV_P_GAMMA = 4,
       EV_P_ITER = 5,
       EV_D_CHUZR = 11, EV_D_CHUZC = 12, EV_D_TROW = 13, EV_D_GAMMA = 14,
       EV_D_ITER = 15 };
struct Hook { hook_fn fn = nullptr; void *user = nullptr; };

void prob_add_rows(Prob &P, int nrs);
void prob_add_cols(Prob &P, int ncs);
void prob_set_row_bnds(Prob &P, int i, int type, double lb, double ub);
void prob_set_col_bnds(Prob &P, int j, int type, double lb, double ub);
void prob_set_mat_row(Prob &P, int i, int len, const int *ind,
                      const double *val);
void prob_sort_matrix(Prob &P);
void prob_set_row_stat(Prob &P, int i, int stat);
void prob_set_col_stat(Prob &P, int j, int stat);
void prob_std_basis(Prob &P);
int prob_factorize(Prob &P);
int prob_get_status(const Prob &P);
int read_lp(Prob &P, const std::string &text, std::string &err);
std::string write_lp(const Prob &P);

int spx_primal(Prob &lp, const SMCP &parm, const Hook *hook);
int spx_dual(Prob &lp, const SMCP &parm, const Hook *hook);
int simplex(Prob &P, const SMCP &parm, const Hook *hook);
int intopt(Prob &P, const IOCP &parm, long *n_nodes);

/* Knuth subtractive RNG (lib/glprng01.js, glprng02.js) */
struct RNG { int A[56]; int fptr; };
void rng_init(RNG &r, int seed);
int rng_next(RNG &r);
int rng_unif_rand(RNG &r, int m);
double rng_unif_01(RNG &r);

double xtime_ms();

} /* namespace glpo */
#endif
