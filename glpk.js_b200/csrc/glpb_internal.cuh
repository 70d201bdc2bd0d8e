/* glpb_internal.cuh -- shared declarations of libglpb200 (sm_100a).
 *
 * Device data layout (all 0-based; "k" numbers variables 0..m+n-1, auxiliary
 * rows first, exactly the reference's 1..m+n shifted by one):
 *
 *   A   CSC  a_ptr[n+1] a_ind[nnz] a_val[nnz]   scaled values rii*a*sjj
 *   A'  CSR  at_ptr[m+1] at_ind[nnz] at_val[nnz]
 *   type[m+n] i8, lb/ub/coef[m+n] f64 (scaled; coef = working costs)
 *   head[m+n] i32  (first m basic, then n non-basic), bind[m+n] = head^-1
 *   stat[n] i8, bbar[m], cbar[n], gamma[max(m,n)], refsp[m+n] i8
 *
 * Basis inverse ("structural kernel" form).  With P_S the basic positions
 * that hold structural columns and R_N the rows whose auxiliary variable is
 * non-basic (|P_S| = |R_N| = k), the only non-trivial block of inv(B) is
 *      T = inv(B)[P_S, R_N] = -inv(A[R_N, J_B])          (k x k, dense)
 * kept column-major in HBM with leading dimension ldt; rslot/slot_pos map
 * positions <-> rows of T, cslot/slot_row map constraint rows <-> columns.
 * FTRAN / BTRAN are one GEMV with T plus one gather pass over A' / A; a basis
 * change is a rank-1 update of T plus O(k) bookkeeping.
 */
#ifndef GLPB_INTERNAL_CUH
#define GLPB_INTERNAL_CUH

#include <cuda_runtime.h>
#include <cfloat>
#include <climits>
#include <cstdint>
#include <map>
#include <string>
#include <vector>
#include "../../include/glpb200.h"

/* GLP_* constants (lib/glpk.js) */
enum { GLP_MIN = 1, GLP_MAX = 2 };
enum { GLP_CV = 1, GLP_IV = 2 };
enum { GLP_FR = 1, GLP_LO = 2, GLP_UP = 3, GLP_DB = 4, GLP_FX = 5 };
enum { GLP_BS = 1, GLP_NL = 2, GLP_NU = 3, GLP_NF = 4, GLP_NS = 5 };
enum { GLP_UNDEF = 1, GLP_FEAS = 2, GLP_INFEAS = 3, GLP_NOFEAS = 4, GLP_OPT = 5, GLP_UNBND = 6 };
enum { GLP_PRIMAL = 1, GLP_DUALP = 2, GLP_DUAL = 3 };
enum { GLP_PT_STD = 0x11, GLP_PT_PSE = 0x22 };
enum { GLP_RT_STD = 0x11, GLP_RT_HAR = 0x22 };
enum { GLP_BR_FFV = 1, GLP_BR_LFV = 2, GLP_BR_MFV = 3, GLP_BR_DTH = 4, GLP_BR_PCH = 5 };
enum { GLP_BT_DFS = 1, GLP_BT_BFS = 2, GLP_BT_BLB = 3, GLP_BT_BPH = 4 };
enum { GLP_PP_NONE = 0, GLP_PP_ROOT = 1, GLP_PP_ALL = 2 };
enum { GLP_EBADB = 1, GLP_ESING = 2, GLP_ECOND = 3, GLP_EBOUND = 4, GLP_EFAIL = 5,
       GLP_EOBJLL = 6, GLP_EOBJUL = 7, GLP_EITLIM = 8, GLP_ETMLIM = 9, GLP_ENOPFS = 10,
       GLP_ENODFS = 11, GLP_EROOT = 12, GLP_ESTOP = 13, GLP_EMIPGAP = 14 };

#define GLPB_KAPPA 0.10   /* lib/glpspx01.js:3 */

/* iteration status written by the device, read by the host loop controller */
enum {
    ST_OK = 0,
    ST_NONE1 = 1,      /* pricing found nothing (primal chuzc q / dual chuzr p) */
    ST_D1D2 = 2,       /* primal: reduced cost of xN[q] inaccurate             */
    ST_NONE2 = 3,      /* ratio test found nothing                             */
    ST_PIVSMALL = 4,   /* pivot below 1e-5(1+0.01 max)                         */
    ST_PIV12 = 5,      /* tcol[p] and trow[q] disagree                         */
    ST_SINGULAR = 6,   /* refactorisation met a zero pivot                     */
    /* raised at the end of a completed iteration (batch mode) */
    ST_REFAC = 7,      /* refactorisation period reached                       */
    ST_LIMIT = 8,      /* iteration limit reached                              */
    ST_REFSP = 9,      /* reference space must be reset (refct == 0)           */
    ST_OBJLIM = 10,    /* dual: objective crossed obj_ll / obj_ul              */
    ST_PHASE = 12,     /* phase 1: the iteration just completed left nothing infeasible; the host
                          repeats the reference's check_feas and switches to phase 2             */
    /* raised inside an iteration by the engines only */
    ST_TIE = 11,       /* exact tie in a ratio test: the host repeats the iteration with the
                          reference's sort_tcol / sort_trow list order (k_sort_list)          */
};

#define P_NONE (-1)
#define P_FLIP (-2)

/* device control block; one per handle */
struct Ctrl {
    int status;
    int q, p, p_stat;
    int phase;
    int k;              /* current size of T                                  */
    int flag;           /* generic boolean result of check_* kernels          */
    int cnt;            /* generic counter result (set_aux_obj)               */
    int skip2;          /* ratio test: second pass not needed                 */
    int sing;           /* refactorisation: singular                          */
    /* loop state kept on the device so that several iterations can be enqueued
       back to back; the host seeds it at the start of a batch */
    int it_cnt, it_max, refct, upd_cnt, period, n_done;
    int rigorous, bbar_fresh, cbar_fresh, binv_fresh, pse;
    int list_num;       /* entries of the ratio-test list k_sort_list left          */
    double obj_ll, obj_ul, zeta;
    double teta, delta, new_dq, tmax;
    double tcol_max, trow_max, eps;
    double piv1, piv2, d1, d2;
    double gamma_q, delta_q;
    double obj;         /* dual: tracked objective (bbar[0] in the reference)  */
    double big, scal;   /* scratch scalars                                    */
    double max_a;       /* largest |a| seen when building the kernel matrix   */
    /* optional pivot log (parity tests): entry it = (q, p) of iteration it, 0-based
       positions (p = P_FLIP for a bound flip); written by iter_end on every path */
    int *piv_log;
    int piv_cap, piv_pad;
    unsigned int ticket[12];
};

/* key used by every arg-reduction */
struct Key {
    double a, b, c;
    int pos, aux;
};

struct glpb_mip;
void glpb_mip_free(glpb_mip *T);
struct glpb_bnb;
void glpb_bnb_free(glpb_bnb *T);

struct glpb_prob {
    int device = 0;
    cudaStream_t stream = nullptr;
    int m = 0, n = 0, nnz = 0, dir = GLP_MIN;
    double c0 = 0.0;
    int ldt = 0;               /* leading dimension / capacity of T */
    /* ---- host copies of the unscaled problem (the glp_prob fields) ---- */
    std::vector<int> h_type, h_kind, h_stat;     /* [m+n] / [n] / [m+n] GLP_BS.. */
    std::vector<double> h_lb, h_ub, h_coef, h_rii, h_sjj;
    std::vector<int> h_aptr, h_aind;             /* host CSC (for B&B preprocessing) */
    std::vector<double> h_aval;
    std::vector<int> h_atptr, h_atind;           /* host CSR */
    std::vector<double> h_atval;
    std::vector<int> h_head;                     /* [m], values k+1 */
    int valid = 0;
    int t_ok = 0;                /* the device's T / slot maps are consistent with h_head */
    int pbs_stat = GLP_UNDEF, dbs_stat = GLP_UNDEF;
    double obj_val = 0.0;
    int it_cnt = 0, some = 0;
    std::vector<double> h_prim, h_dual;          /* [m+n] */
    int mip_stat = GLP_UNDEF;
    double mip_obj = 0.0;
    std::vector<double> h_mipx;
    long mip_nodes = 0;
    glpb_bfcp bfcp = {100, 0.10, 1e-6};
    /* ---- device arrays ---- */
    int *a_ptr = nullptr, *a_ind = nullptr, *at_ptr = nullptr, *at_ind = nullptr;
    double *a_val = nullptr, *at_val = nullptr;
    signed char *type = nullptr, *orig_type = nullptr, *stat = nullptr, *refsp = nullptr;
    double *lb = nullptr, *ub = nullptr, *coef = nullptr, *orig_lb = nullptr, *orig_ub = nullptr;
    double *obj = nullptr;            /* [n+1]: obj[0] = c0, obj[1+j] scaled original costs */
    int *head = nullptr, *bind = nullptr;
    double *bbar = nullptr, *cbar = nullptr, *gamma = nullptr;
    double *tcol = nullptr, *trow = nullptr, *rho = nullptr;
    double *svec = nullptr;           /* [n] PSE inner products s_j (primal) */
    double *w1 = nullptr, *w2 = nullptr, *w3 = nullptr, *w4 = nullptr, *w5 = nullptr; /* [m] work */
    double *yk = nullptr, *wk = nullptr;   /* [ldt] work in kernel space */
    double *yk2 = nullptr, *zn = nullptr;  /* [ldt] engine work in kernel space */
    struct EngSlot *eng_slots = nullptr;   /* engine: barrier flags + partials [ENG_RING][ENG_MAXG] */
    double *eng_cols = nullptr;            /* engine: ycol, ycol2, trowcol [3n] and vrow [m] */
    int *sort_list = nullptr, *sort_tmp = nullptr;   /* ratio-test list in the reference's order [max(m,n)], scratch [2 max(m,n) + 2] */
    long n_tie = 0;                        /* iterations the engines handed over because of an exact tie */
    double *eng_fr = nullptr;              /* engine: deferred terms Fd, Rd [2][ENG_DB][ldt] and zbuf [ENG_DB] */
    long long *eng_cyc = nullptr;          /* engine: SM cycles per phase [16] */
    double *eng_bytes = nullptr;           /* engine: algorithmic bytes per phase [16] */
    int sm_count = 0, eng_dcap = 0, eng_smem = 0, eng_ready = 0, eng_hdr = 0, eng_tma = 0;
    long n_eng_launch = 0, n_eng_prof_iter[2] = {0, 0};
    double *T = nullptr, *T2 = nullptr, *partial = nullptr;   /* T2: output buffer of the refactorisation */
    struct RefSlot *ref_slots = nullptr;
    unsigned int *ref_flags = nullptr;
    double *ref_xp = nullptr;              /* [REF_NB][ldt] pivot rows of a refactorisation round */
    int partial_rows = 0;
    int *rslot = nullptr, *slot_pos = nullptr, *cslot = nullptr, *slot_row = nullptr;
    int *gj_piv = nullptr;
    double *gj_row = nullptr;        /* scratch of the refactorisation (gather list) */
    Key *scratch = nullptr;           /* block partials of reductions */
    Ctrl *ctrl = nullptr;             /* device */
    Ctrl *h_ctrl = nullptr;           /* pinned host mirror */
    unsigned char *h_stage = nullptr; /* pinned staging for the per-solve uploads / read-backs */
    size_t stage_bytes = 0;
    int k_host = 0;                   /* kernel size as of the last read-back of ctrl */
    double zeta = 1.0;
    /* ---- counters ---- */
    long n_iter = 0, n_refac = 0, n_launch = 0, n_sync = 0, n_update = 0;
    double last_solve_us = 0.0;
    int trace = 0;
    /* ---- optional per-kernel timing with CUDA events on the solve stream ---- */
    struct ProfRec { const char *name; cudaEvent_t e0, e1; double bytes; };
    struct ProfAcc { double ms = 0.0, bytes = 0.0; long count = 0; };
    int prof = 0;
    double next_bytes = 0.0;        /* algorithmic bytes of the next launch */
    std::vector<ProfRec> prof_recs;
    std::map<std::string, ProfAcc> prof_acc;
    std::string prof_text;
    int *piv_log = nullptr;        /* device pivot log (glpb_set_pivot_log) */
    int piv_cap = 0;
    glpb_mip *mip = nullptr;       /* branch-and-bound tree while glp_intopt runs */
    glpb_bnb *bnb = nullptr;       /* batched branch-and-bound (bnbpool.cuh) while it runs */
    /* ---- replayed launch sequences (small LPs: the node LPs of branch-and-bound).
       The fresh recomputations of bbar / cbar and the start-of-solve block of the dual
       loop are fixed sequences of 7-22 tiny kernels whose arguments never change for
       a handle except the T/T2 flip; they are captured once into CUDA graphs and
       replayed with one call (host launch cost, not arithmetic, is the node time). */
    struct GraphSlot { cudaGraphExec_t exec = nullptr; const void *T = nullptr; double key = 0.0; int launches = 0; };
    enum { GK_BBAR = 0, GK_CBAR = 1, GK_DSTART = 2, GK_DFINAL = 3, GK_KINDS = 4 };
    GraphSlot graphs[GK_KINDS][2];
    int graph_ok = 0, capturing = 0;
    long n_graph = 0;
};

void glpb_set_error(const char *fmt, ...);

#define CK(call)                                                               \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) {                                              \
            glpb_set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call,         \
                           cudaGetErrorString(e__));                           \
            return GLPB_ENODEV;                                                \
        }                                                                      \
    } while (0)

#endif
