/* ne_emul.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the batched
 * branch-and-bound of the product (csrc/nodeengine.cuh + csrc/bnbpool.cuh)
 * for the HOST with one "thread" per CTA (-DNE_EMUL), so that the node logic
 * and the tree / migration code can be exercised by `-m "not gpu"` tests and
 * by the gloo world-size-2 tests.  The product never loads this library; on
 * the device the same source runs as k_bnb_nodes.
 */
#define NE_EMUL 1
#include "../../include/glpb200.h"
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <vector>

enum { GLP_MIN = 1, GLP_MAX = 2 };
enum { GLP_CV = 1, GLP_IV = 2 };
enum { GLP_BT_DFS = 1, GLP_BT_BFS = 2, GLP_BT_BLB = 3, GLP_BT_BPH = 4 };
enum { GLP_EFAIL = 5, GLP_ETMLIM = 9, GLP_ESTOP = 13, GLP_EMIPGAP = 14 };

struct glpb_prob {
    int m = 0, n = 0, dir = GLP_MIN;
    double c0 = 0.0;
    std::vector<int> h_type, h_kind, h_stat, h_aptr, h_aind;
    std::vector<double> h_lb, h_ub, h_coef, h_rii, h_sjj, h_aval;
};

static char g_err[512];
void glpb_set_error(const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}
static double glpb_now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

#include "../../glpk.js_b200/csrc/bnbpool.cuh"

struct Emul { glpb_prob P; glpb_bnb *T = nullptr; };

#define API extern "C" __attribute__((visibility("default")))

API void *ne_emul_create(int m, int n, int dir, double c0, const int *type, const double *lb, const double *ub,
                         const int *stat, const double *coef, const int *kind, const double *rii, const double *sjj,
                         const int *aptr, const int *aind, const double *aval,
                         int br_tech, int bt_tech, int pp_tech, double tol_int, double tol_obj, double mip_gap,
                         long node_lim, int batch, int cap)
{
    Emul *E = new Emul();
    glpb_prob &P = E->P;
    P.m = m; P.n = n; P.dir = dir; P.c0 = c0;
    P.h_type.assign(type, type + m + n); P.h_lb.assign(lb, lb + m + n); P.h_ub.assign(ub, ub + m + n);
    P.h_stat.assign(stat, stat + m + n); P.h_coef.assign(coef, coef + n); P.h_kind.assign(kind, kind + n);
    P.h_rii.assign(m, 1.0); P.h_sjj.assign(n, 1.0);
    if (rii) P.h_rii.assign(rii, rii + m);
    if (sjj) P.h_sjj.assign(sjj, sjj + n);
    P.h_aptr.assign(aptr, aptr + n + 1); P.h_aind.assign(aind, aind + aptr[n]); P.h_aval.assign(aval, aval + aptr[n]);
    glpb_iocp parm;
    memset(&parm, 0, sizeof parm);
    parm.br_tech = br_tech; parm.bt_tech = bt_tech; parm.pp_tech = pp_tech; parm.tol_int = tol_int; parm.tol_obj = tol_obj;
    parm.mip_gap = mip_gap; parm.node_lim = node_lim; parm.tm_lim = INT_MAX;
    E->T = new glpb_bnb();
    E->T->tm_beg = glpb_now_ms();
    if (E->T->init(&E->P, parm, batch, cap) != 0) { delete E->T; delete E; return nullptr; }
    return E;
}

API void ne_emul_destroy(void *h) { Emul *E = (Emul *)h; if (E) { delete E->T; delete E; } }
API int ne_emul_round(void *h, long max_tasks, long *done) { return ((Emul *)h)->T->round(max_tasks, done); }
API int ne_emul_open(void *h) { return ((Emul *)h)->T->open_count(); }
API void ne_emul_incumbent(void *h, int *have_sol, int *have_cut, double *obj, double *x)
{
    glpb_bnb *T = ((Emul *)h)->T;
    *have_sol = T->have_sol; *have_cut = T->have_cut; *obj = T->mip_obj;
    if (x) memcpy(x, T->mipx.data(), T->mn * sizeof(double));
}
API void ne_emul_set_cutoff(void *h, double obj) { ((Emul *)h)->T->set_cutoff(obj); }
API void ne_emul_clear(void *h) { ((Emul *)h)->T->clear(); }
API long ne_emul_record_bytes(void *h) { return (long)((Emul *)h)->T->record_bytes(); }
API int ne_emul_export(void *h, int max_count, void *buf, int *count) { return ((Emul *)h)->T->export_nodes(max_count, (unsigned char *)buf, count); }
API int ne_emul_import(void *h, const void *buf, int count) { return ((Emul *)h)->T->import_nodes((const unsigned char *)buf, count); }
API void ne_emul_stats(void *h, long *out5)
{
    glpb_bnb *T = ((Emul *)h)->T;
    out5[0] = T->solved; out5[1] = T->tasks_done; out5[2] = T->rounds; out5[3] = T->iters; out5[4] = T->refacs;
}
API const char *ne_emul_error(void) { return g_err; }
