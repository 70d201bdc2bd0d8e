/* presolve.cpp -- the LP / MIP presolver on the input side of the simplex path
 * (VERDICT r1 "missing" item 3).
 *
 * Replaces, for `presolve: GLP_ON` (lib/glpapi06.js:41-146, lib/glpapi09.js:116-256):
 *     npp_load_prob      lib/glpnpp01.js:262-394   -> glpb_npp_load_prob
 *     npp_simplex        lib/glpnpp05.js:430-435   -> glpb_npp_simplex
 *     npp_integer        lib/glpnpp05.js:437-521   -> glpb_npp_integer
 *     npp_build_prob     lib/glpnpp01.js:396-472   -> glpb_npp_get_size + glpb_npp_build_prob
 *     npp_postprocess    lib/glpnpp01.js:474-570   -> glpb_npp_postprocess
 * and the transformations those drive (lib/glpnpp02.js, glpnpp03.js, glpnpp04.js).
 * npp_unload_sol's copy into the caller's problem object stays with the host
 * binding (it needs the caller's own rows / columns).
 *
 * Host code by design: symbolic list surgery, O(nnz), once per solve.  What the
 * DEVICE gets out of it is a smaller LP -- the reduced problem this file emits is
 * what glpb_create uploads.
 *
 * Parity contract: the reduced problem must be the reference's bit for bit AND in
 * the reference's order (row order, column order, element order inside every
 * column), because scaling, the crash basis and the simplex pivots downstream
 * depend on those orders.  The reference keeps rows, columns and elements in
 * doubly linked lists of heap objects and moves "active" rows / columns to the
 * list heads; here the same lists live in three flat arenas addressed by 32-bit
 * indices (0 = null), a row's / column's arena index IS its reference number, and
 * the recovery records are tagged structs with their coefficient lists in one
 * shared pool (walked backwards: the reference prepends).  Arithmetic is kept
 * expression by expression (compiled with -ffp-contract=off).
 */
#include "../../include/glpb200.h"
#include <cfloat>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

namespace npp_detail {

constexpr int BS = 1, NL = 2, NU = 3, NF = 4, NS = 5;      /* lib/glpk.js:23-27 */
constexpr int FR = 1, LO = 2, UP = 3, DB = 4, FX = 5;      /* lib/glpk.js:16-20 */
constexpr int SOL = 1, MIP = 3;                            /* lib/glpk.js:37-39 */
constexpr int ENOPFS = 0x0A, ENODFS = 0x0B;                /* lib/glpk.js:132-133 */
constexpr int IV = 2;
constexpr double INF = DBL_MAX;

struct Row {
    double lb, ub;
    int ptr, prev, next;
    int temp;              /* activity flag; ordinal in the reduced problem after build */
};
struct Col {
    double lb, ub, coef, ll, uu;
    int ptr, prev, next;
    int temp, is_int;
};
struct Elem {
    int row, col;
    double val;
    int r_prev, r_next, c_prev, c_next;
};
struct Lfe { int ref; double val; };
struct FCol { int j, stat; double a, c; int l0, ln; };

enum Kind {
    K_FREE_ROW, K_FIXED_COL, K_MAKE_EQ, K_MAKE_FIXED, K_EMPTY_COL, K_EQ_SINGLET, K_INEQ_SINGLET,
    K_IMPLIED_SLACK, K_IMPLIED_FREE, K_FORCING_ROW, K_INACTIVE_BOUND, K_LBND_COL, K_BINARIZE
};
struct Tse {
    int kind, p, q, stat, lb_changed, ub_changed, n, j;
    double apq, b, c, lb, ub, s;
    int l0, ln;   /* coefficient list in the Lfe pool (K_FORCING_ROW: range in the FCol pool) */
};

struct Form { double aj; int xj; };   /* one term of a copied linear form (copy_form, glpnpp04.js:101-114) */

} // namespace npp_detail
using namespace npp_detail;

struct glpb_npp {
    int orig_dir = 0, orig_m = 0, orig_n = 0, orig_nnz = 0;
    double c0 = 0.0;
    int nrows = 0, ncols = 0;
    int r_head = 0, r_tail = 0, c_head = 0, c_tail = 0;
    int sol = 0;
    int m = 0, n = 0, nnz = 0;     /* reduced problem (after build) */
    bool built = false;
    std::vector<Row> row{Row()};
    std::vector<Col> col{Col()};
    std::vector<Elem> el{Elem()};
    std::vector<Tse> stack;
    std::vector<Lfe> lfe;
    std::vector<FCol> fcol;
    std::vector<int> row_ref, col_ref;
    /* counters the reference prints from npp_integer */
    int n_packing = 0, n_covering = 0, n_reduced = 0, bin_vars = 0, bin_bins = 0, bin_rows = 0, bin_fails = 0;
    /* recovered solution */
    std::vector<signed char> r_stat, c_stat;
    std::vector<double> r_pi, c_value;

    /* ---- lists (glpnpp01.js:24-134) ---- */
    void insert_row(int r, int where)
    {
        Row &x = row[r];
        if (where == 0) {
            x.prev = 0; x.next = r_head;
            if (x.next == 0) r_tail = r; else row[x.next].prev = r;
            r_head = r;
        } else {
            x.prev = r_tail; x.next = 0;
            if (x.prev == 0) r_head = r; else row[x.prev].next = r;
            r_tail = r;
        }
    }
    void remove_row(int r)
    {
        Row &x = row[r];
        if (x.prev == 0) r_head = x.next; else row[x.prev].next = x.next;
        if (x.next == 0) r_tail = x.prev; else row[x.next].prev = x.prev;
    }
    void activate_row(int r)   { if (!row[r].temp) { row[r].temp = 1; remove_row(r); insert_row(r, 0); } }
    void deactivate_row(int r) { if (row[r].temp)  { row[r].temp = 0; remove_row(r); insert_row(r, 1); } }
    void insert_col(int c, int where)
    {
        Col &x = col[c];
        if (where == 0) {
            x.prev = 0; x.next = c_head;
            if (x.next == 0) c_tail = c; else col[x.next].prev = c;
            c_head = c;
        } else {
            x.prev = c_tail; x.next = 0;
            if (x.prev == 0) c_head = c; else col[x.prev].next = c;
            c_tail = c;
        }
    }
    void remove_col(int c)
    {
        Col &x = col[c];
        if (x.prev == 0) c_head = x.next; else col[x.prev].next = x.next;
        if (x.next == 0) c_tail = x.prev; else col[x.next].prev = x.prev;
    }
    void activate_col(int c)   { if (!col[c].temp) { col[c].temp = 1; remove_col(c); insert_col(c, 0); } }
    void deactivate_col(int c) { if (col[c].temp)  { col[c].temp = 0; remove_col(c); insert_col(c, 1); } }

    /* ---- construction (glpnpp01.js:136-260) ---- */
    int add_row()
    {
        Row x{}; x.lb = -INF; x.ub = +INF;
        row.push_back(x);
        int r = ++nrows;
        insert_row(r, 1);
        return r;
    }
    int add_col()
    {
        Col x{};
        col.push_back(x);
        int c = ++ncols;
        insert_col(c, 1);
        return c;
    }
    int add_aij(int r, int c, double val)
    {
        Elem e{};
        e.row = r; e.col = c; e.val = val;
        e.r_prev = 0; e.r_next = row[r].ptr;
        e.c_prev = 0; e.c_next = col[c].ptr;
        int k = (int)el.size();
        el.push_back(e);
        if (e.r_next) el[e.r_next].r_prev = k;
        if (e.c_next) el[e.c_next].c_prev = k;
        row[r].ptr = col[c].ptr = k;
        return k;
    }
    void erase_row(int r)
    {
        while (row[r].ptr) {
            Elem &a = el[row[r].ptr];
            row[r].ptr = a.r_next;
            if (a.c_prev == 0) col[a.col].ptr = a.c_next; else el[a.c_prev].c_next = a.c_next;
            if (a.c_next) el[a.c_next].c_prev = a.c_prev;
        }
    }
    void del_row(int r) { erase_row(r); remove_row(r); }
    void del_col(int c)
    {
        while (col[c].ptr) {
            Elem &a = el[col[c].ptr];
            col[c].ptr = a.c_next;
            if (a.r_prev == 0) row[a.row].ptr = a.r_next; else el[a.r_prev].r_next = a.r_next;
            if (a.r_next) el[a.r_next].r_prev = a.r_prev;
        }
        remove_col(c);
    }
    Tse &push(int kind)
    {
        Tse t{}; t.kind = kind; t.l0 = (int)lfe.size(); t.ln = 0;
        stack.push_back(t);
        return stack.back();
    }
    /* save a[i,q] of column q except element `skip` (reference: prepended list) */
    void save_col(Tse &t, int q, int skip)
    {
        t.l0 = (int)lfe.size();
        for (int a = col[q].ptr; a; a = el[a].c_next) {
            if (a == skip) continue;
            lfe.push_back(Lfe{el[a].row, el[a].val});
        }
        t.ln = (int)lfe.size() - t.l0;
    }
    /* subtract a fixed value of column q from the bounds of its rows (glpnpp02.js:379-389) */
    void shift_rows(int q, double v)
    {
        for (int a = col[q].ptr; a; a = el[a].c_next) {
            Row &i = row[el[a].row];
            if (i.lb == i.ub)
                i.ub = (i.lb -= el[a].val * v);
            else {
                if (i.lb != -INF) i.lb -= el[a].val * v;
                if (i.ub != +INF) i.ub -= el[a].val * v;
            }
        }
    }

    /* ---- transformations: glpnpp02.js ---- */
    void free_row(int p)                      /* :2-21 */
    {
        Tse &t = push(K_FREE_ROW);
        t.p = p;
        del_row(p);
    }
    void lbnd_col(int q)                      /* :196-244 */
    {
        Col &x = col[q];
        Tse &t = push(K_LBND_COL);
        t.q = q; t.b = x.lb;
        c0 += x.coef * x.lb;
        shift_rows(q, x.lb);
        if (x.ub != +INF) x.ub -= x.lb;
        x.lb = 0.0;
    }
    void fixed_col(int q)                     /* :357-392 */
    {
        Col &x = col[q];
        Tse &t = push(K_FIXED_COL);
        t.q = q; t.s = x.lb;
        c0 += x.coef * x.lb;
        shift_rows(q, x.lb);
        del_col(q);
    }
    int make_equality(int p)                  /* :394-435 */
    {
        Row &x = row[p];
        double eps = 1e-9 + 1e-12 * std::fabs(x.lb);
        if (x.ub - x.lb > eps) return 0;
        Tse &t = push(K_MAKE_EQ);
        t.p = p;
        double b = 0.5 * (x.ub + x.lb);
        double nint = std::floor(b + 0.5);
        if (std::fabs(b - nint) <= eps) b = nint;
        x.lb = x.ub = b;
        return 1;
    }
    int make_fixed(int q)                     /* :437-500 */
    {
        Col &x = col[q];
        double eps = 1e-9 + 1e-12 * std::fabs(x.lb);
        if (x.ub - x.lb > eps) return 0;
        Tse &t = push(K_MAKE_FIXED);
        t.q = q; t.c = x.coef;
        if (sol == SOL) save_col(t, q, 0);
        double s = 0.5 * (x.ub + x.lb);
        double nint = std::floor(s + 0.5);
        if (std::fabs(s - nint) <= eps) s = nint;
        x.lb = x.ub = s;
        return 1;
    }

    /* ---- transformations: glpnpp03.js ---- */
    int empty_row(int p)                      /* :1-14 */
    {
        const double eps = 1e-3;
        Row &x = row[p];
        if (x.lb > +eps || x.ub < -eps) return 1;
        x.lb = -INF; x.ub = +INF;
        free_row(p);
        return 0;
    }
    int empty_col(int q)                      /* :16-78 */
    {
        const double eps = 1e-3;
        Col &x = col[q];
        if (x.coef > +eps && x.lb == -INF) return 1;
        if (x.coef < -eps && x.ub == +INF) return 1;
        Tse &t = push(K_EMPTY_COL);
        t.q = q;
        bool lo;
        if (x.lb == -INF && x.ub == +INF) {
            t.stat = NF; x.lb = x.ub = 0.0;
        } else {
            if (x.ub == +INF) lo = true;
            else if (x.lb == -INF) lo = false;
            else if (x.lb != x.ub) {
                if (x.coef >= +DBL_EPSILON) lo = true;
                else if (x.coef <= -DBL_EPSILON) lo = false;
                else lo = std::fabs(x.lb) <= std::fabs(x.ub);
            } else {
                t.stat = NS;
                fixed_col(q);
                return 0;
            }
            if (lo) { t.stat = NL; x.ub = x.lb; } else { t.stat = NU; x.lb = x.ub; }
        }
        fixed_col(q);
        return 0;
    }
    int implied_value(int q, double s)        /* :80-119 */
    {
        Col &x = col[q];
        double eps, nint;
        if (x.is_int) {
            nint = std::floor(s + 0.5);
            if (std::fabs(s - nint) <= 1e-5) s = nint; else return 2;
        }
        if (x.lb != -INF) {
            eps = (x.is_int ? 1e-5 : 1e-5 + 1e-8 * std::fabs(x.lb));
            if (s < x.lb - eps) return 1;
            if (s < x.lb + 1e-3 * eps) { x.ub = x.lb; return 0; }
        }
        if (x.ub != +INF) {
            eps = (x.is_int ? 1e-5 : 1e-5 + 1e-8 * std::fabs(x.ub));
            if (s > x.ub + eps) return 1;
            if (s > x.ub - 1e-3 * eps) { x.lb = x.ub; return 0; }
        }
        x.lb = x.ub = s;
        return 0;
    }
    int eq_singlet(int p)                     /* :121-184 */
    {
        int a = row[p].ptr;
        int q = el[a].col;
        double s = row[p].lb / el[a].val;
        int ret = implied_value(q, s);
        if (ret != 0) return ret;
        Tse &t = push(K_EQ_SINGLET);
        t.p = p; t.q = q; t.apq = el[a].val; t.c = col[q].coef;
        if (sol != MIP) save_col(t, q, a);
        del_row(p);
        return 0;
    }
    int implied_lower(int q, double l)        /* :186-237 */
    {
        Col &x = col[q];
        int ret;
        double eps, nint;
        if (x.is_int) {
            nint = std::floor(l + 0.5);
            if (std::fabs(l - nint) <= 1e-5) l = nint; else l = std::ceil(l);
        }
        if (x.lb != -INF) {
            eps = (x.is_int ? 1e-3 : 1e-3 + 1e-6 * std::fabs(x.lb));
            if (l < x.lb + eps) return 0;
        }
        if (x.ub != +INF) {
            eps = (x.is_int ? 1e-5 : 1e-5 + 1e-8 * std::fabs(x.ub));
            if (l > x.ub + eps) return 4;
            if (l > x.ub - 1e-3 * eps) { x.lb = x.ub; return 3; }
        }
        if (x.lb == -INF) ret = 2;
        else if (x.is_int && l > x.lb + 0.5) ret = 2;
        else if (l > x.lb + 0.30 * (1.0 + std::fabs(x.lb))) ret = 2;
        else ret = 1;
        x.lb = l;
        return ret;
    }
    int implied_upper(int q, double u)        /* :239-289 */
    {
        Col &x = col[q];
        int ret;
        double eps, nint;
        if (x.is_int) {
            nint = std::floor(u + 0.5);
            if (std::fabs(u - nint) <= 1e-5) u = nint; else u = std::floor(u);
        }
        if (x.ub != +INF) {
            eps = (x.is_int ? 1e-3 : 1e-3 + 1e-6 * std::fabs(x.ub));
            if (u > x.ub - eps) return 0;
        }
        if (x.lb != -INF) {
            eps = (x.is_int ? 1e-5 : 1e-5 + 1e-8 * std::fabs(x.lb));
            if (u < x.lb - eps) return 4;
            if (u < x.lb + 1e-3 * eps) { x.ub = x.lb; return 3; }
        }
        if (x.ub == +INF) ret = 2;
        else if (x.is_int && u < x.ub - 0.5) ret = 2;
        else if (u < x.ub - 0.30 * (1.0 + std::fabs(x.ub))) ret = 2;
        else ret = 1;
        x.ub = u;
        return ret;
    }
    int ineq_singlet(int p)                   /* :291-508 */
    {
        Row &x = row[p];
        int apq = x.ptr;
        int q = el[apq].col;
        double v = el[apq].val, ll, uu;
        int lb_changed, ub_changed;
        if (v > 0.0) {
            ll = (x.lb == -INF ? -INF : x.lb / v);
            uu = (x.ub == +INF ? +INF : x.ub / v);
        } else {
            ll = (x.ub == +INF ? -INF : x.ub / v);
            uu = (x.lb == -INF ? +INF : x.lb / v);
        }
        if (ll == -INF) lb_changed = 0;
        else {
            lb_changed = implied_lower(q, ll);
            if (lb_changed == 4) return 4;
        }
        if (uu == +INF) ub_changed = 0;
        else if (lb_changed == 3) ub_changed = 0;
        else {
            ub_changed = implied_upper(q, uu);
            if (ub_changed == 4) return 4;
        }
        if (!lb_changed && !ub_changed) {
            x.lb = -INF; x.ub = +INF;
            free_row(p);
            return 0;
        }
        Tse &t = push(K_INEQ_SINGLET);
        t.p = p; t.q = q; t.apq = v; t.c = col[q].coef;
        t.lb = x.lb; t.ub = x.ub;
        t.lb_changed = lb_changed; t.ub_changed = ub_changed;
        if (sol != MIP) save_col(t, q, apq);
        del_row(p);
        return lb_changed >= ub_changed ? lb_changed : ub_changed;
    }
    void implied_slack(int q)                 /* :510-592 */
    {
        int a = col[q].ptr;
        int p = el[a].row;
        Tse &t = push(K_IMPLIED_SLACK);
        t.p = p; t.q = q; t.apq = el[a].val; t.b = row[p].lb; t.c = col[q].coef;
        const double apq = t.apq, b = t.b, c = t.c;
        t.l0 = (int)lfe.size();
        for (int k = row[p].ptr; k; k = el[k].r_next) {
            if (el[k].col == q) continue;
            lfe.push_back(Lfe{el[k].col, el[k].val});
            col[el[k].col].coef -= c * (el[k].val / apq);
        }
        stack.back().ln = (int)lfe.size() - stack.back().l0;
        c0 += c * (b / apq);
        Row &x = row[p];
        const Col &y = col[q];
        if (apq > 0.0) {
            x.lb = (y.ub == +INF ? -INF : b - apq * y.ub);
            x.ub = (y.lb == -INF ? +INF : b - apq * y.lb);
        } else {
            x.lb = (y.lb == -INF ? -INF : b - apq * y.lb);
            x.ub = (y.ub == +INF ? +INF : b - apq * y.ub);
        }
        del_col(q);
    }
    int implied_free(int q)                   /* :594-744 */
    {
        Col &y = col[q];
        int apq = y.ptr;
        int p = el[apq].row;
        Row &x = row[p];
        double alfa, beta, l, u, pi, eps;
        const double v = el[apq].val;
        alfa = x.lb;
        if (alfa != -INF) {
            for (int a = x.ptr; a; a = el[a].r_next) {
                if (a == apq) continue;
                const Col &z = col[el[a].col];
                if (el[a].val > 0.0) {
                    if (z.ub == +INF) { alfa = -INF; break; }
                    alfa -= el[a].val * z.ub;
                } else {
                    if (z.lb == -INF) { alfa = -INF; break; }
                    alfa -= el[a].val * z.lb;
                }
            }
        }
        beta = x.ub;
        if (beta != +INF) {
            for (int a = x.ptr; a; a = el[a].r_next) {
                if (a == apq) continue;
                const Col &z = col[el[a].col];
                if (el[a].val > 0.0) {
                    if (z.lb == -INF) { beta = +INF; break; }
                    beta -= el[a].val * z.lb;
                } else {
                    if (z.ub == +INF) { beta = +INF; break; }
                    beta -= el[a].val * z.ub;
                }
            }
        }
        if (v > 0.0) l = (alfa == -INF ? -INF : alfa / v);
        else         l = (beta == +INF ? -INF : beta / v);
        if (v > 0.0) u = (beta == +INF ? +INF : beta / v);
        else         u = (alfa == -INF ? +INF : alfa / v);
        if (y.lb != -INF) {
            eps = 1e-9 + 1e-12 * std::fabs(y.lb);
            if (l < y.lb - eps) return 1;
        }
        if (y.ub != +INF) {
            eps = 1e-9 + 1e-12 * std::fabs(y.ub);
            if (u > y.ub + eps) return 1;
        }
        y.lb = -INF; y.ub = +INF;
        Tse &t = push(K_IMPLIED_FREE);
        t.p = p; t.stat = -1;
        pi = y.coef / v;
        int act;   /* 0 = lower bound made active, 1 = upper */
        if (pi > +DBL_EPSILON) {
            if (x.lb != -INF) act = 0;
            else { if (pi > +1e-5) return 2; act = 1; }
        } else if (pi < -DBL_EPSILON) {
            if (x.ub != +INF) act = 1;
            else { if (pi < -1e-5) return 2; act = 0; }
        } else {
            if (x.ub == +INF) act = 0;
            else if (x.lb == -INF) act = 1;
            else act = (std::fabs(x.lb) <= std::fabs(x.ub)) ? 0 : 1;
        }
        if (act == 0) { t.stat = NL; x.ub = x.lb; } else { t.stat = NU; x.lb = x.ub; }
        return 0;
    }
    int forcing_row(int p, int at)            /* :861-1013 */
    {
        Row &x = row[p];
        double big = 1.0;
        for (int a = x.ptr; a; a = el[a].r_next)
            if (big < std::fabs(el[a].val)) big = std::fabs(el[a].val);
        for (int a = x.ptr; a; a = el[a].r_next)
            if (std::fabs(el[a].val) < 1e-7 * big) return 1;
        Tse &t0 = push(K_FORCING_ROW);
        t0.p = p;
        if (x.lb == x.ub) t0.stat = NS;
        else if (at == 0) t0.stat = NL;
        else t0.stat = NU;
        const int f0 = (int)fcol.size();
        for (int a = x.ptr; a; a = el[a].r_next) {
            const int j = el[a].col;
            Col &y = col[j];
            const bool lower = (at == 0 && el[a].val < 0.0) || (at != 0 && el[a].val > 0.0);
            if (lower) y.ub = y.lb; else y.lb = y.ub;
            if (sol != MIP) {
                FCol f{};
                f.j = j; f.stat = lower ? NL : NU; f.a = el[a].val; f.c = y.coef;
                f.l0 = (int)lfe.size();
                for (int k = y.ptr; k; k = el[k].c_next) {
                    if (k == a) continue;
                    lfe.push_back(Lfe{el[k].row, el[k].val});
                }
                f.ln = (int)lfe.size() - f.l0;
                fcol.push_back(f);
            }
        }
        stack.back().l0 = f0;
        stack.back().ln = (int)fcol.size() - f0;
        x.lb = -INF; x.ub = +INF;
        return 0;
    }
    int analyze_row(int p)                    /* :1015-1096 */
    {
        const Row &x = row[p];
        int ret = 0x00;
        double l, u, eps;
        l = 0.0;
        for (int a = x.ptr; a; a = el[a].r_next) {
            const Col &z = col[el[a].col];
            if (el[a].val > 0.0) {
                if (z.lb == -INF) { l = -INF; break; }
                l += el[a].val * z.lb;
            } else {
                if (z.ub == +INF) { l = -INF; break; }
                l += el[a].val * z.ub;
            }
        }
        u = 0.0;
        for (int a = x.ptr; a; a = el[a].r_next) {
            const Col &z = col[el[a].col];
            if (el[a].val > 0.0) {
                if (z.ub == +INF) { u = +INF; break; }
                u += el[a].val * z.ub;
            } else {
                if (z.lb == -INF) { u = +INF; break; }
                u += el[a].val * z.lb;
            }
        }
        if (x.lb != -INF) {
            eps = 1e-3 + 1e-6 * std::fabs(x.lb);
            if (x.lb - eps > u) return 0x33;
        }
        if (x.ub != +INF) {
            eps = 1e-3 + 1e-6 * std::fabs(x.ub);
            if (x.ub + eps < l) return 0x33;
        }
        if (x.lb != -INF) {
            eps = 1e-9 + 1e-12 * std::fabs(x.lb);
            if (x.lb - eps > l) {
                if (x.lb + eps <= u) ret |= 0x01; else ret |= 0x02;
            }
        }
        if (x.ub != +INF) {
            eps = 1e-9 + 1e-12 * std::fabs(x.ub);
            if (x.ub + eps < u) {
                if (x.ub - eps >= l) ret |= 0x10; else ret |= 0x20;
            }
        }
        return ret;
    }
    void inactive_bound(int p, int which)     /* :1098-1138 */
    {
        Row &x = row[p];
        if (sol == SOL) {
            Tse &t = push(K_INACTIVE_BOUND);
            t.p = p;
            if (x.ub == +INF) t.stat = NL;
            else if (x.lb == -INF) t.stat = NU;
            else if (x.lb != x.ub) t.stat = (which == 0 ? NU : NL);
            else t.stat = NS;
        }
        if (which == 0) x.lb = -INF; else x.ub = +INF;
    }
    void implied_bounds(int p)                /* :1140-1260 */
    {
        const Row &x = row[p];
        double big = 1.0, eps, temp;
        for (int a = x.ptr; a; a = el[a].r_next) {
            Col &z = col[el[a].col];
            z.ll = -INF; z.uu = +INF;
            if (big < std::fabs(el[a].val)) big = std::fabs(el[a].val);
        }
        eps = 1e-6 * big;
        for (int side = 0; side < 2; side++) {
            /* side 0: row lower bound, side 1: row upper bound */
            if (side == 0 ? x.lb == -INF : x.ub == +INF) continue;
            int apk = 0;
            bool skip = false;
            for (int a = x.ptr; a; a = el[a].r_next) {
                const Col &z = col[el[a].col];
                const double v = el[a].val;
                const bool open = side == 0 ? ((v > 0.0 && z.ub == +INF) || (v < 0.0 && z.lb == -INF))
                                            : ((v > 0.0 && z.lb == -INF) || (v < 0.0 && z.ub == +INF));
                if (open) {
                    if (apk == 0) apk = a; else { skip = true; break; }
                }
            }
            if (skip) continue;
            temp = side == 0 ? x.lb : x.ub;
            for (int a = x.ptr; a; a = el[a].r_next) {
                if (a == apk) continue;
                const Col &z = col[el[a].col];
                const double v = el[a].val;
                if (side == 0) { if (v > 0.0) temp -= v * z.ub; else temp -= v * z.lb; }
                else           { if (v > 0.0) temp -= v * z.lb; else temp -= v * z.ub; }
            }
            if (apk == 0) {
                for (int a = x.ptr; a; a = el[a].r_next) {
                    Col &z = col[el[a].col];
                    const double v = el[a].val;
                    if (side == 0) {
                        if (v >= +eps) z.ll = z.ub + temp / v;
                        else if (v <= -eps) z.uu = z.lb + temp / v;
                    } else {
                        if (v >= +eps) z.uu = z.lb + temp / v;
                        else if (v <= -eps) z.ll = z.ub + temp / v;
                    }
                }
            } else {
                Col &z = col[el[apk].col];
                const double v = el[apk].val;
                if (side == 0) {
                    if (v >= +eps) z.ll = temp / v; else if (v <= -eps) z.uu = temp / v;
                } else {
                    if (v >= +eps) z.uu = temp / v; else if (v <= -eps) z.ll = temp / v;
                }
            }
        }
    }

    /* ---- driver: glpnpp05.js ---- */
    void clean_prob()                         /* :2-62 */
    {
        int r, c, nx;
        for (r = r_head; r; r = nx) {
            nx = row[r].next;
            if (row[r].lb == -INF && row[r].ub == +INF) free_row(r);
        }
        for (r = r_head; r; r = nx) {
            nx = row[r].next;
            if (row[r].lb != -INF && row[r].ub != +INF && row[r].lb < row[r].ub) make_equality(r);
        }
        for (c = c_head; c; c = nx) {
            nx = col[c].next;
            if (col[c].lb == col[c].ub) fixed_col(c);
        }
        for (c = c_head; c; c = nx) {
            nx = col[c].next;
            if (col[c].lb != -INF && col[c].ub != +INF && col[c].lb < col[c].ub)
                if (make_fixed(c) == 1) fixed_col(c);
        }
    }
    /* columns of a forcing row were fixed, the row was made free (fixup(), :209-227) */
    int fixup(int p)
    {
        int nx;
        for (int a = row[p].ptr; a; a = nx) {
            const int c = el[a].col;
            nx = el[a].r_next;
            for (int k = col[c].ptr; k; k = el[k].c_next) activate_row(el[k].row);
            fixed_col(c);
        }
        free_row(p);
        return 0;
    }
    int improve_bounds(int p, int flag)       /* :231-293 */
    {
        int count = 0, nx, ret;
        implied_bounds(p);
        for (int a = row[p].ptr; a; a = nx) {
            const int c = el[a].col;
            nx = el[a].r_next;
            for (int kase = 0; kase <= 1; kase++) {
                const double lb = col[c].lb, ub = col[c].ub;
                if (kase == 0) {
                    if (col[c].ll == -INF) continue;
                    ret = implied_lower(c, col[c].ll);
                } else {
                    if (col[c].uu == +INF) continue;
                    ret = implied_upper(c, col[c].uu);
                }
                if (ret == 0 || ret == 1) {
                    col[c].lb = lb; col[c].ub = ub;
                } else if (ret == 2 || ret == 3) {
                    count++;
                    if (flag)
                        for (int k = col[c].ptr; k; k = el[k].c_next)
                            if (el[k].row != p) activate_row(el[k].row);
                    if (ret == 3) { fixed_col(c); break; }
                } else
                    return -1;
            }
        }
        return count;
    }
    int process_row(int p, int hard)          /* :64-229 */
    {
        int ret;
        if (row[p].ptr == 0) {
            ret = empty_row(p);
            return ret == 0 ? 0 : ENOPFS;
        }
        if (el[row[p].ptr].r_next == 0) {
            const int c = el[row[p].ptr].col;
            if (row[p].lb == row[p].ub) {
                ret = eq_singlet(p);
                if (ret == 0) {
                    for (int k = col[c].ptr; k; k = el[k].c_next) activate_row(el[k].row);
                    fixed_col(c);
                    return 0;
                }
                return ENOPFS;
            }
            ret = ineq_singlet(p);
            if (0 <= ret && ret <= 3) {
                activate_col(c);
                if (ret >= 2)
                    for (int k = col[c].ptr; k; k = el[k].c_next) activate_row(el[k].row);
                if (ret == 3) fixed_col(c);
                return 0;
            }
            return ENOPFS;
        }
        ret = analyze_row(p);
        if (ret == 0x33) return ENOPFS;
        if ((ret & 0x0F) == 0x00) {
            if (row[p].lb != -INF) inactive_bound(p, 0);
        } else if ((ret & 0x0F) == 0x02) {
            if (forcing_row(p, 0) == 0) return fixup(p);
        }
        if ((ret & 0xF0) == 0x00) {
            if (row[p].ub != +INF) inactive_bound(p, 1);
        } else if ((ret & 0xF0) == 0x20) {
            if (forcing_row(p, 1) == 0) return fixup(p);
        }
        if (row[p].lb == -INF && row[p].ub == +INF) {
            for (int a = row[p].ptr; a; a = el[a].r_next) activate_col(el[a].col);
            free_row(p);
            return 0;
        }
        if (sol == MIP && hard)
            if (improve_bounds(p, 1) < 0) return ENOPFS;
        return 0;
    }
    /* implied slack variable: slack(), :323-343 */
    int slack(int q, int p)
    {
        implied_slack(q);
        if (row[p].lb == -INF && row[p].ub == +INF) {
            for (int a = row[p].ptr; a; a = el[a].r_next) activate_col(el[a].col);
            free_row(p);
        } else
            activate_row(p);
        return 0;
    }
    int process_col(int q)                    /* :295-374 */
    {
        int ret;
        if (col[q].ptr == 0) {
            ret = empty_col(q);
            return ret == 0 ? 0 : ENODFS;
        }
        if (el[col[q].ptr].c_next == 0) {
            const int p = el[col[q].ptr].row;
            if (row[p].lb == row[p].ub) {
                if (!col[q].is_int) return slack(q, p);
            } else if (!col[q].is_int) {
                ret = implied_free(q);
                if (ret == 0) return slack(q, p);
                if (ret == 2) return ENODFS;
            }
        }
        return 0;
    }
    int process_prob(int hard)                /* :376-428 */
    {
        int ret;
        clean_prob();
        for (int r = r_head; r; r = row[r].next) row[r].temp = 1;
        for (int c = c_head; c; c = col[c].next) col[c].temp = 1;
        bool processing = true;
        while (processing) {
            processing = false;
            for (;;) {
                const int r = r_head;
                if (r == 0 || !row[r].temp) break;
                deactivate_row(r);
                ret = process_row(r, hard);
                if (ret != 0) return ret;
                processing = true;
            }
            for (;;) {
                const int c = c_head;
                if (c == 0 || !col[c].temp) break;
                deactivate_col(c);
                ret = process_col(c);
                if (ret != 0) return ret;
                processing = true;
            }
        }
        if (sol == MIP && !hard)
            for (int r = r_head; r; r = row[r].next)
                if (improve_bounds(r, 0) < 0) return ENOPFS;
        return 0;
    }

    /* ---- MIP-only transformations: glpnpp04.js ---- */
    bool binary(int c) const { return col[c].is_int && col[c].lb == 0.0 && col[c].ub == 1.0; }
    int binarize_prob()                       /* :2-99 */
    {
        bin_fails = bin_vars = bin_bins = bin_rows = 0;
        for (int c = c_tail; c; c = col[c].prev) {
            if (!col[c].is_int) continue;
            if (col[c].lb == col[c].ub) continue;
            if (col[c].lb == 0.0 && col[c].ub == 1.0) continue;
            if (col[c].lb < -1e6 || col[c].ub > +1e6 || col[c].ub - col[c].lb > 4095.0) {
                bin_fails++;
                continue;
            }
            bin_vars++;
            if (col[c].lb != 0.0) lbnd_col(c);
            const int u = (int)col[c].ub;
            if (u == 1) continue;
            int n = 2, temp = 4;
            while (u >= temp) { n++; temp += temp; }
            bin_bins += n;
            const size_t ti = stack.size();
            { Tse &t = push(K_BINARIZE); t.q = c; t.j = 0; t.n = n; }
            int r = 0;
            if (u < temp - 1) {
                r = add_row(); bin_rows++;
                row[r].lb = -INF; row[r].ub = u;
            }
            col[c].ub = 1.0;
            if (r) add_aij(r, c, 1.0);
            int k;
            for (k = 1, temp = 2; k < n; k++, temp += temp) {
                const int b = add_col();
                col[b].is_int = 1;
                col[b].lb = 0.0; col[b].ub = 1.0;
                col[b].coef = temp * col[c].coef;
                if (stack[ti].j == 0) stack[ti].j = b;
                for (int a = col[c].ptr; a; a = el[a].c_next)
                    add_aij(el[a].row, b, temp * el[a].val);
            }
        }
        return bin_fails;
    }
    void copy_form(int r, double s, std::vector<Form> &f) const     /* :101-114: the copy is the row reversed */
    {
        f.clear();
        for (int a = row[r].ptr; a; a = el[a].r_next) f.push_back(Form{s * el[a].val, el[a].col});
        for (size_t i = 0, j = f.size(); i + 1 < j; i++, j--) std::swap(f[i], f[j - 1]);
    }
    static bool unit_form(const std::vector<Form> &f, int &neg)
    {
        neg = 0;
        for (const Form &e : f) {
            if (e.aj == +1.0) continue;
            if (e.aj == -1.0) neg++; else return false;
        }
        return true;
    }
    static void to_unit(std::vector<Form> &f, double &b)
    {
        b = 1.0;
        for (Form &e : f) {
            if (e.aj > 0.0) e.aj = +1.0; else { e.aj = -1.0; b -= 1.0; }
        }
    }
    static int hidden_packing_form(std::vector<Form> &f, double &bout)   /* :141-210 */
    {
        double b = bout;
        int neg;
        if (unit_form(f, neg) && b == (1 - neg)) return 1;
        for (const Form &e : f) if (e.aj < 0) b -= e.aj;
        for (const Form &e : f) if (std::fabs(e.aj) > b) return 0;
        int ej = -1, ek = -1;
        for (int i = 0; i < (int)f.size(); i++)
            if (ej < 0 || std::fabs(f[ej].aj) > std::fabs(f[i].aj)) ej = i;
        for (int i = 0; i < (int)f.size(); i++)
            if (i != ej)
                if (ek < 0 || std::fabs(f[ek].aj) > std::fabs(f[i].aj)) ek = i;
        const double eps = 1e-3 + 1e-6 * std::fabs(b);
        if (std::fabs(f[ej].aj) + std::fabs(f[ek].aj) <= b + eps) return 0;
        to_unit(f, bout);
        return 2;
    }
    static int hidden_covering_form(std::vector<Form> &f, double &bout)  /* :386-445 */
    {
        double b = bout;
        int neg;
        if (unit_form(f, neg) && b == (1 - neg)) return 1;
        for (const Form &e : f) if (e.aj < 0) b -= e.aj;
        if (b < 1e-3) return 0;
        const double eps = 1e-9 + 1e-12 * std::fabs(b);
        for (const Form &e : f) if (std::fabs(e.aj) < b - eps) return 0;
        to_unit(f, bout);
        return 2;
    }
    int reduce_form(std::vector<Form> &f, double &bout) const           /* :546-607 */
    {
        int count = 0;
        double h = 0.0, inf_t, new_a, b = bout;
        for (const Form &e : f) {
            if (e.aj > 0.0) {
                if (col[e.xj].lb == -INF) return count;
                h += e.aj * col[e.xj].lb;
            } else {
                if (col[e.xj].ub == +INF) return count;
                h += e.aj * col[e.xj].ub;
            }
        }
        for (Form &e : f) {
            if (!binary(e.xj)) continue;
            if (e.aj > 0.0) {
                inf_t = h;
                if (b - e.aj < inf_t && inf_t < b) {
                    new_a = b - inf_t;
                    if (new_a >= +1e-3 && e.aj - new_a >= 0.01 * (1.0 + e.aj)) {
                        e.aj = new_a;
                        count++;
                    }
                }
            } else {
                inf_t = h - e.aj;
                if (b < inf_t && inf_t < b - e.aj) {
                    new_a = e.aj + (inf_t - b);
                    if (new_a <= -1e-3 && new_a - e.aj >= 0.01 * (1.0 - e.aj)) {
                        e.aj = new_a;
                        h += (inf_t - b);
                        b = inf_t;
                        count++;
                    }
                }
            }
        }
        bout = b;
        return count;
    }
    /* replace row r by the form f with the single bound b on `side` (0: <= b, 1: >= b); a double-sided row
     * first leaves a copy carrying its other bound, which the caller continues with (:253-281, :488-516, :639-668) */
    int replace_row(int r, const std::vector<Form> &f, double b, int side, bool other_is_lower)
    {
        int copy = 0;
        if (!(row[r].lb == -INF || row[r].ub == +INF)) {
            copy = add_row();
            if (other_is_lower) { row[copy].lb = row[r].lb; row[copy].ub = +INF; }
            else                { row[copy].lb = -INF; row[copy].ub = row[r].ub; }
            for (int a = row[r].ptr; a; a = el[a].r_next) add_aij(copy, el[a].col, el[a].val);
        }
        erase_row(r);
        if (side == 0) { row[r].lb = -INF; row[r].ub = b; } else { row[r].lb = b; row[r].ub = +INF; }
        for (const Form &e : f) add_aij(r, e.xj, e.aj);
        return copy;
    }
    int hidden_packing(int r)                 /* :212-285 */
    {
        std::vector<Form> f;
        int count = 0;
        for (int kase = 0; kase <= 1; kase++) {
            double b;
            if (kase == 0) {
                if (row[r].ub == +INF) continue;
                copy_form(r, +1.0, f); b = +row[r].ub;
            } else {
                if (row[r].lb == -INF) continue;
                copy_form(r, -1.0, f); b = -row[r].lb;
            }
            const int ret = hidden_packing_form(f, b);
            if ((kase == 1 && ret == 1) || ret == 2) {
                count++;
                const int copy = replace_row(r, f, b, 0, kase == 0);
                if (copy) r = copy;
            }
        }
        return count;
    }
    int hidden_covering(int r)                /* :447-520 */
    {
        std::vector<Form> f;
        int count = 0;
        for (int kase = 0; kase <= 1; kase++) {
            double b;
            if (kase == 0) {
                if (row[r].lb == -INF) continue;
                copy_form(r, +1.0, f); b = +row[r].lb;
            } else {
                if (row[r].ub == +INF) continue;
                copy_form(r, -1.0, f); b = -row[r].ub;
            }
            const int ret = hidden_covering_form(f, b);
            if ((kase == 1 && ret == 1) || ret == 2) {
                count++;
                const int copy = replace_row(r, f, b, 1, kase != 0);
                if (copy) r = copy;
            }
        }
        return count;
    }
    int reduce_ineq_coef(int r)               /* :609-672 */
    {
        std::vector<Form> f;
        int total = 0;
        for (int kase = 0; kase <= 1; kase++) {
            double b;
            if (kase == 0) {
                if (row[r].lb == -INF) continue;
                copy_form(r, +1.0, f); b = +row[r].lb;
            } else {
                if (row[r].ub == +INF) continue;
                copy_form(r, -1.0, f); b = -row[r].ub;
            }
            const int cnt = reduce_form(f, b);
            if (cnt > 0) {
                const int copy = replace_row(r, f, b, 1, kase != 0);
                if (copy) r = copy;
            }
            total += cnt;
        }
        return total;
    }
    int integer(int binarize)                 /* glpnpp05.js:437-521 */
    {
        int ret = process_prob(1);
        if (ret != 0) return ret;
        if (binarize) binarize_prob();
        int r, pv;
        n_packing = 0;
        for (r = r_tail; r; r = pv) {
            pv = row[r].prev;
            if (row[r].lb == -INF && row[r].ub == +INF) continue;
            if (row[r].lb == row[r].ub) continue;
            if (row[r].ptr == 0 || el[row[r].ptr].r_next == 0) continue;
            int a;
            for (a = row[r].ptr; a; a = el[a].r_next) if (!binary(el[a].col)) break;
            if (a) continue;
            n_packing += hidden_packing(r);
        }
        n_covering = 0;
        for (r = r_tail; r; r = pv) {
            pv = row[r].prev;
            if (row[r].lb == -INF && row[r].ub == +INF) continue;
            if (row[r].lb == row[r].ub) continue;
            if (row[r].ptr == 0 || el[row[r].ptr].r_next == 0 || el[el[row[r].ptr].r_next].r_next == 0) continue;
            int a;
            for (a = row[r].ptr; a; a = el[a].r_next) if (!binary(el[a].col)) break;
            if (a) continue;
            n_covering += hidden_covering(r);
        }
        n_reduced = 0;
        for (r = r_tail; r; r = pv) {
            pv = row[r].prev;
            if (row[r].lb == row[r].ub) continue;
            n_reduced += reduce_ineq_coef(r);
        }
        return 0;
    }

    /* ---- recovery (the closures pushed by the transformations above) ---- */
    double minus_dot_pi(double start, int l0, int ln) const
    {
        double t = start;
        for (int k = l0 + ln - 1; k >= l0; k--) t -= lfe[k].val * r_pi[lfe[k].ref];
        return t;
    }
    int recover(const Tse &t)
    {
        switch (t.kind) {
        case K_FREE_ROW:                      /* glpnpp02.js:9-16 */
            if (sol == SOL) r_stat[t.p] = BS;
            if (sol != MIP) r_pi[t.p] = 0.0;
            return 0;
        case K_LBND_COL:                      /* glpnpp02.js:207-222 */
            if (sol == SOL && !(c_stat[t.q] == BS || c_stat[t.q] == NL || c_stat[t.q] == NU)) return 1;
            c_value[t.q] = t.b + c_value[t.q];
            return 0;
        case K_FIXED_COL:                     /* glpnpp02.js:366-372 */
            if (sol == SOL) c_stat[t.q] = NS;
            c_value[t.q] = t.s;
            return 0;
        case K_MAKE_EQ:                       /* glpnpp02.js:408-425 */
            if (sol == SOL) {
                if (r_stat[t.p] == BS) ;
                else if (r_stat[t.p] == NS) r_stat[t.p] = (r_pi[t.p] >= 0.0 ? NL : NU);
                else return 1;
            }
            return 0;
        case K_MAKE_FIXED:                    /* glpnpp02.js:453-477 */
            if (sol == SOL) {
                if (c_stat[t.q] == BS) ;
                else if (c_stat[t.q] == NS) {
                    const double lambda = minus_dot_pi(t.c, t.l0, t.ln);
                    c_stat[t.q] = (lambda >= 0.0 ? NL : NU);
                } else return 1;
            }
            return 0;
        case K_EMPTY_COL:                     /* glpnpp03.js:29-34 */
            if (sol == SOL) c_stat[t.q] = (signed char)t.stat;
            return 0;
        case K_EQ_SINGLET:                    /* glpnpp03.js:141-162 */
            if (sol == SOL) {
                if (c_stat[t.q] != NS) return 1;
                r_stat[t.p] = NS;
                c_stat[t.q] = BS;
            }
            if (sol != MIP) r_pi[t.p] = minus_dot_pi(t.c, t.l0, t.ln) / t.apq;
            return 0;
        case K_INEQ_SINGLET: {                /* glpnpp03.js:345-482 */
            if (sol == MIP) return 0;
            const double lambda = minus_dot_pi(t.c, t.l0, t.ln);
            if (sol != SOL) return 0;
            int cs = c_stat[t.q];
            if (cs == NS) {
                if (lambda > +1e-7 &&
                    ((t.apq > 0.0 && t.lb != -INF) || (t.apq < 0.0 && t.ub != +INF) || !t.lb_changed)) {
                    c_stat[t.q] = NL; cs = NL;
                } else if (lambda < -1e-7 &&
                    ((t.apq > 0.0 && t.ub != +INF) || (t.apq < 0.0 && t.lb != -INF) || !t.ub_changed)) {
                    c_stat[t.q] = NU; cs = NU;
                } else {
                    if (t.lb != -INF && t.ub == +INF) r_stat[t.p] = NL;
                    else if (t.lb == -INF && t.ub != +INF) r_stat[t.p] = NU;
                    else if (t.lb != -INF && t.ub != +INF)
                        r_stat[t.p] = (t.apq * c_value[t.q] <= 0.5 * (t.lb + t.ub)) ? NL : NU;
                    else return 1;
                    c_stat[t.q] = BS;
                    r_pi[t.p] = lambda / t.apq;
                    return 0;
                }
            }
            if (cs == BS) {
                r_stat[t.p] = BS; r_pi[t.p] = 0.0;
            } else if (cs == NL || cs == NU) {
                const bool implied = cs == NL ? t.lb_changed != 0 : t.ub_changed != 0;
                if (implied) {
                    if (cs == NL) r_stat[t.p] = (t.apq > 0.0 ? NL : NU);
                    else          r_stat[t.p] = (t.apq > 0.0 ? NU : NL);
                    c_stat[t.q] = BS;
                    r_pi[t.p] = lambda / t.apq;
                } else {
                    r_stat[t.p] = BS; r_pi[t.p] = 0.0;
                }
            } else return 1;
            return 0;
        }
        case K_IMPLIED_SLACK: {               /* glpnpp03.js:526-557 */
            if (sol == SOL) {
                const int rs = r_stat[t.p];
                if (rs == BS || rs == NF) c_stat[t.q] = (signed char)rs;
                else if (rs == NL) c_stat[t.q] = (t.apq > 0.0 ? NU : NL);
                else if (rs == NU) c_stat[t.q] = (t.apq > 0.0 ? NL : NU);
                else return 1;
                r_stat[t.p] = NS;
            }
            if (sol != MIP) r_pi[t.p] += t.c / t.apq;
            double temp = t.b;
            for (int k = t.l0 + t.ln - 1; k >= t.l0; k--) temp -= lfe[k].val * c_value[lfe[k].ref];
            c_value[t.q] = temp / t.apq;
            return 0;
        }
        case K_IMPLIED_FREE:                  /* glpnpp03.js:674-689 */
            if (sol == SOL) {
                if (r_stat[t.p] == BS) ;
                else if (r_stat[t.p] == NS) {
                    if (!(t.stat == NL || t.stat == NU)) return 1;
                    r_stat[t.p] = (signed char)t.stat;
                } else return 1;
            }
            return 0;
        case K_FORCING_ROW: {                 /* glpnpp03.js:880-947 */
            if (sol == MIP) return 0;
            if (sol == SOL) {
                if (r_stat[t.p] != BS) return 1;
                for (int k = t.l0 + t.ln - 1; k >= t.l0; k--) {
                    if (c_stat[fcol[k].j] != NS) return 1;
                    c_stat[fcol[k].j] = (signed char)fcol[k].stat;
                }
            }
            int piv = -1;
            double big = 0.0, pd = 0.0;
            for (int k = t.l0 + t.ln - 1; k >= t.l0; k--) {
                const FCol &f = fcol[k];
                const double d = minus_dot_pi(f.c, f.l0, f.ln);
                const double temp = std::fabs(d / f.a);
                if (f.stat == NL) {
                    if (d < 0.0 && big < temp) { piv = k; big = temp; pd = d; }
                } else if (f.stat == NU) {
                    if (d > 0.0 && big < temp) { piv = k; big = temp; pd = d; }
                } else return 1;
            }
            if (piv >= 0) {
                if (sol == SOL) {
                    r_stat[t.p] = (signed char)t.stat;
                    c_stat[fcol[piv].j] = BS;
                }
                r_pi[t.p] = pd / fcol[piv].a;
            }
            return 0;
        }
        case K_INACTIVE_BOUND:                /* glpnpp03.js:1104-1115 */
            if (sol != SOL) return 1;
            if (r_stat[t.p] != BS) r_stat[t.p] = (signed char)t.stat;
            return 0;
        case K_BINARIZE: {                    /* glpnpp04.js:46-55 */
            double sum = c_value[t.q];
            int k, temp;
            for (k = 1, temp = 2; k < t.n; k++, temp += temp) sum += temp * c_value[t.j + (k - 1)];
            c_value[t.q] = sum;
            return 0;
        }
        }
        return 1;
    }
};

/* ------------------------------------------------------------------ C ABI */

extern "C" {

glpb_npp *glpb_npp_create(void) { return new (std::nothrow) glpb_npp(); }

void glpb_npp_destroy(glpb_npp *npp) { delete npp; }

int glpb_npp_load_prob(glpb_npp *npp, int m, int n, int dir, double c0, const int *type, const double *lb,
                       const double *ub, const double *coef, const int *kind, const int *A_ptr,
                       const int *A_ind, const double *A_val, int sol)
{
    try {
        if (!npp || m < 0 || n < 0 || !(dir == 1 || dir == 2) || !(sol == SOL || sol == MIP)) return GLPB_EINVAL;
        if ((m + n > 0 && (!type || !lb || !ub)) || (n > 0 && (!coef || !A_ptr))) return GLPB_EINVAL;
        if (npp->nrows || npp->ncols) return GLPB_EINVAL;
        if (n > 0) {
            if (A_ptr[0] != 0) return GLPB_EINVAL;
            for (int j = 0; j < n; j++) if (A_ptr[j + 1] < A_ptr[j]) return GLPB_EINVAL;
            if (A_ptr[n] > 0 && (!A_ind || !A_val)) return GLPB_EINVAL;
            for (int e = 0; e < A_ptr[n]; e++) if (A_ind[e] < 0 || A_ind[e] >= m) return GLPB_EINVAL;
        }
        for (int k = 0; k < m + n; k++) if (type[k] < FR || type[k] > FX) return GLPB_EINVAL;
        const double sgn = dir == 1 ? +1.0 : -1.0;
        npp->orig_dir = dir; npp->orig_m = m; npp->orig_n = n; npp->orig_nnz = n > 0 ? A_ptr[n] : 0;
        npp->c0 = sgn * c0;
        npp->sol = sol;
        npp->row.reserve(m + 1); npp->col.reserve(n + 1); npp->el.reserve(npp->orig_nnz + 1);
        auto bounds = [&](int k, double &l, double &u) {
            switch (type[k]) {
            case FR: l = -INF; u = +INF; break;
            case LO: l = lb[k]; u = +INF; break;
            case UP: l = -INF; u = ub[k]; break;
            case DB: l = lb[k]; u = ub[k]; break;
            default: l = u = lb[k]; break;
            }
        };
        for (int i = 0; i < m; i++) {
            const int r = npp->add_row();
            bounds(i, npp->row[r].lb, npp->row[r].ub);
        }
        for (int j = 0; j < n; j++) {
            const int c = npp->add_col();
            if (sol == MIP) npp->col[c].is_int = (kind && kind[j] == IV) ? 1 : 0;
            bounds(m + j, npp->col[c].lb, npp->col[c].ub);
            npp->col[c].coef = sgn * coef[j];
            for (int e = A_ptr[j]; e < A_ptr[j + 1]; e++) npp->add_aij(A_ind[e] + 1, c, A_val[e]);
        }
        return 0;
    } catch (const std::bad_alloc &) {
        return GLPB_ENOMEM;
    }
}

int glpb_npp_simplex(glpb_npp *npp)
{
    try {
        if (!npp || npp->sol != SOL || npp->built) return GLPB_EINVAL;
        return npp->process_prob(0);
    } catch (const std::bad_alloc &) {
        return GLPB_ENOMEM;
    }
}

int glpb_npp_integer(glpb_npp *npp, int binarize)
{
    try {
        if (!npp || npp->sol != MIP || npp->built) return GLPB_EINVAL;
        return npp->integer(binarize);
    } catch (const std::bad_alloc &) {
        return GLPB_ENOMEM;
    }
}

int glpb_npp_get_counts(glpb_npp *npp, int *out, int count)
{
    if (!npp || !out) return GLPB_EINVAL;
    int v[8 + K_BINARIZE + 1] = {npp->n_packing, npp->n_covering, npp->n_reduced, npp->bin_vars, npp->bin_bins,
                                 npp->bin_rows, npp->bin_fails, (int)npp->stack.size()};
    for (const Tse &t : npp->stack) v[8 + t.kind]++;
    for (int k = 0; k < count && k < 8 + K_BINARIZE + 1; k++) out[k] = v[k];
    return 0;
}

int glpb_npp_get_size(glpb_npp *npp, int *m, int *n, int *nnz)
{
    if (!npp) return GLPB_EINVAL;
    int mm = 0, nn = 0, nz = 0;
    for (int r = npp->r_head; r; r = npp->row[r].next) mm++;
    for (int c = npp->c_head; c; c = npp->col[c].next) {
        nn++;
        for (int a = npp->col[c].ptr; a; a = npp->el[a].c_next) nz++;
    }
    if (m) *m = mm;
    if (n) *n = nn;
    if (nnz) *nnz = nz;
    return 0;
}

int glpb_npp_build_prob(glpb_npp *npp, double *c0, int *type, double *lb, double *ub, double *coef, int *kind,
                        int *A_ptr, int *A_ind, double *A_val, int *row_ref, int *col_ref)
{
    try {
        if (!npp || npp->built || !c0 || !A_ptr) return GLPB_EINVAL;
        const double sgn = npp->orig_dir == 1 ? +1.0 : -1.0;
        *c0 = sgn * npp->c0;
        auto kind_of = [](double l, double u) {
            if (l == -INF && u == +INF) return FR;
            if (u == +INF) return LO;
            if (l == -INF) return UP;
            if (l != u) return DB;
            return FX;
        };
        int m = 0, n = 0, nz = 0;
        for (int r = npp->r_head; r; r = npp->row[r].next) m++;
        npp->row_ref.assign(1, 0);
        npp->col_ref.assign(1, 0);
        int i = 0;
        for (int r = npp->r_head; r; r = npp->row[r].next, i++) {
            Row &x = npp->row[r];
            x.temp = i;
            type[i] = kind_of(x.lb, x.ub); lb[i] = x.lb; ub[i] = x.ub;
            npp->row_ref.push_back(r);
            if (row_ref) row_ref[i] = r;
        }
        A_ptr[0] = 0;
        for (int c = npp->c_head; c; c = npp->col[c].next, n++) {
            const Col &y = npp->col[c];
            type[m + n] = kind_of(y.lb, y.ub); lb[m + n] = y.lb; ub[m + n] = y.ub;
            coef[n] = sgn * y.coef;
            if (kind) kind[n] = y.is_int ? IV : 1;
            for (int a = y.ptr; a; a = npp->el[a].c_next, nz++) {
                A_ind[nz] = npp->row[npp->el[a].row].temp;
                A_val[nz] = npp->el[a].val;
            }
            A_ptr[n + 1] = nz;
            npp->col_ref.push_back(c);
            if (col_ref) col_ref[n] = c;
        }
        npp->m = m; npp->n = n; npp->nnz = nz;
        npp->built = true;
        npp->c0 = 0.0;
        npp->r_head = npp->r_tail = npp->c_head = npp->c_tail = 0;
        return 0;
    } catch (const std::bad_alloc &) {
        return GLPB_ENOMEM;
    }
}

int glpb_npp_postprocess(glpb_npp *npp, const int *r_stat, const double *r_dual, const int *c_stat,
                         const double *c_value, int *out_r_stat, double *out_r_dual, int *out_c_stat,
                         double *out_c_value)
{
    try {
        if (!npp || !npp->built) return GLPB_EINVAL;
        if (npp->n > 0 && !c_value) return GLPB_EINVAL;
        const bool basic = npp->sol == SOL;
        if (basic && ((npp->m > 0 && (!r_stat || !r_dual)) || (npp->n > 0 && !c_stat))) return GLPB_EINVAL;
        const double sgn = npp->orig_dir == 1 ? +1.0 : -1.0;
        if (basic) {
            npp->r_stat.assign(npp->nrows + 1, 0);
            npp->c_stat.assign(npp->ncols + 1, 0);
            npp->r_pi.assign(npp->nrows + 1, DBL_MAX);
        }
        npp->c_value.assign(npp->ncols + 1, DBL_MAX);
        for (int i = 1; i <= npp->m; i++)
            if (basic) {
                const int k = npp->row_ref[i];
                npp->r_stat[k] = (signed char)r_stat[i - 1];
                npp->r_pi[k] = sgn * r_dual[i - 1];
            }
        for (int j = 1; j <= npp->n; j++) {
            const int k = npp->col_ref[j];
            if (basic) npp->c_stat[k] = (signed char)c_stat[j - 1];
            npp->c_value[k] = c_value[j - 1];
        }
        for (size_t t = npp->stack.size(); t-- > 0;)
            if (npp->recover(npp->stack[t]) != 0) return GLPB_ESTATE;
        for (int i = 1; i <= npp->orig_m && basic; i++) {
            if (out_r_stat) out_r_stat[i - 1] = npp->r_stat[i];
            if (out_r_dual) out_r_dual[i - 1] = sgn * npp->r_pi[i];
        }
        for (int j = 1; j <= npp->orig_n; j++) {
            if (basic && out_c_stat) out_c_stat[j - 1] = npp->c_stat[j];
            if (out_c_value) out_c_value[j - 1] = npp->c_value[j];
        }
        return 0;
    } catch (const std::bad_alloc &) {
        return GLPB_ENOMEM;
    }
}

} // extern "C"
