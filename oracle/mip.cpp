/* mip.cpp -- CPU ORACLE (test infrastructure only; see glpo.h).
 * Restatement of the branch-and-bound path of glp_intopt with presolve OFF and
 * the default options (no cuts, no heuristics, no callback):
 *   solve_mip            lib/glpapi09.js:62-114
 *   ios_driver           lib/glpios03.js:507-951
 *   tree / node LP       lib/glpios01.js
 *   node preprocessing   lib/glpios02.js
 *   branching            lib/glpios09.js:28-270 (FFV/LFV/MFV/DTH)
 *   node selection       lib/glpios12.js
 *   tableau row, dual ratio test  lib/glpapi12.js:401-453,687-762
 */
#include "glpo.h"
#include <algorithm>
#include <cassert>
#include <cmath>
#include <memory>

namespace glpo {

namespace {

struct BndChange { int k, type; double lb, ub; };
struct StatChange { int k, stat; };

struct Node {
    int p = 0;
    Node *up = nullptr;
    int level = 0, count = 0;
    std::vector<BndChange> b;
    std::vector<StatChange> s;
    double lp_obj = 0, bound = 0;
    int ii_cnt = 0;
    double ii_sum = 0;
    Node *prev = nullptr, *next = nullptr, *temp = nullptr;
    bool alive = true;
};

enum { NO_BRNCH = 0, DN_BRNCH = 1, UP_BRNCH = 2 };

struct Tree {
    Prob &mip;
    const IOCP &parm;
    int m, n;
    std::vector<std::unique_ptr<Node>> store;   /* slot p-1 */
    std::vector<int> avail;
    Node *head = nullptr, *tail = nullptr, *curr = nullptr;
    int a_cnt = 0, n_cnt = 0, t_cnt = 0;
    std::vector<int> orig_type, orig_stat, root_type, root_stat, pred_type, pred_stat;
    std::vector<double> orig_lb, orig_ub, orig_prim, orig_dual, root_lb, root_ub, pred_lb, pred_ub;
    double orig_obj = 0;
    bool root_frozen = false;
    std::vector<char> non_int;
    int child = 0;
    long solved = 0;
    double tm_beg;

    Tree(Prob &P, const IOCP &pr) : mip(P), parm(pr), m(P.m), n(P.n) {}

    /* unified access to row/column attributes by k = 1..m+n */
    int &type(int k) { return k <= m ? mip.r_type[k] : mip.c_type[k - m]; }
    double &lb(int k) { return k <= m ? mip.r_lb[k] : mip.c_lb[k - m]; }
    double &ub(int k) { return k <= m ? mip.r_ub[k] : mip.c_ub[k - m]; }
    int &stat(int k) { return k <= m ? mip.r_stat[k] : mip.c_stat[k - m]; }
    double &prim(int k) { return k <= m ? mip.r_prim[k] : mip.c_prim[k - m]; }
    double &dual(int k) { return k <= m ? mip.r_dual[k] : mip.c_dual[k - m]; }
    void set_bnds(int k, int t, double l, double u)
    {
        if (k <= m) prob_set_row_bnds(mip, k, t, l, u); else prob_set_col_bnds(mip, k - m, t, l, u);
    }
    void set_stat(int k, int s)
    {
        if (k <= m) prob_set_row_stat(mip, k, s); else prob_set_col_stat(mip, k - m, s);
    }
    Node *node(int p) { return store[p - 1].get(); }

    /* lib/glpios01.js:1-50 new_node */
    Node *new_node(Node *parent)
    {
        int p;
        if (!avail.empty()) { p = avail.back(); avail.pop_back(); store[p - 1].reset(new Node()); }
        else { store.emplace_back(new Node()); p = (int)store.size(); }
        Node *nd = node(p);
        nd->p = p; nd->up = parent;
        nd->level = parent ? parent->level + 1 : 0;
        nd->lp_obj = parent ? parent->lp_obj : (mip.dir == GLP_MIN ? -DBL_MAX : +DBL_MAX);
        nd->bound = parent ? parent->bound : (mip.dir == GLP_MIN ? -DBL_MAX : +DBL_MAX);
        nd->prev = tail; nd->next = nullptr;
        if (!head) head = nd; else tail->next = nd;
        tail = nd;
        a_cnt++; n_cnt++; t_cnt++;
        if (parent) parent->count++;
        return nd;
    }

    void unlink(Node *nd)
    {
        if (!nd->prev) head = nd->next; else nd->prev->next = nd->next;
        if (!nd->next) tail = nd->prev; else nd->next->prev = nd->prev;
        nd->prev = nd->next = nullptr;
        a_cnt--;
    }

    /* lib/glpios01.js:83-174 ios_create_tree */
    void create()
    {
        orig_type.assign(1 + m + n, 0); orig_stat = orig_type;
        orig_lb.assign(1 + m + n, 0.0); orig_ub = orig_prim = orig_dual = orig_lb;
        for (int k = 1; k <= m + n; k++) {
            orig_type[k] = type(k); orig_lb[k] = lb(k); orig_ub[k] = ub(k);
            orig_stat[k] = stat(k); orig_prim[k] = prim(k); orig_dual[k] = dual(k);
        }
        orig_obj = mip.obj_val;
        non_int.assign(1 + n, 0);
        tm_beg = xtime_ms();
        new_node(nullptr);
    }

    /* lib/glpios01.js:176-308 ios_revive_node */
    void revive(int p)
    {
        Node *nd = node(p);
        assert(nd->count == 0 && curr == nullptr);
        curr = nd;
        Node *root = node(1);
        if (nd == root) return;
        nd->temp = nullptr;
        for (Node *t = nd; t; t = t->up) if (t->up) t->up->temp = t;
        for (Node *t = root; t; t = t->temp) {
            if (t->temp == nullptr) {
                pred_type.assign(1 + m + n, 0); pred_stat = pred_type;
                pred_lb.assign(1 + m + n, 0.0); pred_ub = pred_lb;
                for (int k = 1; k <= m + n; k++) {
                    pred_type[k] = type(k); pred_lb[k] = lb(k); pred_ub[k] = ub(k); pred_stat[k] = stat(k);
                }
            }
            for (const BndChange &b : t->b) set_bnds(b.k, b.type, b.lb, b.ub);
            for (const StatChange &s : t->s) set_stat(s.k, s.stat);
        }
        curr->b.clear();
        curr->s.clear();
    }

    /* lib/glpios01.js:310-464 ios_freeze_node */
    void freeze()
    {
        Node *nd = curr;
        assert(nd);
        if (nd->up == nullptr) {
            assert(!root_frozen);
            root_frozen = true;
            root_type.assign(1 + m + n, 0); root_stat = root_type;
            root_lb.assign(1 + m + n, 0.0); root_ub = root_lb;
            for (int k = 1; k <= m + n; k++) {
                root_type[k] = type(k); root_lb[k] = lb(k); root_ub[k] = ub(k); root_stat[k] = stat(k);
            }
        } else {
            /* the reference prepends to its lists, so they end up in descending k */
            for (int k = m + n; k >= 1; k--) {
                if (!(pred_type[k] == type(k) && pred_lb[k] == lb(k) && pred_ub[k] == ub(k)))
                    nd->b.push_back(BndChange{k, type(k), lb(k), ub(k)});
                if (pred_stat[k] != stat(k)) nd->s.push_back(StatChange{k, stat(k)});
            }
            for (int k = 1; k <= m + n; k++) {
                set_bnds(k, root_type[k], root_lb[k], root_ub[k]);
                set_stat(k, root_stat[k]);
            }
        }
        curr = nullptr;
    }

    /* lib/glpios01.js:466-487 ios_clone_node */
    void clone(int p, int nnn, int *ref)
    {
        Node *nd = node(p);
        assert(nd->count == 0 && curr != nd);
        unlink(nd);
        for (int k = 1; k <= nnn; k++) ref[k] = new_node(nd)->p;
    }

    /* lib/glpios01.js:489-570 ios_delete_node */
    void del(int p)
    {
        Node *nd = node(p);
        assert(nd->count == 0 && curr != nd);
        unlink(nd);
        for (;;) {
            Node *up = nd->up;
            avail.push_back(nd->p);
            nd->alive = false;
            n_cnt--;
            nd = up;
            if (nd) {
                nd->count--;
                if (nd->count == 0) continue;
            }
            break;
        }
    }

    /* lib/glpios01.js:572-613 ios_delete_tree */
    void destroy()
    {
        for (int k = 1; k <= m + n; k++) {
            set_bnds(k, orig_type[k], orig_lb[k], orig_ub[k]);
            set_stat(k, orig_stat[k]);
            prim(k) = orig_prim[k]; dual(k) = orig_dual[k];
        }
        mip.pbs_stat = mip.dbs_stat = GLP_FEAS;
        mip.obj_val = orig_obj;
    }

    /* lib/glpios01.js:789-819 ios_is_hopeful */
    bool is_hopeful(double bound)
    {
        if (mip.mip_stat == GLP_FEAS) {
            double eps = parm.tol_obj * (1.0 + fabs(mip.mip_obj));
            if (mip.dir == GLP_MIN) { if (bound >= mip.mip_obj - eps) return false; }
            else { if (bound <= mip.mip_obj + eps) return false; }
        } else {
            if (mip.dir == GLP_MIN) { if (bound == +DBL_MAX) return false; }
            else { if (bound == -DBL_MAX) return false; }
        }
        return true;
    }

    static int gcd2(int x, int y) { while (y > 0) { int r = x % y; x = y; y = r; } return x; }

    /* lib/glpios01.js:730-787 ios_round_bound (gcdn: lib/glplib03.js:1-24) */
    double round_bound(double bound)
    {
        std::vector<int> c;
        double s = mip.c0;
        int d = 0;
        for (int j = 1; j <= n; j++) {
            double cf = mip.c_coef[j];
            if (cf == 0.0) continue;
            if (mip.c_type[j] == GLP_FX) s += cf * mip.c_prim[j];
            else {
                if (mip.c_kind[j] != GLP_IV) return bound;
                if (cf != floor(cf)) return bound;
                if (fabs(cf) <= (double)INT_MAX) c.push_back((int)fabs(cf)); else d = 1;
            }
        }
        if (d == 0) {
            if (c.empty()) return bound;
            for (int x : c) d = gcd2(d, x);
        }
        assert(d > 0);
        if (mip.dir == GLP_MIN) {
            if (bound != +DBL_MAX) {
                double h = (bound - s) / d;
                if (h >= floor(h) + 0.001) { h = ceil(h); bound = d * h + s; }
            }
        } else {
            if (bound != -DBL_MAX) {
                double h = (bound - s) / d;
                if (h <= ceil(h) - 0.001) { h = floor(h); bound = d * h + s; }
            }
        }
        return bound;
    }

    /* lib/glpapi12.js:401-453 glp_eval_tab_row (with the scaling wrappers of
       glp_btran, lib/glpapi12.js:222-244) */
    int eval_tab_row(int k, std::vector<int> &ind, std::vector<double> &val)
    {
        int i = (k <= m ? mip.r_bind[k] : mip.c_bind[k - m]);
        assert(1 <= i && i <= m && mip.valid);
        std::vector<double> rho(1 + m, 0.0);
        rho[i] = 1.0;
        for (int t = 1; t <= m; t++) {
            int kk = mip.head[t];
            if (kk <= m) rho[t] /= mip.r_rii[kk]; else rho[t] *= mip.c_sjj[kk - m];
        }
        bfd_btran(*mip.bfd, rho.data());
        for (int t = 1; t <= m; t++) rho[t] *= mip.r_rii[t];
        int len = 0;
        for (int kk = 1; kk <= m + n; kk++) {
            double alfa;
            if (kk <= m) {
                if (mip.r_stat[kk] == GLP_BS) continue;
                alfa = -rho[kk];
            } else {
                if (mip.c_stat[kk - m] == GLP_BS) continue;
                alfa = 0.0;
                for (const Elem &e : mip.col_list[kk - m]) alfa += rho[e.idx] * e.val;
            }
            if (alfa != 0.0) { len++; ind[len] = kk; val[len] = alfa; }
        }
        return len;
    }

    /* lib/glpapi12.js:687-762 glp_dual_rtest; returns position in the list */
    int dual_rtest(int len, const std::vector<int> &ind, const std::vector<double> &val, int dir, double eps)
    {
        double obj = (mip.dir == GLP_MIN ? +1.0 : -1.0);
        int piv = 0;
        double teta = DBL_MAX, big = 0.0;
        for (int t = 1; t <= len; t++) {
            int k = ind[t];
            int st = stat(k);
            double cost = dual(k);
            double alfa = (dir > 0 ? +val[t] : -val[t]), temp;
            if (st == GLP_NL) { if (alfa < +eps) continue; temp = (obj * cost) / alfa; }
            else if (st == GLP_NU) { if (alfa > -eps) continue; temp = (obj * cost) / alfa; }
            else if (st == GLP_NF) { if (-eps < alfa && alfa < +eps) continue; temp = 0.0; }
            else continue;
            if (temp < 0.0) temp = 0.0;
            if (teta > temp || (teta == temp && big < fabs(alfa))) { piv = t; teta = temp; big = fabs(alfa); }
        }
        return piv;
    }

    /* lib/glpios01.js:615-728 ios_eval_degrad */
    void eval_degrad(int j, double &dn, double &up)
    {
        std::vector<int> ind(1 + n);
        std::vector<double> val(1 + n);
        double beta = mip.c_prim[j];
        int len = eval_tab_row(m + j, ind, val);
        for (int kase = -1; kase <= +1; kase += 2) {
            int t = dual_rtest(len, ind, val, kase, 1e-9);
            if (t == 0) {
                double inf = (mip.dir == GLP_MIN ? +DBL_MAX : -DBL_MAX);
                if (kase < 0) dn = inf; else up = inf;
                continue;
            }
            int k = ind[t];
            double alfa = val[t];
            int st = stat(k);
            double gamma = dual(k);
            if (mip.dir == GLP_MIN) {
                if ((st == GLP_NL && gamma < 0.0) || (st == GLP_NU && gamma > 0.0) || st == GLP_NF) gamma = 0.0;
            } else {
                if ((st == GLP_NL && gamma > 0.0) || (st == GLP_NU && gamma < 0.0) || st == GLP_NF) gamma = 0.0;
            }
            double delta = (kase < 0 ? floor(beta) : ceil(beta)) - beta;
            delta /= alfa;
            double dz = gamma * delta;
            if (kase < 0) dn = mip.obj_val + dz; else up = mip.obj_val + dz;
        }
    }

    /* lib/glpios01.js:866-910 ios_solve_node */
    int solve_node()
    {
        SMCP sp;
        sp.msg_lev = GLP_MSG_OFF;
        sp.meth = GLP_DUALP;
        if (mip.mip_stat == GLP_FEAS) {
            if (mip.dir == GLP_MIN) sp.obj_ul = mip.mip_obj; else sp.obj_ll = mip.mip_obj;
        }
        solved++;
        return simplex(mip, sp, nullptr);
    }

    /* lib/glpios03.js:56-116 check_integrality */
    void check_integrality()
    {
        int ii_cnt = 0;
        double ii_sum = 0.0;
        for (int j = 1; j <= n; j++) {
            non_int[j] = 0;
            if (mip.c_kind[j] != GLP_IV) continue;
            if (mip.c_stat[j] != GLP_BS) continue;
            int t = mip.c_type[j];
            double l = mip.c_lb[j], u = mip.c_ub[j], x = mip.c_prim[j];
            if (t == GLP_LO || t == GLP_DB || t == GLP_FX) {
                if (l - parm.tol_int <= x && x <= l + parm.tol_int) continue;
                if (x < l) continue;
            }
            if (t == GLP_UP || t == GLP_DB || t == GLP_FX) {
                if (u - parm.tol_int <= x && x <= u + parm.tol_int) continue;
                if (x > u) continue;
            }
            double r = floor(x + 0.5);
            if (r - parm.tol_int <= x && x <= r + parm.tol_int) continue;
            non_int[j] = 1;
            ii_cnt++;
            double t1 = x - floor(x), t2 = ceil(x) - x;
            ii_sum += (t1 <= t2 ? t1 : t2);
        }
        curr->ii_cnt = ii_cnt;
        curr->ii_sum = ii_sum;
    }

    /* lib/glpios03.js:118-139 record_solution */
    void record_solution()
    {
        mip.mip_stat = GLP_FEAS;
        mip.mip_obj = mip.obj_val;
        for (int i = 1; i <= m; i++) mip.r_mipx[i] = mip.r_prim[i];
        for (int j = 1; j <= n; j++)
            mip.c_mipx[j] = (mip.c_kind[j] == GLP_IV) ? floor(mip.c_prim[j] + 0.5) : mip.c_prim[j];
    }

    /* lib/glpios03.js:307-377 fix_by_red_cost */
    void fix_by_red_cost()
    {
        double obj = mip.obj_val;
        for (int j = 1; j <= n; j++) {
            if (mip.c_kind[j] != GLP_IV) continue;
            double l = mip.c_lb[j], u = mip.c_ub[j], dj = mip.c_dual[j];
            int st = mip.c_stat[j];
            if (mip.dir == GLP_MIN) {
                if (st == GLP_NL) { if (dj < 0.0) dj = 0.0; if (obj + dj >= mip.mip_obj) prob_set_col_bnds(mip, j, GLP_FX, l, l); }
                else if (st == GLP_NU) { if (dj > 0.0) dj = 0.0; if (obj - dj >= mip.mip_obj) prob_set_col_bnds(mip, j, GLP_FX, u, u); }
            } else {
                if (st == GLP_NL) { if (dj > 0.0) dj = 0.0; if (obj + dj <= mip.mip_obj) prob_set_col_bnds(mip, j, GLP_FX, l, l); }
                else if (st == GLP_NU) { if (dj < 0.0) dj = 0.0; if (obj - dj <= mip.mip_obj) prob_set_col_bnds(mip, j, GLP_FX, u, u); }
            }
        }
    }

    /* lib/glpios09.js:60-82 branch_mostf */
    int branch_mostf(int &next)
    {
        int jj = 0;
        double most = DBL_MAX;
        for (int j = 1; j <= n; j++)
            if (non_int[j]) {
                double beta = mip.c_prim[j], temp = floor(beta) + 0.5;
                if (most > fabs(beta - temp)) {
                    jj = j; most = fabs(beta - temp);
                    next = (beta < temp) ? DN_BRNCH : UP_BRNCH;
                }
            }
        return jj;
    }

    /* lib/glpios09.js:84-270 branch_drtom (Driebeck-Tomlin) */
    int branch_drtom(int &next)
    {
        std::vector<int> ind(1 + n);
        std::vector<double> val(1 + n);
        int jj = 0;
        double degrad = -1.0, dz_dn = 0, dz_up = 0;
        for (int j = 1; j <= n; j++) {
            if (!non_int[j]) continue;
            double x = mip.c_prim[j];
            int len = eval_tab_row(m + j, ind, val);
            for (int kase = -1; kase <= +1; kase += 2) {
                int t = dual_rtest(len, ind, val, kase, 1e-9);
                double delta_z;
                if (t == 0) delta_z = (mip.dir == GLP_MIN ? +DBL_MAX : -DBL_MAX);
                else {
                    int k = ind[t];
                    double alfa = val[t];
                    double delta_j = (kase < 0 ? floor(x) : ceil(x)) - x;
                    double delta_k = delta_j / alfa;
                    if (k > m && mip.c_kind[k - m] != GLP_CV) {
                        if (fabs(delta_k - floor(delta_k + 0.5)) > 1e-3)
                            delta_k = (delta_k > 0.0) ? ceil(delta_k) : floor(delta_k);
                    }
                    int st = stat(k);
                    double dk = dual(k);
                    if (mip.dir == GLP_MIN) {
                        if ((st == GLP_NL && dk < 0.0) || (st == GLP_NU && dk > 0.0) || st == GLP_NF) dk = 0.0;
                    } else {
                        if ((st == GLP_NL && dk > 0.0) || (st == GLP_NU && dk < 0.0) || st == GLP_NF) dk = 0.0;
                    }
                    delta_z = dk * delta_k;
                }
                if (kase < 0) dz_dn = delta_z; else dz_up = delta_z;
            }
            if (degrad < fabs(dz_dn) || degrad < fabs(dz_up)) {
                jj = j;
                if (fabs(dz_dn) < fabs(dz_up)) { next = DN_BRNCH; degrad = fabs(dz_up); }
                else { next = UP_BRNCH; degrad = fabs(dz_dn); }
                if (degrad == DBL_MAX) break;
            }
        }
        assert(1 <= jj && jj <= n);
        if (degrad < 1e-6 * (1.0 + 0.001 * fabs(mip.obj_val))) jj = branch_mostf(next);
        return jj;
    }

    /* lib/glpios09.js:1-26 ios_choose_var */
    int choose_var(int &next)
    {
        if (parm.br_tech == GLP_BR_FFV || parm.br_tech == GLP_BR_LFV) {
            int j;
            if (parm.br_tech == GLP_BR_FFV) { for (j = 1; j <= n; j++) if (non_int[j]) break; }
            else { for (j = n; j >= 1; j--) if (non_int[j]) break; }
            double beta = mip.c_prim[j];
            next = (beta - floor(beta) < ceil(beta) - beta) ? DN_BRNCH : UP_BRNCH;
            return j;
        }
        if (parm.br_tech == GLP_BR_MFV) return branch_mostf(next);
        return branch_drtom(next);
    }

    /* lib/glpios12.js ios_choose_node */
    int choose_node()
    {
        if (parm.bt_tech == GLP_BT_DFS) return tail->p;
        if (parm.bt_tech == GLP_BT_BFS) return head->p;
        if (parm.bt_tech == GLP_BT_BLB) {
            Node *best = nullptr;
            if (mip.dir == GLP_MIN) {
                double bound = +DBL_MAX;
                for (Node *t = head; t; t = t->next) if (bound > t->bound) bound = t->bound;
                double eps = 0.001 * (1.0 + fabs(bound));
                for (Node *t = head; t; t = t->next)
                    if (t->bound <= bound + eps)
                        if (!best || best->up->ii_sum > t->up->ii_sum) best = t;
            } else {
                double bound = -DBL_MAX;
                for (Node *t = head; t; t = t->next) if (bound < t->bound) bound = t->bound;
                double eps = 0.001 * (1.0 + fabs(bound));
                for (Node *t = head; t; t = t->next)
                    if (t->bound >= bound - eps)
                        if (!best || best->lp_obj < t->lp_obj) best = t;
            }
            return best->p;
        }
        /* GLP_BT_BPH */
        int p = 0;
        double best = DBL_MAX;
        if (mip.mip_stat == GLP_UNDEF) {
            for (Node *t = head; t; t = t->next)
                if (best > t->up->ii_sum) { p = t->p; best = t->up->ii_sum; }
        } else {
            Node *root = node(1);
            double deg = (mip.mip_obj - root->bound) / root->ii_sum;
            for (Node *t = head; t; t = t->next) {
                double obj = t->up->bound + deg * t->up->ii_sum;
                if (mip.dir == GLP_MAX) obj = -obj;
                if (best > obj) { p = t->p; best = obj; }
            }
        }
        return p;
    }

    /* ---- lib/glpios02.js ios_preprocess_node ---- */
    struct RowInfo { double f_min, f_max; int j_min, j_max; };

    static void prepare_row_info(int len, const double *a, const double *l, const double *u, RowInfo &f)
    {
        f.f_min = 0.0; f.j_min = 0;
        for (int j = 1; j <= len; j++) {
            if (a[j] > 0.0) {
                if (l[j] == -DBL_MAX) { if (f.j_min == 0) f.j_min = j; else { f.f_min = -DBL_MAX; f.j_min = 0; break; } }
                else f.f_min += a[j] * l[j];
            } else {
                if (u[j] == +DBL_MAX) { if (f.j_min == 0) f.j_min = j; else { f.f_min = -DBL_MAX; f.j_min = 0; break; } }
                else f.f_min += a[j] * u[j];
            }
        }
        f.f_max = 0.0; f.j_max = 0;
        for (int j = 1; j <= len; j++) {
            if (a[j] > 0.0) {
                if (u[j] == +DBL_MAX) { if (f.j_max == 0) f.j_max = j; else { f.f_max = +DBL_MAX; f.j_max = 0; break; } }
                else f.f_max += a[j] * u[j];
            } else {
                if (l[j] == -DBL_MAX) { if (f.j_max == 0) f.j_max = j; else { f.f_max = +DBL_MAX; f.j_max = 0; break; } }
                else f.f_max += a[j] * l[j];
            }
        }
    }

    static int check_row_bounds(const RowInfo &f, double &L, double &U)
    {
        double LL = (f.j_min == 0 ? f.f_min : -DBL_MAX), UU = (f.j_max == 0 ? f.f_max : +DBL_MAX);
        if (L != -DBL_MAX) { double eps = 1e-3 * (1.0 + fabs(L)); if (UU < L - eps) return 1; }
        if (U != +DBL_MAX) { double eps = 1e-3 * (1.0 + fabs(U)); if (LL > U + eps) return 1; }
        if (L != -DBL_MAX) { double eps = 1e-12 * (1.0 + fabs(L)); if (LL > L - eps) L = -DBL_MAX; }
        if (U != +DBL_MAX) { double eps = 1e-12 * (1.0 + fabs(U)); if (UU < U + eps) U = +DBL_MAX; }
        return 0;
    }

    static void col_implied_bounds(const RowInfo &f, const double *a, double L, double U, const double *l,
                                   const double *u, int k, double &ll, double &uu)
    {
        double ilb, iub;
        if (L == -DBL_MAX || f.f_max == +DBL_MAX) ilb = -DBL_MAX;
        else if (f.j_max == 0) ilb = L - (f.f_max - a[k] * (a[k] > 0.0 ? u[k] : l[k]));
        else if (f.j_max == k) ilb = L - f.f_max;
        else ilb = -DBL_MAX;
        if (U == +DBL_MAX || f.f_min == -DBL_MAX) iub = +DBL_MAX;
        else if (f.j_min == 0) iub = U - (f.f_min - a[k] * (a[k] > 0.0 ? l[k] : u[k]));
        else if (f.j_min == k) iub = U - f.f_min;
        else iub = +DBL_MAX;
        if (fabs(a[k]) < 1e-6) { ll = -DBL_MAX; uu = +DBL_MAX; }
        else if (a[k] > 0.0) { ll = (ilb == -DBL_MAX ? -DBL_MAX : ilb / a[k]); uu = (iub == +DBL_MAX ? +DBL_MAX : iub / a[k]); }
        else { ll = (iub == +DBL_MAX ? -DBL_MAX : iub / a[k]); uu = (ilb == -DBL_MAX ? +DBL_MAX : ilb / a[k]); }
    }

    static int check_col_bounds(const RowInfo &f, const double *a, double L, double U, const double *l,
                                const double *u, int flag, int j, double &lj_out, double &uj_out)
    {
        double lj = l[j], uj = u[j], ll, uu;
        col_implied_bounds(f, a, L, U, l, u, j, ll, uu);
        if (flag) {
            if (ll != -DBL_MAX) ll = (ll - floor(ll) < 1e-3 ? floor(ll) : ceil(ll));
            if (uu != +DBL_MAX) uu = (ceil(uu) - uu < 1e-3 ? ceil(uu) : floor(uu));
        }
        if (lj != -DBL_MAX) { double eps = 1e-3 * (1.0 + fabs(lj)); if (uu < lj - eps) return 1; }
        if (uj != +DBL_MAX) { double eps = 1e-3 * (1.0 + fabs(uj)); if (ll > uj + eps) return 1; }
        if (ll != -DBL_MAX) { double eps = 1e-3 * (1.0 + fabs(ll)); if (lj < ll - eps) lj = ll; }
        if (uu != +DBL_MAX) { double eps = 1e-3 * (1.0 + fabs(uu)); if (uj > uu + eps) uj = uu; }
        if (!(lj == -DBL_MAX || uj == +DBL_MAX)) {
            double t1 = fabs(lj), t2 = fabs(uj);
            double eps = 1e-10 * (1.0 + (t1 <= t2 ? t1 : t2));
            if (lj > uj - eps) {
                if (lj == l[j]) uj = lj;
                else if (uj == u[j]) lj = uj;
                else if (t1 <= t2) uj = lj;
                else lj = uj;
            }
        }
        lj_out = lj; uj_out = uj;
        return 0;
    }

    static int check_efficiency(int flag, double l, double u, double ll, double uu)
    {
        int eff = 0;
        if (l < ll) {
            if (flag || l == -DBL_MAX) eff++;
            else {
                double r = (u == +DBL_MAX) ? 1.0 + fabs(l) : 1.0 + (u - l);
                if (ll - l >= 0.25 * r) eff++;
            }
        }
        if (u > uu) {
            if (flag || u == +DBL_MAX) eff++;
            else {
                double r = (l == -DBL_MAX) ? 1.0 + fabs(u) : 1.0 + (u - l);
                if (u - uu >= 0.25 * r) eff++;
            }
        }
        return eff;
    }

    int preprocess_node(int max_pass)
    {
        std::vector<double> L(1 + m), U(1 + m), l(1 + n), u(1 + n);
        if (mip.mip_stat == GLP_FEAS) {
            if (mip.dir == GLP_MIN) { L[0] = -DBL_MAX; U[0] = mip.mip_obj - mip.c0; }
            else { L[0] = mip.mip_obj - mip.c0; U[0] = +DBL_MAX; }
        } else { L[0] = -DBL_MAX; U[0] = +DBL_MAX; }
        auto get_lb = [](int t, double v) { return (t == GLP_FR || t == GLP_UP) ? -DBL_MAX : v; };
        auto get_ub = [](int t, double lbv, double ubv) { return (t == GLP_FR || t == GLP_LO) ? +DBL_MAX : (t == GLP_FX ? lbv : ubv); };
        for (int i = 1; i <= m; i++) { L[i] = get_lb(mip.r_type[i], mip.r_lb[i]); U[i] = get_ub(mip.r_type[i], mip.r_lb[i], mip.r_ub[i]); }
        for (int j = 1; j <= n; j++) { l[j] = get_lb(mip.c_type[j], mip.c_lb[j]); u[j] = get_ub(mip.c_type[j], mip.c_lb[j], mip.c_ub[j]); }
        /* basic_preprocessing with all rows 0..m on the list */
        std::vector<int> list(2 + m), mark(2 + m, 0), pass(2 + m, 0), ind(1 + n);
        std::vector<double> val(1 + n), lbv(1 + n), ubv(1 + n);
        int size = 0;
        for (int k = 1; k <= m + 1; k++) { list[++size] = k - 1; mark[k - 1] = 1; }
        while (size > 0) {
            int i = list[size--];
            mark[i] = 0;
            pass[i]++;
            if (L[i] == -DBL_MAX && U[i] == +DBL_MAX) continue;
            int len = 0;
            if (i == 0) {
                for (int j = 1; j <= n; j++) if (mip.c_coef[j] != 0.0) { len++; ind[len] = j; val[len] = mip.c_coef[j]; }
            } else
                for (const Elem &e : mip.row_list[i]) { len++; ind[len] = e.idx; val[len] = e.val; }
            for (int k = 1; k <= len; k++) { lbv[k] = l[ind[k]]; ubv[k] = u[ind[k]]; }
            RowInfo f;
            prepare_row_info(len, val.data(), lbv.data(), ubv.data(), f);
            if (check_row_bounds(f, L[i], U[i])) return 1;
            if (L[i] == -DBL_MAX && U[i] == +DBL_MAX) continue;
            for (int k = 1; k <= len; k++) {
                int j = ind[k];
                int flag = mip.c_kind[j] != GLP_CV;
                double ll, uu;
                if (check_col_bounds(f, val.data(), L[i], U[i], lbv.data(), ubv.data(), flag, k, ll, uu)) return 1;
                int eff = check_efficiency(flag, l[j], u[j], ll, uu);
                l[j] = ll; u[j] = uu;
                if (eff > 0)
                    for (const Elem &e : mip.col_list[j]) {
                        int ii = e.idx;
                        if (pass[ii] >= max_pass) continue;
                        if (L[ii] == -DBL_MAX && U[ii] == +DBL_MAX) continue;
                        if (mark[ii] == 0) { list[++size] = ii; mark[ii] = 1; }
                    }
            }
        }
        for (int i = 1; i <= m; i++)
            if (mip.r_stat[i] == GLP_BS) {
                if (L[i] == -DBL_MAX && U[i] == +DBL_MAX) prob_set_row_bnds(mip, i, GLP_FR, 0.0, 0.0);
                else if (U[i] == +DBL_MAX) prob_set_row_bnds(mip, i, GLP_LO, L[i], 0.0);
                else if (L[i] == -DBL_MAX) prob_set_row_bnds(mip, i, GLP_UP, 0.0, U[i]);
            }
        for (int j = 1; j <= n; j++) {
            int type;
            if (l[j] == -DBL_MAX && u[j] == +DBL_MAX) type = GLP_FR;
            else if (u[j] == +DBL_MAX) type = GLP_LO;
            else if (l[j] == -DBL_MAX) type = GLP_UP;
            else if (l[j] != u[j]) type = GLP_DB;
            else type = GLP_FX;
            prob_set_col_bnds(mip, j, type, l[j], u[j]);
        }
        return 0;
    }

    void improve_bound(Node *nd, double bnd)
    {
        if (mip.dir == GLP_MIN) { if (nd->bound < bnd) nd->bound = bnd; }
        else { if (nd->bound > bnd) nd->bound = bnd; }
    }

    /* lib/glpios03.js:141-305 branch_on */
    int branch_on(int j, int next)
    {
        int type = mip.c_type[j], dn_type, up_type;
        double l = mip.c_lb[j], u = mip.c_ub[j], beta = mip.c_prim[j];
        double new_ub = floor(beta), new_lb = ceil(beta);
        switch (type) {
        case GLP_FR: dn_type = GLP_UP; up_type = GLP_LO; break;
        case GLP_LO: dn_type = (l == new_ub ? GLP_FX : GLP_DB); up_type = GLP_LO; break;
        case GLP_UP: dn_type = GLP_UP; up_type = (new_lb == u ? GLP_FX : GLP_DB); break;
        case GLP_DB: dn_type = (l == new_ub ? GLP_FX : GLP_DB); up_type = (new_lb == u ? GLP_FX : GLP_DB); break;
        default: assert(!"bad type"); return 2;
        }
        double dn_lp = 0, up_lp = 0;
        eval_degrad(j, dn_lp, up_lp);
        double dn_bnd = round_bound(dn_lp), up_bnd = round_bound(up_lp);
        bool dn_bad = !is_hopeful(dn_bnd), up_bad = !is_hopeful(up_bnd);
        if (dn_bad && up_bad) return 2;
        if (up_bad) {
            prob_set_col_bnds(mip, j, dn_type, l, new_ub);
            curr->lp_obj = dn_lp;
            improve_bound(curr, dn_bnd);
            return 1;
        }
        if (dn_bad) {
            prob_set_col_bnds(mip, j, up_type, new_lb, u);
            curr->lp_obj = up_lp;
            improve_bound(curr, up_bnd);
            return 1;
        }
        int p = curr->p;
        freeze();
        int ref[3];
        clone(p, 2, ref);
        Node *nd = node(ref[1]);
        nd->b.push_back(BndChange{m + j, dn_type, l, new_ub});
        nd->lp_obj = dn_lp;
        improve_bound(nd, dn_bnd);
        nd = node(ref[2]);
        nd->b.push_back(BndChange{m + j, up_type, new_lb, u});
        nd->lp_obj = up_lp;
        improve_bound(nd, up_bnd);
        child = (next == NO_BRNCH ? 0 : (next == DN_BRNCH ? ref[1] : ref[2]));
        return 0;
    }

    void cleanup_the_tree()
    {
        Node *next;
        for (Node *nd = head; nd; nd = next) {
            next = nd->next;
            if (!is_hopeful(nd->bound)) del(nd->p);
        }
    }

    /* lib/glpios03.js:507-951 ios_driver (states loop/more/fath) */
    int driver()
    {
        enum { LOOP, MORE, FATH } state = LOOP;
        int p = 0, ret;
        for (;;) {
            if (state == LOOP) {
                if (head == nullptr) return 0;
                int next_p;
                if (a_cnt == 1) next_p = head->p;
                else if (child != 0) next_p = child;
                else next_p = choose_node();
                revive(next_p);
                child = 0;
                p = curr->p;
                state = MORE;
            }
            if (state == MORE) {
                if (parm.node_lim >= 0 && solved >= parm.node_lim) return GLP_ESTOP;
                if (parm.tm_lim < INT_MAX && (parm.tm_lim - 1) <= (xtime_ms() - tm_beg)) return GLP_ETMLIM;
                if (parm.pp_tech == GLP_PP_ROOT) { if (curr->level == 0 && preprocess_node(100)) { state = FATH; continue; } }
                else if (parm.pp_tech == GLP_PP_ALL) { if (preprocess_node(curr->level == 0 ? 100 : 10)) { state = FATH; continue; } }
                if (!is_hopeful(curr->bound)) { state = FATH; continue; }
                ret = solve_node();
                if (!(ret == 0 || ret == GLP_EOBJLL || ret == GLP_EOBJUL)) return GLP_EFAIL;
                int p_stat = mip.pbs_stat, d_stat = mip.dbs_stat;
                if (p_stat == GLP_FEAS && d_stat == GLP_FEAS) {}
                else if (d_stat == GLP_NOFEAS) return GLP_EFAIL;
                else if (p_stat == GLP_INFEAS && d_stat == GLP_FEAS) { state = FATH; continue; }
                else if (p_stat == GLP_NOFEAS) { state = FATH; continue; }
                else assert(!"unexpected node LP status");
                curr->lp_obj = mip.obj_val;
                improve_bound(curr, round_bound(mip.obj_val));
                if (!is_hopeful(curr->bound)) { state = FATH; continue; }
                check_integrality();
                if (curr->ii_cnt == 0) { record_solution(); state = FATH; continue; }
                if (mip.mip_stat == GLP_FEAS) fix_by_red_cost();
                int next = NO_BRNCH;
                int jv = choose_var(next);
                ret = branch_on(jv, next);
                if (ret == 0) { state = LOOP; continue; }
                if (ret == 1) { state = MORE; continue; }
                state = FATH;
            }
            if (state == FATH) {
                freeze();
                del(p);
                if (mip.mip_stat == GLP_FEAS) cleanup_the_tree();
                state = LOOP;
            }
        }
    }
};

} /* namespace */

/* glp_intopt with presolve OFF: lib/glpapi09.js:61-114,256-390 */
int intopt(Prob &P, const IOCP &parm, long *n_nodes)
{
    P.mip_stat = GLP_UNDEF;
    P.mip_obj = 0.0;
    for (int j = 1; j <= P.n; j++)
        if (P.c_kind[j] == GLP_IV) {
            int t = P.c_type[j];
            if ((t == GLP_LO || t == GLP_DB) && P.c_lb[j] != floor(P.c_lb[j])) return GLP_EBOUND;
            if ((t == GLP_UP || t == GLP_DB) && P.c_ub[j] != floor(P.c_ub[j])) return GLP_EBOUND;
            if (t == GLP_FX && P.c_lb[j] != floor(P.c_lb[j])) return GLP_EBOUND;
        }
    if (prob_get_status(P) != GLP_OPT) return GLP_EROOT;
    Tree T(P, parm);
    T.create();
    int ret = T.driver();
    T.destroy();
    if (n_nodes) *n_nodes = T.solved;
    if (ret == 0) P.mip_stat = (P.mip_stat == GLP_FEAS) ? GLP_OPT : GLP_NOFEAS;
    return ret;
}

} /* namespace glpo */
