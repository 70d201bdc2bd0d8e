/* mip.cu -- branch-and-bound driver of glp_intopt (presolve OFF, default
 * options: no cuts, no heuristics, no callback), SURVEY.md 8a row a18.
 *
 * The search tree lives on the host (it is pointer surgery on a few hundred
 * bytes per node); every node LP is a warm-started dual simplex on the device
 * handle (ios_solve_node, lib/glpios01.js:866-910), and the simplex-table rows
 * that Driebeck-Tomlin branching needs (glp_eval_tab_row,
 * lib/glpapi12.js:401-453) are one k_rho + one k_trow launch each.
 *
 * Multi-GPU: the tree is resumable.  glpb_mip_begin / _run / _end
 * let a host run the search in slices; between slices ranks exchange the
 * incumbent objective (glpb_mip_set_cutoff) and migrate open nodes
 * (glpb_mip_export_nodes / _import_nodes, self-contained records), which is
 * all the communication branch-and-bound needs.  glpb_intopt is begin +
 * run-to-completion + end on one GPU.
 */
#include "kernels.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>

namespace {

struct BndChange { int k, type; double lb, ub; };   /* k = 0..m+n-1 */
struct StatChange { int k, stat; };

struct MNode {
    int up = -1, level = 0, count = 0;
    std::vector<BndChange> b;
    std::vector<StatChange> s;
    double lp_obj = 0, bound = 0, ii_sum = 0;
    int ii_cnt = 0;
    int prev = -1, next = -1, temp = -1;
    bool used = false;
};

enum { NO_BRNCH = 0, DN_BRNCH = 1, UP_BRNCH = 2 };
enum { S_LOOP, S_MORE, S_FATH };

} /* namespace */

struct glpb_mip {
    glpb_prob *P;
    glpb_iocp parm;
    int m, n;
    std::vector<MNode> pool;
    std::vector<int> avail;
    int head = -1, tail = -1, curr = -1, child = -1, a_cnt = 0;
    std::vector<int> orig_type, orig_stat, root_type, root_stat, pred_type, pred_stat;
    std::vector<double> orig_lb, orig_ub, orig_prim, orig_dual, root_lb, root_ub, pred_lb, pred_ub;
    double orig_obj = 0;
    bool root_frozen = false;
    std::vector<char> non_int;
    long solved = 0;
    double tm_beg = 0;
    int state = S_LOOP, cur_p = -1;
    /* incumbent: own solution or a cut-off value received from another rank */
    bool have_sol = false;      /* mipx holds a solution found here            */
    bool have_cut = false;      /* P->mip_obj is a valid incumbent objective   */
    std::vector<double> trow;   /* host copy of a simplex-table row            */
    std::vector<int> dev_head;
    /* batched table rows of one node (k_tab_rows) */
    int *d_tab_pos = nullptr;
    double *d_tab_rho = nullptr, *d_tab_trow = nullptr;
    int tab_cap = 0;
    std::vector<int> tab_js;            /* structural j of every cached row, in order */
    std::vector<double> tab_rows_h;     /* [rows][n] scaled table rows                */
    ~glpb_mip()
    {
        if (d_tab_pos) cudaFree(d_tab_pos);
        if (d_tab_rho) cudaFree(d_tab_rho);
        if (d_tab_trow) cudaFree(d_tab_trow);
    }

    explicit glpb_mip(glpb_prob *P_, const glpb_iocp &pr) : P(P_), parm(pr), m(P_->m), n(P_->n) {}

    int dir() const { return P->dir; }
    bool feas() const { return have_cut; }

    /* ---- tree storage: lib/glpios01.js:1-82 ---- */
    int new_node(int parent)
    {
        int p;
        if (!avail.empty()) { p = avail.back(); avail.pop_back(); pool[p] = MNode(); }
        else { pool.emplace_back(); p = (int)pool.size() - 1; }
        MNode &nd = pool[p];
        nd.used = true; nd.up = parent;
        nd.level = parent >= 0 ? pool[parent].level + 1 : 0;
        double inf = (dir() == GLP_MIN ? -DBL_MAX : +DBL_MAX);
        nd.lp_obj = parent >= 0 ? pool[parent].lp_obj : inf;
        nd.bound = parent >= 0 ? pool[parent].bound : inf;
        nd.prev = tail; nd.next = -1;
        if (head < 0) head = p; else pool[tail].next = p;
        tail = p;
        a_cnt++;
        if (parent >= 0) pool[parent].count++;
        return p;
    }

    void unlink(int p)
    {
        MNode &nd = pool[p];
        if (nd.prev < 0) head = nd.next; else pool[nd.prev].next = nd.next;
        if (nd.next < 0) tail = nd.prev; else pool[nd.next].prev = nd.prev;
        nd.prev = nd.next = -1;
        a_cnt--;
    }

    void set_bnds(int k, int t, double l, double u)
    {
        int kk = k + 1;
        glpb_set_bounds(P, 1, &kk, &t, &l, &u);
    }

    /* glp_set_row_stat / glp_set_col_stat for one variable, lib/glpapi05.js:1-47 */
    void set_stat(int k, int s)
    {
        if (s != GLP_BS) {
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

    /* lib/glpios01.js:83-174 ios_create_tree */
    void create()
    {
        orig_type = P->h_type; orig_stat = P->h_stat;
        orig_lb = P->h_lb; orig_ub = P->h_ub; orig_prim = P->h_prim; orig_dual = P->h_dual;
        orig_obj = P->obj_val;
        non_int.assign(n, 0);
        trow.assign(n, 0.0);
        dev_head.assign(m + n, 0);
        tm_beg = now_ms();
        new_node(-1);
    }

    /* lib/glpios01.js:176-308 ios_revive_node */
    void revive(int p)
    {
        curr = p;
        int root = 0;
        if (p == root) return;
        pool[p].temp = -1;
        for (int t = p; t >= 0; t = pool[t].up) if (pool[t].up >= 0) pool[pool[t].up].temp = t;
        for (int t = root; t >= 0; t = pool[t].temp) {
            if (pool[t].temp < 0) {
                pred_type = P->h_type; pred_lb = P->h_lb; pred_ub = P->h_ub; pred_stat = P->h_stat;
            }
            for (const BndChange &b : pool[t].b) set_bnds(b.k, b.type, b.lb, b.ub);
            for (const StatChange &s : pool[t].s) set_stat(s.k, s.stat);
        }
        pool[p].b.clear();
        pool[p].s.clear();
    }

    /* lib/glpios01.js:310-464 ios_freeze_node */
    void freeze()
    {
        MNode &nd = pool[curr];
        if (nd.up < 0 && !root_frozen) {
            root_frozen = true;
            root_type = P->h_type; root_lb = P->h_lb; root_ub = P->h_ub; root_stat = P->h_stat;
        } else if (nd.up >= 0) {
            for (int k = m + n - 1; k >= 0; k--) {
                if (!(pred_type[k] == P->h_type[k] && pred_lb[k] == P->h_lb[k] && pred_ub[k] == P->h_ub[k]))
                    nd.b.push_back(BndChange{k, P->h_type[k], P->h_lb[k], P->h_ub[k]});
                if (pred_stat[k] != P->h_stat[k]) nd.s.push_back(StatChange{k, P->h_stat[k]});
            }
            for (int k = 0; k < m + n; k++) {
                set_bnds(k, root_type[k], root_lb[k], root_ub[k]);
                set_stat(k, root_stat[k]);
            }
        }
        curr = -1;
    }

    /* lib/glpios01.js:489-570 ios_delete_node */
    void del(int p)
    {
        unlink(p);
        for (;;) {
            int up = pool[p].up;
            pool[p].used = false;
            pool[p].b.clear(); pool[p].s.clear();
            avail.push_back(p);
            p = up;
            if (p >= 0) {
                pool[p].count--;
                if (pool[p].count == 0 && p != 0) continue;
                if (pool[p].count == 0 && p == 0) { pool[p].used = false; }
            }
            break;
        }
    }

    /* lib/glpios01.js:572-613 ios_delete_tree */
    void destroy()
    {
        for (int k = 0; k < m + n; k++) {
            set_bnds(k, orig_type[k], orig_lb[k], orig_ub[k]);
            set_stat(k, orig_stat[k]);
        }
        P->h_prim = orig_prim; P->h_dual = orig_dual;
        P->pbs_stat = P->dbs_stat = GLP_FEAS;
        P->obj_val = orig_obj;
    }

    /* lib/glpios01.js:789-819 ios_is_hopeful */
    bool is_hopeful(double bound) const
    {
        if (feas()) {
            double eps = parm.tol_obj * (1.0 + fabs(P->mip_obj));
            if (dir() == GLP_MIN) { if (bound >= P->mip_obj - eps) return false; }
            else { if (bound <= P->mip_obj + eps) return false; }
        } else {
            if (dir() == GLP_MIN) { if (bound == +DBL_MAX) return false; }
            else { if (bound == -DBL_MAX) return false; }
        }
        return true;
    }

    /* lib/glpios01.js:730-787 ios_round_bound */
    double round_bound(double bound) const
    {
        std::vector<long> c;
        double s = P->c0;
        long d = 0;
        for (int j = 0; j < n; j++) {
            double cf = P->h_coef[j];
            if (cf == 0.0) continue;
            if (P->h_type[m + j] == GLP_FX) s += cf * P->h_prim[m + j];
            else {
                if (P->h_kind[j] != GLP_IV) return bound;
                if (cf != floor(cf)) return bound;
                if (fabs(cf) <= (double)INT_MAX) c.push_back((long)fabs(cf)); else d = 1;
            }
        }
        if (d == 0) {
            if (c.empty()) return bound;
            for (long x : c) { long a = d, b2 = x; while (b2 > 0) { long r = a % b2; a = b2; b2 = r; } d = a; }
        }
        if (dir() == GLP_MIN) {
            if (bound != +DBL_MAX) {
                double h = (bound - s) / d;
                if (h >= floor(h) + 0.001) bound = d * ceil(h) + s;
            }
        } else {
            if (bound != -DBL_MAX) {
                double h = (bound - s) / d;
                if (h <= ceil(h) - 0.001) bound = d * floor(h) + s;
            }
        }
        return bound;
    }

    /* glp_eval_tab_row (lib/glpapi12.js:401-453) for the basic variable k on
       the device: rho = row of inv(B) (k_rho), then -rho' N_j for every
       non-basic j (k_trow).  Result: alfa[kk] for kk = 0..m+n-1 (0 for basic). */
    int tab_row(int k, std::vector<double> &alfa)
    {
        /* position of variable k in the basis header */
        int pos = -1;
        for (int i = 0; i < m; i++) if (P->h_head[i] - 1 == k) { pos = i; break; }
        if (pos < 0 || !P->valid) return GLPB_ESTATE;
        Dev D(P);
        LAUNCH(P, k_clear_ctrl, 1, 1, 0, P->ctrl, 2);
        CK(cudaMemcpyAsync(&P->ctrl->p, &pos, sizeof(int), cudaMemcpyHostToDevice, P->stream));
        launch_rho(P, P->h_ctrl->k);   /* k_clear_ctrl leaves pse/refct alone: k_trow gets u = NULL anyway */
        GROUP_DISPATCH(D.gc, LAUNCH(P, k_trow<GG>, cdiv((long)n * GG, 256), 256, 0, P->ctrl, m, n, P->a_ptr,
                                    P->a_ind, P->a_val, P->head, P->stat, P->rho, (const double *)nullptr,
                                    P->trow, P->svec, 0));
        CK(cudaMemcpyAsync(trow.data(), P->trow, n * sizeof(double), cudaMemcpyDeviceToHost, P->stream));
        CK(cudaMemcpyAsync(dev_head.data(), P->head, (m + n) * sizeof(int), cudaMemcpyDeviceToHost, P->stream));
        CK(cudaStreamSynchronize(P->stream));
        P->n_sync++;
        /* un-scale: alfa_unscaled = alfa_scaled * scale(basic) / scale(non-basic) */
        double sb = (k < m) ? 1.0 / P->h_rii[k] : P->h_sjj[k - m];
        std::fill(alfa.begin(), alfa.end(), 0.0);
        for (int j = 0; j < n; j++) {
            int kk = dev_head[m + j];
            double sn = (kk < m) ? 1.0 / P->h_rii[kk] : P->h_sjj[kk - m];
            alfa[kk] = trow[j] * sb / sn;
        }
        return 0;
    }

    /* table rows of all the structural variables js at once: one launch per chunk of
       tab_cap rows, one read-back; fills tab_rows_h / dev_head */
    int tab_rows_batch(const std::vector<int> &js)
    {
        tab_js = js;
        const int F = (int)js.size();
        tab_rows_h.assign((size_t)F * n, 0.0);
        if (F == 0) return 0;
        if (!P->valid) return GLPB_ESTATE;
        if (!d_tab_pos) {
            tab_cap = std::max(1, std::min(m, 64));
            if (cudaMalloc((void **)&d_tab_pos, tab_cap * sizeof(int)) != cudaSuccess ||
                cudaMalloc((void **)&d_tab_rho, (size_t)tab_cap * m * sizeof(double)) != cudaSuccess ||
                cudaMalloc((void **)&d_tab_trow, (size_t)tab_cap * n * sizeof(double)) != cudaSuccess) {
                glpb_set_error("branch-and-bound: device allocation failed");
                return GLPB_ENOMEM;
            }
        }
        std::vector<int> posof(m + n, -1), pos(F);
        for (int i = 0; i < m; i++) posof[P->h_head[i] - 1] = i;
        for (int f = 0; f < F; f++) {
            pos[f] = posof[m + js[f]];
            if (pos[f] < 0) return GLPB_ESTATE;
        }
        for (int f0 = 0; f0 < F; f0 += tab_cap) {
            const int fc = std::min(tab_cap, F - f0);
            CK(cudaMemcpyAsync(d_tab_pos, pos.data() + f0, fc * sizeof(int), cudaMemcpyHostToDevice, P->stream));
            LAUNCH(P, k_tab_rows, fc, 256, 0, P->ctrl, m, n, d_tab_pos, P->T, P->ldt, P->a_ptr, P->a_ind, P->a_val,
                   P->at_ptr, P->at_ind, P->at_val, P->head, P->bind, P->stat, P->rslot, P->cslot, d_tab_rho,
                   d_tab_trow);
            CK(cudaMemcpyAsync(tab_rows_h.data() + (size_t)f0 * n, d_tab_trow, (size_t)fc * n * sizeof(double),
                               cudaMemcpyDeviceToHost, P->stream));
        }
        CK(cudaMemcpyAsync(dev_head.data(), P->head, (m + n) * sizeof(int), cudaMemcpyDeviceToHost, P->stream));
        CK(cudaStreamSynchronize(P->stream));
        P->n_sync++;
        return 0;
    }

    /* row f of the batch, un-scaled like tab_row does */
    void cached_row(int f, std::vector<double> &alfa) const
    {
        const int k = m + tab_js[f];
        const double sb = P->h_sjj[k - m];
        const double *tr = tab_rows_h.data() + (size_t)f * n;
        std::fill(alfa.begin(), alfa.end(), 0.0);
        for (int j = 0; j < n; j++) {
            const int kk = dev_head[m + j];
            const double sn = (kk < m) ? 1.0 / P->h_rii[kk] : P->h_sjj[kk - m];
            alfa[kk] = tr[j] * sb / sn;
        }
    }

    /* glp_dual_rtest, lib/glpapi12.js:687-762, over all non-basic kk in
       ascending order (the order glp_eval_tab_row lists them); returns kk or -1 */
    int dual_rtest(const std::vector<double> &alfa_row, int dirn, double eps) const
    {
        double obj = (dir() == GLP_MIN ? +1.0 : -1.0);
        int piv = -1;
        double teta = DBL_MAX, big = 0.0;
        for (int kk = 0; kk < m + n; kk++) {
            double v = alfa_row[kk];
            if (v == 0.0) continue;
            int st = P->h_stat[kk];
            if (st == GLP_BS) continue;
            double cost = P->h_dual[kk];
            double alfa = (dirn > 0 ? +v : -v), temp;
            if (st == GLP_NL) { if (alfa < +eps) continue; temp = (obj * cost) / alfa; }
            else if (st == GLP_NU) { if (alfa > -eps) continue; temp = (obj * cost) / alfa; }
            else if (st == GLP_NF) { if (-eps < alfa && alfa < +eps) continue; temp = 0.0; }
            else continue;
            if (temp < 0.0) temp = 0.0;
            if (teta > temp || (teta == temp && big < fabs(alfa))) { piv = kk; teta = temp; big = fabs(alfa); }
        }
        return piv;
    }

    double fixed_sign_dual(int kk) const
    {
        int st = P->h_stat[kk];
        double g = P->h_dual[kk];
        if (dir() == GLP_MIN) { if ((st == GLP_NL && g < 0.0) || (st == GLP_NU && g > 0.0) || st == GLP_NF) g = 0.0; }
        else { if ((st == GLP_NL && g > 0.0) || (st == GLP_NU && g < 0.0) || st == GLP_NF) g = 0.0; }
        return g;
    }

    /* lib/glpios01.js:615-728 ios_eval_degrad */
    int eval_degrad(int j, double &dn, double &up)
    {
        std::vector<double> alfa(m + n);
        int rc = tab_row(m + j, alfa);
        if (rc) return rc;
        double beta = P->h_prim[m + j];
        for (int kase = -1; kase <= +1; kase += 2) {
            int kk = dual_rtest(alfa, kase, 1e-9);
            if (kk < 0) {
                double inf = (dir() == GLP_MIN ? +DBL_MAX : -DBL_MAX);
                if (kase < 0) dn = inf; else up = inf;
                continue;
            }
            double delta = ((kase < 0 ? floor(beta) : ceil(beta)) - beta) / alfa[kk];
            double dz = fixed_sign_dual(kk) * delta;
            if (kase < 0) dn = P->obj_val + dz; else up = P->obj_val + dz;
        }
        return 0;
    }

    /* lib/glpios01.js:866-910 ios_solve_node */
    int solve_node()
    {
        glpb_smcp sp;
        glpb_init_smcp(&sp);
        sp.msg_lev = 0;
        sp.meth = GLP_DUALP;
        if (feas()) { if (dir() == GLP_MIN) sp.obj_ul = P->mip_obj; else sp.obj_ll = P->mip_obj; }
        solved++;
        return glpb_simplex(P, &sp);
    }

    /* lib/glpios03.js:56-116 check_integrality */
    void check_integrality()
    {
        int ii_cnt = 0;
        double ii_sum = 0.0;
        for (int j = 0; j < n; j++) {
            int k = m + j;
            non_int[j] = 0;
            if (P->h_kind[j] != GLP_IV || P->h_stat[k] != GLP_BS) continue;
            int t = P->h_type[k];
            double l = P->h_lb[k], u = P->h_ub[k], x = P->h_prim[k];
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
        pool[curr].ii_cnt = ii_cnt;
        pool[curr].ii_sum = ii_sum;
    }

    /* lib/glpios03.js:118-139 record_solution */
    void record_solution()
    {
        P->mip_stat = GLP_FEAS;
        P->mip_obj = P->obj_val;
        have_sol = have_cut = true;
        for (int i = 0; i < m; i++) P->h_mipx[i] = P->h_prim[i];
        for (int j = 0; j < n; j++)
            P->h_mipx[m + j] = (P->h_kind[j] == GLP_IV) ? floor(P->h_prim[m + j] + 0.5) : P->h_prim[m + j];
    }

    /* lib/glpios03.js:307-377 fix_by_red_cost */
    void fix_by_red_cost()
    {
        double obj = P->obj_val;
        for (int j = 0; j < n; j++) {
            int k = m + j;
            if (P->h_kind[j] != GLP_IV) continue;
            double l = P->h_lb[k], u = P->h_ub[k], dj = P->h_dual[k];
            int st = P->h_stat[k];
            if (dir() == GLP_MIN) {
                if (st == GLP_NL) { if (dj < 0.0) dj = 0.0; if (obj + dj >= P->mip_obj) set_bnds(k, GLP_FX, l, l); }
                else if (st == GLP_NU) { if (dj > 0.0) dj = 0.0; if (obj - dj >= P->mip_obj) set_bnds(k, GLP_FX, u, u); }
            } else {
                if (st == GLP_NL) { if (dj > 0.0) dj = 0.0; if (obj + dj <= P->mip_obj) set_bnds(k, GLP_FX, l, l); }
                else if (st == GLP_NU) { if (dj < 0.0) dj = 0.0; if (obj - dj <= P->mip_obj) set_bnds(k, GLP_FX, u, u); }
            }
        }
    }

    /* lib/glpios09.js:60-82 branch_mostf */
    int branch_mostf(int &next) const
    {
        int jj = -1;
        double most = DBL_MAX;
        for (int j = 0; j < n; j++)
            if (non_int[j]) {
                double beta = P->h_prim[m + j], temp = floor(beta) + 0.5;
                if (most > fabs(beta - temp)) { jj = j; most = fabs(beta - temp); next = (beta < temp) ? DN_BRNCH : UP_BRNCH; }
            }
        return jj;
    }

    /* lib/glpios09.js:84-270 branch_drtom */
    int branch_drtom(int &next, int &jj_out)
    {
        std::vector<double> alfa(m + n);
        int jj = -1;
        double degrad = -1.0, dz_dn = 0, dz_up = 0;
        /* all the fractional variables' table rows in one device round trip */
        std::vector<int> js;
        for (int j = 0; j < n; j++) if (non_int[j]) js.push_back(j);
        int rcb = tab_rows_batch(js);
        if (rcb) return rcb;
        int f = -1;
        for (int j = 0; j < n; j++) {
            if (!non_int[j]) continue;
            double x = P->h_prim[m + j];
            cached_row(++f, alfa);
            for (int kase = -1; kase <= +1; kase += 2) {
                int kk = dual_rtest(alfa, kase, 1e-9);
                double delta_z;
                if (kk < 0) delta_z = (dir() == GLP_MIN ? +DBL_MAX : -DBL_MAX);
                else {
                    double delta_k = ((kase < 0 ? floor(x) : ceil(x)) - x) / alfa[kk];
                    if (kk >= m && P->h_kind[kk - m] != GLP_CV)
                        if (fabs(delta_k - floor(delta_k + 0.5)) > 1e-3)
                            delta_k = (delta_k > 0.0) ? ceil(delta_k) : floor(delta_k);
                    delta_z = fixed_sign_dual(kk) * delta_k;
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
        if (degrad < 1e-6 * (1.0 + 0.001 * fabs(P->obj_val))) jj = branch_mostf(next);
        jj_out = jj;
        return 0;
    }

    int choose_var(int &next, int &jj)
    {
        if (parm.br_tech == GLP_BR_FFV || parm.br_tech == GLP_BR_LFV) {
            int j;
            if (parm.br_tech == GLP_BR_FFV) { for (j = 0; j < n; j++) if (non_int[j]) break; }
            else { for (j = n - 1; j >= 0; j--) if (non_int[j]) break; }
            double beta = P->h_prim[m + j];
            next = (beta - floor(beta) < ceil(beta) - beta) ? DN_BRNCH : UP_BRNCH;
            jj = j;
            return 0;
        }
        if (parm.br_tech == GLP_BR_MFV) { jj = branch_mostf(next); return 0; }
        return branch_drtom(next, jj);
    }

    /* lib/glpios12.js ios_choose_node; nodes imported from other ranks hang
       directly under the root, so "up" always exists for non-root nodes */
    int choose_node() const
    {
        if (parm.bt_tech == GLP_BT_DFS) return tail;
        if (parm.bt_tech == GLP_BT_BFS) return head;
        auto up_sum = [&](int t) { return pool[t].up >= 0 ? pool[pool[t].up].ii_sum : 0.0; };
        if (parm.bt_tech == GLP_BT_BLB) {
            int best = -1;
            if (dir() == GLP_MIN) {
                double bound = +DBL_MAX;
                for (int t = head; t >= 0; t = pool[t].next) if (bound > pool[t].bound) bound = pool[t].bound;
                double eps = 0.001 * (1.0 + fabs(bound));
                for (int t = head; t >= 0; t = pool[t].next)
                    if (pool[t].bound <= bound + eps)
                        if (best < 0 || up_sum(best) > up_sum(t)) best = t;
            } else {
                double bound = -DBL_MAX;
                for (int t = head; t >= 0; t = pool[t].next) if (bound < pool[t].bound) bound = pool[t].bound;
                double eps = 0.001 * (1.0 + fabs(bound));
                for (int t = head; t >= 0; t = pool[t].next)
                    if (pool[t].bound >= bound - eps)
                        if (best < 0 || pool[best].lp_obj < pool[t].lp_obj) best = t;
            }
            return best;
        }
        int p = -1;
        double best = DBL_MAX;
        if (!feas()) {
            for (int t = head; t >= 0; t = pool[t].next) if (best > up_sum(t)) { p = t; best = up_sum(t); }
        } else {
            double deg = pool[0].ii_sum > 0 ? (P->mip_obj - pool[0].bound) / pool[0].ii_sum : 0.0;
            for (int t = head; t >= 0; t = pool[t].next) {
                int u = pool[t].up >= 0 ? pool[t].up : t;
                double obj = pool[u].bound + deg * pool[u].ii_sum;
                if (dir() == GLP_MAX) obj = -obj;
                if (best > obj) { p = t; best = obj; }
            }
        }
        return p;
    }

    /* ---- node preprocessing: lib/glpios02.js ---- */
    struct RowInfo { double f_min, f_max; int j_min, j_max; };

    static void row_info(int len, const double *a, const double *l, const double *u, RowInfo &f)
    {
        f.f_min = 0.0; f.j_min = -1;
        for (int j = 0; j < len; j++) {
            double bnd = a[j] > 0.0 ? l[j] : u[j];
            if (bnd == (a[j] > 0.0 ? -DBL_MAX : +DBL_MAX)) {
                if (f.j_min < 0) f.j_min = j; else { f.f_min = -DBL_MAX; f.j_min = -1; break; }
            } else f.f_min += a[j] * bnd;
        }
        f.f_max = 0.0; f.j_max = -1;
        for (int j = 0; j < len; j++) {
            double bnd = a[j] > 0.0 ? u[j] : l[j];
            if (bnd == (a[j] > 0.0 ? +DBL_MAX : -DBL_MAX)) {
                if (f.j_max < 0) f.j_max = j; else { f.f_max = +DBL_MAX; f.j_max = -1; break; }
            } else f.f_max += a[j] * bnd;
        }
    }

    static int row_bounds(const RowInfo &f, double &L, double &U)
    {
        double LL = (f.j_min < 0 ? f.f_min : -DBL_MAX), UU = (f.j_max < 0 ? f.f_max : +DBL_MAX);
        if (L != -DBL_MAX && UU < L - 1e-3 * (1.0 + fabs(L))) return 1;
        if (U != +DBL_MAX && LL > U + 1e-3 * (1.0 + fabs(U))) return 1;
        if (L != -DBL_MAX && LL > L - 1e-12 * (1.0 + fabs(L))) L = -DBL_MAX;
        if (U != +DBL_MAX && UU < U + 1e-12 * (1.0 + fabs(U))) U = +DBL_MAX;
        return 0;
    }

    static int col_bounds(const RowInfo &f, const double *a, double L, double U, const double *l, const double *u,
                          int flag, int k, double &lj, double &uj)
    {
        double ilb, iub, ll, uu;
        if (L == -DBL_MAX || f.f_max == +DBL_MAX) ilb = -DBL_MAX;
        else if (f.j_max < 0) ilb = L - (f.f_max - a[k] * (a[k] > 0.0 ? u[k] : l[k]));
        else if (f.j_max == k) ilb = L - f.f_max;
        else ilb = -DBL_MAX;
        if (U == +DBL_MAX || f.f_min == -DBL_MAX) iub = +DBL_MAX;
        else if (f.j_min < 0) iub = U - (f.f_min - a[k] * (a[k] > 0.0 ? l[k] : u[k]));
        else if (f.j_min == k) iub = U - f.f_min;
        else iub = +DBL_MAX;
        if (fabs(a[k]) < 1e-6) { ll = -DBL_MAX; uu = +DBL_MAX; }
        else if (a[k] > 0.0) { ll = (ilb == -DBL_MAX ? -DBL_MAX : ilb / a[k]); uu = (iub == +DBL_MAX ? +DBL_MAX : iub / a[k]); }
        else { ll = (iub == +DBL_MAX ? -DBL_MAX : iub / a[k]); uu = (ilb == -DBL_MAX ? +DBL_MAX : ilb / a[k]); }
        if (flag) {
            if (ll != -DBL_MAX) ll = (ll - floor(ll) < 1e-3 ? floor(ll) : ceil(ll));
            if (uu != +DBL_MAX) uu = (ceil(uu) - uu < 1e-3 ? ceil(uu) : floor(uu));
        }
        lj = l[k]; uj = u[k];
        if (lj != -DBL_MAX && uu < lj - 1e-3 * (1.0 + fabs(lj))) return 1;
        if (uj != +DBL_MAX && ll > uj + 1e-3 * (1.0 + fabs(uj))) return 1;
        if (ll != -DBL_MAX && lj < ll - 1e-3 * (1.0 + fabs(ll))) lj = ll;
        if (uu != +DBL_MAX && uj > uu + 1e-3 * (1.0 + fabs(uu))) uj = uu;
        if (!(lj == -DBL_MAX || uj == +DBL_MAX)) {
            double t1 = fabs(lj), t2 = fabs(uj);
            double eps = 1e-10 * (1.0 + (t1 <= t2 ? t1 : t2));
            if (lj > uj - eps) {
                if (lj == l[k]) uj = lj;
                else if (uj == u[k]) lj = uj;
                else if (t1 <= t2) uj = lj;
                else lj = uj;
            }
        }
        return 0;
    }

    static int efficiency(int flag, double l, double u, double ll, double uu)
    {
        int eff = 0;
        if (l < ll) {
            if (flag || l == -DBL_MAX) eff++;
            else if (ll - l >= 0.25 * ((u == +DBL_MAX) ? 1.0 + fabs(l) : 1.0 + (u - l))) eff++;
        }
        if (u > uu) {
            if (flag || u == +DBL_MAX) eff++;
            else if (u - uu >= 0.25 * ((l == -DBL_MAX) ? 1.0 + fabs(u) : 1.0 + (u - l))) eff++;
        }
        return eff;
    }

    int preprocess_node(int max_pass)
    {
        /* index 0 is the objective row, rows are 1..m here like the reference */
        std::vector<double> L(1 + m), U(1 + m), l(n), u(n);
        if (feas()) {
            if (dir() == GLP_MIN) { L[0] = -DBL_MAX; U[0] = P->mip_obj - P->c0; }
            else { L[0] = P->mip_obj - P->c0; U[0] = +DBL_MAX; }
        } else { L[0] = -DBL_MAX; U[0] = +DBL_MAX; }
        auto glb = [&](int k) { int t = P->h_type[k]; return (t == GLP_FR || t == GLP_UP) ? -DBL_MAX : P->h_lb[k]; };
        auto gub = [&](int k) { int t = P->h_type[k]; return (t == GLP_FR || t == GLP_LO) ? +DBL_MAX : (t == GLP_FX ? P->h_lb[k] : P->h_ub[k]); };
        for (int i = 0; i < m; i++) { L[1 + i] = glb(i); U[1 + i] = gub(i); }
        for (int j = 0; j < n; j++) { l[j] = glb(m + j); u[j] = gub(m + j); }
        std::vector<int> list(2 + m), mark(1 + m, 0), pass(1 + m, 0), ind(n);
        std::vector<double> val(n), lbv(n), ubv(n);
        int size = 0;
        for (int i = 0; i <= m; i++) { list[size++] = i; mark[i] = 1; }
        while (size > 0) {
            int i = list[--size];
            mark[i] = 0;
            pass[i]++;
            if (L[i] == -DBL_MAX && U[i] == +DBL_MAX) continue;
            int len = 0;
            if (i == 0) {
                for (int j = 0; j < n; j++) if (P->h_coef[j] != 0.0) { ind[len] = j; val[len] = P->h_coef[j]; len++; }
            } else
                for (int t = P->h_atptr[i - 1]; t < P->h_atptr[i]; t++) { ind[len] = P->h_atind[t]; val[len] = P->h_atval[t]; len++; }
            for (int k = 0; k < len; k++) { lbv[k] = l[ind[k]]; ubv[k] = u[ind[k]]; }
            RowInfo f;
            row_info(len, val.data(), lbv.data(), ubv.data(), f);
            if (row_bounds(f, L[i], U[i])) return 1;
            if (L[i] == -DBL_MAX && U[i] == +DBL_MAX) continue;
            for (int k = 0; k < len; k++) {
                int j = ind[k];
                int flag = P->h_kind[j] != GLP_CV;
                double ll, uu;
                if (col_bounds(f, val.data(), L[i], U[i], lbv.data(), ubv.data(), flag, k, ll, uu)) return 1;
                int eff = efficiency(flag, l[j], u[j], ll, uu);
                l[j] = ll; u[j] = uu;
                if (eff > 0)
                    for (int t = P->h_aptr[j]; t < P->h_aptr[j + 1]; t++) {
                        int ii = P->h_aind[t] + 1;
                        if (pass[ii] >= max_pass) continue;
                        if (L[ii] == -DBL_MAX && U[ii] == +DBL_MAX) continue;
                        if (!mark[ii]) { list[size++] = ii; mark[ii] = 1; }
                    }
            }
        }
        for (int i = 0; i < m; i++)
            if (P->h_stat[i] == GLP_BS) {
                double Li = L[1 + i], Ui = U[1 + i];
                if (Li == -DBL_MAX && Ui == +DBL_MAX) set_bnds(i, GLP_FR, 0.0, 0.0);
                else if (Ui == +DBL_MAX) set_bnds(i, GLP_LO, Li, 0.0);
                else if (Li == -DBL_MAX) set_bnds(i, GLP_UP, 0.0, Ui);
            }
        for (int j = 0; j < n; j++) {
            int type;
            if (l[j] == -DBL_MAX && u[j] == +DBL_MAX) type = GLP_FR;
            else if (u[j] == +DBL_MAX) type = GLP_LO;
            else if (l[j] == -DBL_MAX) type = GLP_UP;
            else if (l[j] != u[j]) type = GLP_DB;
            else type = GLP_FX;
            set_bnds(m + j, type, l[j], u[j]);
        }
        return 0;
    }

    void improve_bound(int p, double bnd)
    {
        if (dir() == GLP_MIN) { if (pool[p].bound < bnd) pool[p].bound = bnd; }
        else { if (pool[p].bound > bnd) pool[p].bound = bnd; }
    }

    /* lib/glpios03.js:141-305 branch_on; returns 0/1/2 or a negative error */
    int branch_on(int j, int next)
    {
        int k = m + j, type = P->h_type[k], dn_type, up_type;
        double l = P->h_lb[k], u = P->h_ub[k], beta = P->h_prim[k];
        double new_ub = floor(beta), new_lb = ceil(beta);
        switch (type) {
        case GLP_FR: dn_type = GLP_UP; up_type = GLP_LO; break;
        case GLP_LO: dn_type = (l == new_ub ? GLP_FX : GLP_DB); up_type = GLP_LO; break;
        case GLP_UP: dn_type = GLP_UP; up_type = (new_lb == u ? GLP_FX : GLP_DB); break;
        default: dn_type = (l == new_ub ? GLP_FX : GLP_DB); up_type = (new_lb == u ? GLP_FX : GLP_DB); break;
        }
        double dn_lp = 0, up_lp = 0;
        int rc = eval_degrad(j, dn_lp, up_lp);
        if (rc) return rc;
        double dn_bnd = round_bound(dn_lp), up_bnd = round_bound(up_lp);
        bool dn_bad = !is_hopeful(dn_bnd), up_bad = !is_hopeful(up_bnd);
        if (dn_bad && up_bad) return 2;
        if (up_bad) { set_bnds(k, dn_type, l, new_ub); pool[curr].lp_obj = dn_lp; improve_bound(curr, dn_bnd); return 1; }
        if (dn_bad) { set_bnds(k, up_type, new_lb, u); pool[curr].lp_obj = up_lp; improve_bound(curr, up_bnd); return 1; }
        int p = curr;
        freeze();
        unlink(p);                       /* ios_clone_node: the parent becomes inactive */
        int c1 = new_node(p), c2 = new_node(p);
        pool[c1].b.push_back(BndChange{k, dn_type, l, new_ub});
        pool[c1].lp_obj = dn_lp; improve_bound(c1, dn_bnd);
        pool[c2].b.push_back(BndChange{k, up_type, new_lb, u});
        pool[c2].lp_obj = up_lp; improve_bound(c2, up_bnd);
        child = (next == NO_BRNCH ? -1 : (next == DN_BRNCH ? c1 : c2));
        return 0;
    }

    void cleanup_the_tree()
    {
        for (int t = head, nx; t >= 0; t = nx) {
            nx = pool[t].next;
            if (!is_hopeful(pool[t].bound)) del(t);
        }
    }

    /* ios_driver, lib/glpios03.js:507-951, resumable: returns 0 when the local
       pool is exhausted, 1 when max_nodes node LPs were solved in this slice */
    int run(long max_nodes)
    {
        long budget = max_nodes;
        for (;;) {
            if (state == S_LOOP) {
                if (head < 0) return 0;
                if (budget == 0) return 1;
                int next_p;
                if (a_cnt == 1) next_p = head;
                else if (child >= 0 && pool[child].used) next_p = child;
                else next_p = choose_node();
                revive(next_p);
                child = -1;
                cur_p = curr;
                state = S_MORE;
            }
            if (state == S_MORE) {
                if (parm.node_lim >= 0 && solved >= parm.node_lim) return GLP_ESTOP;
                if (parm.tm_lim < INT_MAX && (parm.tm_lim - 1) <= (now_ms() - tm_beg)) return GLP_ETMLIM;
                /* ios_relative_gap <= mip_gap: lib/glpios03.js:615-625, lib/glpios01.js:821-864 */
                if (feas() && parm.mip_gap > 0.0) {
                    double best = pool[curr].bound;
                    for (int t = head; t >= 0; t = pool[t].next)
                        if (dir() == GLP_MIN ? pool[t].bound < best : pool[t].bound > best) best = pool[t].bound;
                    if (fabs(P->mip_obj - best) / (fabs(P->mip_obj) + DBL_EPSILON) <= parm.mip_gap) return GLP_EMIPGAP;
                }
                int lvl = pool[curr].level;
                if (parm.pp_tech == GLP_PP_ROOT) { if (lvl == 0 && preprocess_node(100)) { state = S_FATH; continue; } }
                else if (parm.pp_tech == GLP_PP_ALL) { if (preprocess_node(lvl == 0 ? 100 : 10)) { state = S_FATH; continue; } }
                if (!is_hopeful(pool[curr].bound)) { state = S_FATH; continue; }
                int ret = solve_node();
                if (budget > 0) budget--;
                if (ret < 0) return ret;
                if (!(ret == 0 || ret == GLP_EOBJLL || ret == GLP_EOBJUL)) return GLP_EFAIL;
                int p_stat = P->pbs_stat, d_stat = P->dbs_stat;
                if (p_stat == GLP_FEAS && d_stat == GLP_FEAS) {}
                else if (d_stat == GLP_NOFEAS) return GLP_EFAIL;
                else if (p_stat == GLP_INFEAS && d_stat == GLP_FEAS) { state = S_FATH; continue; }
                else if (p_stat == GLP_NOFEAS) { state = S_FATH; continue; }
                else return GLP_EFAIL;
                pool[curr].lp_obj = P->obj_val;
                improve_bound(curr, round_bound(P->obj_val));
                if (!is_hopeful(pool[curr].bound)) { state = S_FATH; continue; }
                check_integrality();
                if (pool[curr].ii_cnt == 0) { record_solution(); state = S_FATH; continue; }
                if (feas()) fix_by_red_cost();
                int next = NO_BRNCH, jv = -1;
                int rc = choose_var(next, jv);
                if (rc) return rc;
                ret = branch_on(jv, next);
                if (ret < 0) return ret;
                if (ret == 0) { state = S_LOOP; continue; }
                if (ret == 1) { state = S_MORE; continue; }
                state = S_FATH;
            }
            if (state == S_FATH) {
                freeze();
                if (cur_p == 0) { unlink(0); pool[0].used = false; }
                else del(cur_p);
                if (feas()) cleanup_the_tree();
                state = S_LOOP;
            }
        }
    }

    /* ---- migration: a node is shipped as its complete (type, lb, ub, stat)
           vectors plus (bound, lp_obj, level); 18 (m+n) + 32 bytes ---- */
    size_t record_bytes() const { return (size_t)(m + n) * (1 + 8 + 8 + 1) + 32; }

    /* state of node p as full vectors (replays the diffs on scratch copies) */
    void materialize(int p, std::vector<int> &type, std::vector<double> &lb, std::vector<double> &ub,
                     std::vector<int> &stat) const
    {
        type = root_frozen ? root_type : P->h_type; lb = root_frozen ? root_lb : P->h_lb;
        ub = root_frozen ? root_ub : P->h_ub; stat = root_frozen ? root_stat : P->h_stat;
        std::vector<int> path;
        for (int t = p; t >= 0; t = pool[t].up) path.push_back(t);
        for (auto it = path.rbegin(); it != path.rend(); ++it) {
            for (const BndChange &b : pool[*it].b) { type[b.k] = b.type; lb[b.k] = b.lb; ub[b.k] = b.ub; }
            for (const StatChange &s : pool[*it].s) stat[s.k] = s.stat;
        }
    }
};

void glpb_mip_free(glpb_mip *T) { delete T; }

/* ------------------------------------------------------------------ */
/* C ABI                                                              */
/* ------------------------------------------------------------------ */

static int mip_check(glpb_prob *P)
{
    /* lib/glpapi09.js:337-364 integer bounds; :67-72 optimal root required */
    for (int j = 0; j < P->n; j++)
        if (P->h_kind[j] == GLP_IV) {
            int k = P->m + j, t = P->h_type[k];
            if ((t == GLP_LO || t == GLP_DB) && P->h_lb[k] != floor(P->h_lb[k])) return GLP_EBOUND;
            if ((t == GLP_UP || t == GLP_DB) && P->h_ub[k] != floor(P->h_ub[k])) return GLP_EBOUND;
            if (t == GLP_FX && P->h_lb[k] != floor(P->h_lb[k])) return GLP_EBOUND;
        }
    if (glpb_get_status(P) != GLP_OPT) return GLP_EROOT;
    return 0;
}

extern "C" int glpb_mip_begin(glpb_prob *P, const glpb_iocp *parm_)
{
    if (!P) return GLPB_EINVAL;
    glpb_iocp parm;
    if (parm_) parm = *parm_; else glpb_init_iocp(&parm);
    if (!(parm.br_tech >= GLP_BR_FFV && parm.br_tech <= GLP_BR_DTH)) {
        glpb_set_error("glp_intopt: br_tech = %d not supported on the device path", parm.br_tech);
        return GLPB_EINVAL;
    }
    if (!(parm.bt_tech >= GLP_BT_DFS && parm.bt_tech <= GLP_BT_BPH)) return GLPB_EINVAL;
    if (!(0.0 < parm.tol_int && parm.tol_int < 1.0) || !(0.0 < parm.tol_obj && parm.tol_obj < 1.0)) return GLPB_EINVAL;
    P->mip_stat = GLP_UNDEF;
    P->mip_obj = 0.0;
    int rc = mip_check(P);
    if (rc) return rc;
    delete P->mip;
    P->mip = new glpb_mip(P, parm);
    P->mip->create();
    return 0;
}

/* returns 0: local pool exhausted; 1: slice limit reached; GLP_E*: stopped */
extern "C" int glpb_mip_run(glpb_prob *P, long max_nodes, long *solved)
{
    if (!P || !P->mip) return GLPB_ESTATE;
    long before = P->mip->solved;
    int ret = P->mip->run(max_nodes);
    if (solved) *solved = P->mip->solved - before;
    P->mip_nodes = P->mip->solved;
    return ret;
}

/* incumbent exchange: what this rank knows, and a cut-off learnt elsewhere */
extern "C" int glpb_mip_get_incumbent(glpb_prob *P, int *has_solution, double *obj)
{
    if (!P || !P->mip) return GLPB_ESTATE;
    if (has_solution) *has_solution = P->mip->have_sol;
    if (obj) *obj = P->mip->have_cut ? P->mip_obj : (P->dir == GLP_MIN ? +DBL_MAX : -DBL_MAX);
    return 0;
}

extern "C" int glpb_mip_set_cutoff(glpb_prob *P, double obj)
{
    if (!P || !P->mip) return GLPB_ESTATE;
    glpb_mip &T = *P->mip;
    bool better = !T.have_cut || (P->dir == GLP_MIN ? obj < P->mip_obj : obj > P->mip_obj);
    if (better) {
        P->mip_obj = obj;
        T.have_cut = true;
        T.have_sol = false;       /* the solution vector lives on another rank */
        if (T.curr < 0) T.cleanup_the_tree();
    }
    return 0;
}

extern "C" int glpb_mip_open_count(glpb_prob *P)
{
    if (!P || !P->mip) return GLPB_ESTATE;
    return P->mip->a_cnt + (P->mip->curr >= 0 ? 1 : 0);
}

extern "C" long glpb_mip_record_bytes(glpb_prob *P)
{
    if (!P || !P->mip) return GLPB_ESTATE;
    return (long)P->mip->record_bytes();
}

/* remove up to max_count open nodes (worst bound first is kept; the exported
   ones are taken from the tail of the active list) and serialise them */
extern "C" int glpb_mip_export_nodes(glpb_prob *P, int max_count, void *buf, long cap, int *count)
{
    if (!P || !P->mip || !buf || !count) return GLPB_ESTATE;
    glpb_mip &T = *P->mip;
    if (T.curr >= 0) return GLPB_ESTATE;
    const int mn = T.m + T.n;
    const size_t rb = T.record_bytes();
    char *out = (char *)buf;
    int done = 0;
    std::vector<int> type, stat;
    std::vector<double> lb, ub;
    /* a negative count means "up to -count nodes, the pool may be emptied"
       (used once, to partition the replicated ramp-up pool) */
    const int keep = max_count < 0 ? 0 : 1;
    if (max_count < 0) max_count = -max_count;
    while (done < max_count && T.a_cnt > keep && (long)((done + 1) * rb) <= cap) {
        int p = T.tail;
        if (p == 0) break;
        T.materialize(p, type, lb, ub, stat);
        double hdr[4] = {T.pool[p].bound, T.pool[p].lp_obj, (double)T.pool[p].level, T.pool[T.pool[p].up >= 0 ? T.pool[p].up : p].ii_sum};
        memcpy(out, hdr, 32); out += 32;
        for (int k = 0; k < mn; k++) *out++ = (char)type[k];
        for (int k = 0; k < mn; k++) *out++ = (char)stat[k];
        memcpy(out, lb.data(), mn * 8); out += mn * 8;
        memcpy(out, ub.data(), mn * 8); out += mn * 8;
        T.del(p);
        done++;
    }
    *count = done;
    return 0;
}

extern "C" int glpb_mip_import_nodes(glpb_prob *P, const void *buf, int count)
{
    if (!P || !P->mip || (!buf && count > 0)) return GLPB_ESTATE;
    glpb_mip &T = *P->mip;
    if (T.curr >= 0 || !T.root_frozen) return GLPB_ESTATE;
    const int mn = T.m + T.n;
    const char *in = (const char *)buf;
    /* the root must stay allocated as the common ancestor of imported nodes */
    for (int c = 0; c < count; c++) {
        double hdr[4];
        memcpy(hdr, in, 32); in += 32;
        const char *type = in; in += mn;
        const char *stat = in; in += mn;
        const double *lb = (const double *)in; in += mn * 8;
        const double *ub = (const double *)in; in += mn * 8;
        if (!T.is_hopeful(hdr[0])) continue;
        if (!T.pool[0].used) { T.pool[0].used = true; T.pool[0].count = 0; T.pool[0].prev = T.pool[0].next = -1; }
        int p = T.new_node(0);
        MNode &nd = T.pool[p];
        nd.bound = hdr[0]; nd.lp_obj = hdr[1]; nd.level = (int)hdr[2];
        T.pool[0].ii_sum = hdr[3];
        for (int k = 0; k < mn; k++) {
            double l, u;
            memcpy(&l, &lb[k], 8); memcpy(&u, &ub[k], 8);
            if (!(T.root_type[k] == type[k] && T.root_lb[k] == l && T.root_ub[k] == u))
                nd.b.push_back(BndChange{k, type[k], l, u});
            if (T.root_stat[k] != stat[k]) nd.s.push_back(StatChange{k, stat[k]});
        }
    }
    return 0;
}

/* ios_delete_tree + the status mapping of solve_mip (lib/glpapi09.js:82-112);
   ret is the code the search ended with (0 = tree exhausted everywhere) */
extern "C" int glpb_mip_end(glpb_prob *P, int ret)
{
    if (!P || !P->mip) return GLPB_ESTATE;
    glpb_mip &T = *P->mip;
    if (T.curr >= 0) T.freeze();
    T.destroy();
    P->mip_nodes = T.solved;
    if (T.have_sol) P->mip_stat = (ret == 0) ? GLP_OPT : GLP_FEAS;
    else if (T.have_cut) P->mip_stat = GLP_UNDEF;     /* the optimum was found on another rank */
    else P->mip_stat = (ret == 0) ? GLP_NOFEAS : GLP_UNDEF;
    delete P->mip;
    P->mip = nullptr;
    return ret;
}

static bool bnb_batched(const glpb_prob *P, const glpb_iocp &parm);
static int bnb_intopt_batched(glpb_prob *P, const glpb_iocp *parm);

extern "C" int glpb_intopt(glpb_prob *P, const glpb_iocp *parm)
{
    if (P) {
        glpb_iocp pr;
        if (parm) pr = *parm; else glpb_init_iocp(&pr);
        if (bnb_batched(P, pr)) return bnb_intopt_batched(P, parm);
    }
    int rc = glpb_mip_begin(P, parm);
    if (rc) return rc;
    int ret = glpb_mip_run(P, -1, nullptr);
    if (ret == 1) ret = GLP_ESTOP;
    return glpb_mip_end(P, ret);
}

extern "C" int glpb_get_mip(glpb_prob *P, int *mip_stat, double *mip_obj, double *mipx, long *nodes)
{
    if (!P) return GLPB_EINVAL;
    if (mip_stat) *mip_stat = P->mip_stat;
    if (mip_obj) *mip_obj = P->mip_obj;
    if (mipx) memcpy(mipx, P->h_mipx.data(), (P->m + P->n) * sizeof(double));
    if (nodes) *nodes = P->mip_nodes;
    return 0;
}
