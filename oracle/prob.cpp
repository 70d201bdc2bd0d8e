/* prob.cpp -- CPU ORACLE (test infrastructure only; see glpo.h).
 * Minimal problem object, CPLEX-LP reader/writer, glp_factorize and the
 * glp_simplex driver, restated from lib/glpapi01.js, lib/glpapi05.js,
 * lib/glpapi06.js, lib/glpapi12.js, lib/glpcpx.js, lib/glprng01/02.js.
 */
#include "glpo.h"
#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cctype>
#include <map>
#include <sstream>

namespace glpo {

Prob::~Prob() { delete bfd; }

double xtime_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

/* ---- RNG: lib/glprng01.js:1-60, lib/glprng02.js ---- */
static inline int mod_diff(int x, int y) { return (int)(((unsigned)x - (unsigned)y) & 0x7FFFFFFFu); }

static int flip_cycle(RNG &r)
{
    int ii, jj;
    for (ii = 1, jj = 32; jj <= 55; ii++, jj++) r.A[ii] = mod_diff(r.A[ii], r.A[jj]);
    for (jj = 1; ii <= 55; ii++, jj++) r.A[ii] = mod_diff(r.A[ii], r.A[jj]);
    r.fptr = 54;
    return r.A[55];
}

void rng_init(RNG &r, int seed)
{
    r.A[0] = -1;
    for (int i = 1; i <= 55; i++) r.A[i] = 0;
    r.fptr = 0;
    int prev = seed, next = 1;
    seed = prev = mod_diff(prev, 0);
    r.A[55] = prev;
    for (int i = 21; i; i = (i + 21) % 55) {
        r.A[i] = next;
        next = mod_diff(prev, next);
        if (seed & 1) seed = 0x40000000 + (seed >> 1);
        else seed >>= 1;
        next = mod_diff(next, seed);
        prev = r.A[i];
    }
    for (int t = 0; t < 5; t++) flip_cycle(r);
}

int rng_next(RNG &r) { return r.A[r.fptr] >= 0 ? r.A[r.fptr--] : flip_cycle(r); }

int rng_unif_rand(RNG &r, int m)
{
    const unsigned two31 = 0x80000000u;
    unsigned t = two31 - (two31 % (unsigned)m);
    int x;
    do { x = rng_next(r); } while (t <= (unsigned)x);
    return x % m;
}

double rng_unif_01(RNG &r) { return (double)rng_next(r) / 2147483647.0; }

/* ---- problem object ---- */

/* lib/glpapi01.js:87-123 glp_add_rows: new rows are free and basic */
void prob_add_rows(Prob &P, int nrs)
{
    int m_new = P.m + nrs;
    P.r_type.resize(1 + m_new, GLP_FR); P.r_stat.resize(1 + m_new, GLP_BS);
    P.r_bind.resize(1 + m_new, 0);
    P.r_lb.resize(1 + m_new, 0.0); P.r_ub.resize(1 + m_new, 0.0);
    P.r_rii.resize(1 + m_new, 1.0); P.r_prim.resize(1 + m_new, 0.0);
    P.r_dual.resize(1 + m_new, 0.0); P.r_mipx.resize(1 + m_new, 0.0);
    P.r_name.resize(1 + m_new);
    P.row_list.resize(1 + m_new);
    P.head.resize(1 + m_new, 0);
    P.m = m_new;
    P.valid = 0;
}

/* lib/glpapi01.js:154-171 glp_add_cols: new columns are fixed at 0, NS */
void prob_add_cols(Prob &P, int ncs)
{
    int n_new = P.n + ncs;
    P.c_type.resize(1 + n_new, GLP_FX); P.c_stat.resize(1 + n_new, GLP_NS);
    P.c_bind.resize(1 + n_new, 0); P.c_kind.resize(1 + n_new, GLP_CV);
    P.c_lb.resize(1 + n_new, 0.0); P.c_ub.resize(1 + n_new, 0.0);
    P.c_coef.resize(1 + n_new, 0.0); P.c_sjj.resize(1 + n_new, 1.0);
    P.c_prim.resize(1 + n_new, 0.0); P.c_dual.resize(1 + n_new, 0.0);
    P.c_mipx.resize(1 + n_new, 0.0);
    P.c_name.resize(1 + n_new);
    P.col_list.resize(1 + n_new);
    P.n = n_new;
}

static void set_bnds(int &type_, double &lb_, double &ub_, int &stat, int type, double lb, double ub)
{
    type_ = type;
    switch (type) {
    case GLP_FR: lb_ = ub_ = 0.0; if (stat != GLP_BS) stat = GLP_NF; break;
    case GLP_LO: lb_ = lb; ub_ = 0.0; if (stat != GLP_BS) stat = GLP_NL; break;
    case GLP_UP: lb_ = 0.0; ub_ = ub; if (stat != GLP_BS) stat = GLP_NU; break;
    case GLP_DB:
        lb_ = lb; ub_ = ub;
        if (!(stat == GLP_BS || stat == GLP_NL || stat == GLP_NU))
            stat = (fabs(lb) <= fabs(ub) ? GLP_NL : GLP_NU);
        break;
    case GLP_FX: lb_ = ub_ = lb; if (stat != GLP_BS) stat = GLP_NS; break;
    default: assert(!"bad bound type");
    }
}

/* lib/glpapi01.js:217-248 glp_set_row_bnds (does not touch 'valid') */
void prob_set_row_bnds(Prob &P, int i, int type, double lb, double ub)
{
    set_bnds(P.r_type[i], P.r_lb[i], P.r_ub[i], P.r_stat[i], type, lb, ub);
}

/* lib/glpapi01.js:250-281 glp_set_col_bnds */
void prob_set_col_bnds(Prob &P, int j, int type, double lb, double ub)
{
    set_bnds(P.c_type[j], P.c_lb[j], P.c_ub[j], P.c_stat[j], type, lb, ub);
}

/* lib/glpapi01.js:295-378 glp_set_mat_row: new elements are PREPENDED to the
   row list and to each column list; zeros are dropped afterwards */
void prob_set_mat_row(Prob &P, int i, int len, const int *ind, const double *val)
{
    for (const Elem &e : P.row_list[i]) {
        auto &cl = P.col_list[e.idx];
        for (size_t t = 0; t < cl.size(); t++)
            if (cl[t].idx == i) { cl.erase(cl.begin() + t); break; }
        P.nnz--;
        if (P.c_stat[e.idx] == GLP_BS) P.valid = 0;
    }
    P.row_list[i].clear();
    for (int k = 1; k <= len; k++) {
        if (val[k] == 0.0) continue;
        int j = ind[k];
        P.row_list[i].insert(P.row_list[i].begin(), Elem{j, val[k]});
        P.col_list[j].insert(P.col_list[j].begin(), Elem{i, val[k]});
        P.nnz++;
        if (P.c_stat[j] == GLP_BS) P.valid = 0;
    }
}

/* lib/glpapi01.js:620-649 glp_sort_matrix: rows ascending in j, columns
   ascending in i */
void prob_sort_matrix(Prob &P)
{
    auto cmp = [](const Elem &a, const Elem &b) { return a.idx < b.idx; };
    for (int i = 1; i <= P.m; i++) std::stable_sort(P.row_list[i].begin(), P.row_list[i].end(), cmp);
    for (int j = 1; j <= P.n; j++) std::stable_sort(P.col_list[j].begin(), P.col_list[j].end(), cmp);
}

static int norm_stat(int type, int stat)
{
    if (stat == GLP_BS) return stat;
    switch (type) {
    case GLP_FR: return GLP_NF;
    case GLP_LO: return GLP_NL;
    case GLP_UP: return GLP_NU;
    case GLP_DB: return stat != GLP_NU ? GLP_NL : GLP_NU;
    case GLP_FX: return GLP_NS;
    }
    assert(!"bad type");
    return stat;
}

/* lib/glpapi05.js:1-23 glp_set_row_stat */
void prob_set_row_stat(Prob &P, int i, int stat)
{
    stat = norm_stat(P.r_type[i], stat);
    if ((P.r_stat[i] == GLP_BS) != (stat == GLP_BS)) P.valid = 0;
    P.r_stat[i] = stat;
}

/* lib/glpapi05.js:25-47 glp_set_col_stat */
void prob_set_col_stat(Prob &P, int j, int stat)
{
    stat = norm_stat(P.c_type[j], stat);
    if ((P.c_stat[j] == GLP_BS) != (stat == GLP_BS)) P.valid = 0;
    P.c_stat[j] = stat;
}

/* lib/glpapi05.js:49-63 glp_std_basis */
void prob_std_basis(Prob &P)
{
    for (int i = 1; i <= P.m; i++) prob_set_row_stat(P, i, GLP_BS);
    for (int j = 1; j <= P.n; j++) {
        if (P.c_type[j] == GLP_DB && fabs(P.c_lb[j]) > fabs(P.c_ub[j]))
            prob_set_col_stat(P, j, GLP_NU);
        else
            prob_set_col_stat(P, j, GLP_NL);
    }
}

/* lib/glpapi12.js:7-33 b_col */
static int b_col(void *info, int j, int *ind, double *val)
{
    Prob &lp = *(Prob *)info;
    int k = lp.head[j];
    if (k <= lp.m) { ind[1] = k; val[1] = 1.0; return 1; }
    int len = 0;
    for (const Elem &e : lp.col_list[k - lp.m]) {
        len++;
        ind[len] = e.idx;
        val[len] = -lp.r_rii[e.idx] * e.val * lp.c_sjj[k - lp.m];
    }
    return len;
}

/* lib/glpapi12.js:5-100 glp_factorize */
int prob_factorize(Prob &lp)
{
    const int m = lp.m, n = lp.n;
    lp.valid = 0;
    int j = 0;
    for (int k = 1; k <= m + n; k++) {
        int stat;
        if (k <= m) { stat = lp.r_stat[k]; lp.r_bind[k] = 0; }
        else { stat = lp.c_stat[k - m]; lp.c_bind[k - m] = 0; }
        if (stat == GLP_BS) {
            j++;
            if (j > m) return GLP_EBADB;
            lp.head[j] = k;
            if (k <= m) lp.r_bind[k] = j; else lp.c_bind[k - m] = j;
        }
    }
    if (j < m) return GLP_EBADB;
    if (m > 0) {
        if (lp.bfd == nullptr) lp.bfd = new BFD();
        switch (bfd_factorize(*lp.bfd, m, b_col, &lp)) {
        case 0: break;
        case BFD_ESING: return GLP_ESING;
        case BFD_ECOND: return GLP_ECOND;
        default: assert(!"bad bfd code");
        }
        lp.valid = 1;
    }
    return 0;
}

/* lib/glpapi06.js:398-427 glp_get_status */
int prob_get_status(const Prob &lp)
{
    int status = lp.pbs_stat;
    if (status == GLP_FEAS) {
        if (lp.dbs_stat == GLP_FEAS) status = GLP_OPT;
        else if (lp.dbs_stat == GLP_NOFEAS) status = GLP_UNBND;
    }
    return status;
}

/* lib/glpapi06.js:149-258 trivial_lp */
static void trivial_lp(Prob &P, const SMCP &parm)
{
    P.valid = 0;
    P.pbs_stat = P.dbs_stat = GLP_FEAS;
    P.obj_val = P.c0;
    P.some = 0;
    for (int i = 1; i <= P.m; i++) {
        P.r_stat[i] = GLP_BS;
        P.r_prim[i] = P.r_dual[i] = 0.0;
        int t = P.r_type[i];
        if (t == GLP_LO || t == GLP_DB || t == GLP_FX)
            if (P.r_lb[i] > +parm.tol_bnd) {
                P.pbs_stat = GLP_NOFEAS;
                if (P.some == 0 && parm.meth != GLP_PRIMAL) P.some = i;
            }
        if (t == GLP_UP || t == GLP_DB || t == GLP_FX)
            if (P.r_ub[i] < -parm.tol_bnd) {
                P.pbs_stat = GLP_NOFEAS;
                if (P.some == 0 && parm.meth != GLP_PRIMAL) P.some = i;
            }
    }
    double zeta = 1.0;
    for (int j = 1; j <= P.n; j++)
        if (zeta < fabs(P.c_coef[j])) zeta = fabs(P.c_coef[j]);
    zeta = (P.dir == GLP_MIN ? +1.0 : -1.0) / zeta;
    for (int j = 1; j <= P.n; j++) {
        int t = P.c_type[j];
        bool lo;
        if (t == GLP_FR) { P.c_stat[j] = GLP_NF; P.c_prim[j] = 0.0; }
        else if (t == GLP_FX) { P.c_stat[j] = GLP_NS; P.c_prim[j] = P.c_lb[j]; }
        else {
            if (t == GLP_LO) lo = true;
            else if (t == GLP_UP) lo = false;
            else if (zeta * P.c_coef[j] > 0.0) lo = true;
            else if (zeta * P.c_coef[j] < 0.0) lo = false;
            else lo = (fabs(P.c_lb[j]) <= fabs(P.c_ub[j]));
            if (lo) { P.c_stat[j] = GLP_NL; P.c_prim[j] = P.c_lb[j]; }
            else { P.c_stat[j] = GLP_NU; P.c_prim[j] = P.c_ub[j]; }
        }
        P.c_dual[j] = P.c_coef[j];
        P.obj_val += P.c_coef[j] * P.c_prim[j];
        if (t == GLP_FR || t == GLP_LO)
            if (zeta * P.c_dual[j] < -parm.tol_dj) {
                P.dbs_stat = GLP_NOFEAS;
                if (P.some == 0 && parm.meth == GLP_PRIMAL) P.some = P.m + j;
            }
        if (t == GLP_FR || t == GLP_UP)
            if (zeta * P.c_dual[j] > +parm.tol_dj) {
                P.dbs_stat = GLP_NOFEAS;
                if (P.some == 0 && parm.meth == GLP_PRIMAL) P.some = P.m + j;
            }
    }
}

/* lib/glpapi06.js:3-39 solve_lp */
static int solve_lp(Prob &P, const SMCP &parm, const Hook *hook)
{
    int ret;
    if (!(P.m == 0 || P.valid)) {
        ret = prob_factorize(P);
        if (ret != 0) return ret;
    }
    if (parm.meth == GLP_PRIMAL)
        ret = spx_primal(P, parm, hook);
    else if (parm.meth == GLP_DUALP) {
        ret = spx_dual(P, parm, hook);
        if (ret == GLP_EFAIL && P.valid) ret = spx_primal(P, parm, hook);
    } else
        ret = spx_dual(P, parm, hook);
    return ret;
}

/* lib/glpapi06.js:261-339 glp_simplex (presolve OFF path only: the oracle has no
   restatement of the presolver -- the product's one is checked against the
   reference's own runs directly, tests/golden/ref_npp.json) */
int simplex(Prob &P, const SMCP &parm, const Hook *hook)
{
    P.pbs_stat = P.dbs_stat = GLP_UNDEF;
    P.obj_val = 0.0;
    P.some = 0;
    for (int i = 1; i <= P.m; i++)
        if (P.r_type[i] == GLP_DB && P.r_lb[i] >= P.r_ub[i]) return GLP_EBOUND;
    for (int j = 1; j <= P.n; j++)
        if (P.c_type[j] == GLP_DB && P.c_lb[j] >= P.c_ub[j]) return GLP_EBOUND;
    if (P.nnz == 0) { trivial_lp(P, parm); return 0; }
    return solve_lp(P, parm, hook);
}

/* ---- CPLEX LP reader (lib/glpcpx.js:10-753), own tokenizer ---- */
namespace {

enum Tok { T_EOF, T_MINIMIZE, T_MAXIMIZE, T_SUBJECT_TO, T_BOUNDS, T_GENERAL,
           T_INTEGER, T_BINARY, T_END, T_NAME, T_NUMBER, T_PLUS, T_MINUS,
           T_COLON, T_LE, T_GE, T_EQ };

struct Reader {
    const std::string &s;
    size_t pos = 0;
    int c = '\n';       /* current character (look-ahead) */
    Tok token = T_EOF;
    std::string image;
    double value = 0.0;
    std::string err;
    int line = 0;
    Prob &P;
    std::map<std::string, int> col_index, row_index;
    std::vector<double> lb, ub;

    Reader(Prob &P_, const std::string &text) : s(text), P(P_) {}

    void fail(const std::string &msg)
    {
        if (err.empty()) { std::ostringstream o; o << "line " << line << ": " << msg; err = o.str(); }
        throw 1;
    }

    void read_char()
    {
        if (c == '\n') line++;
        if (pos >= s.size()) { c = (c == '\n' || c == -1) ? -1 : '\n'; return; }
        c = (unsigned char)s[pos++];
        if (c == '\r') { read_char(); return; }
        if (c == '\t') c = ' ';
    }

    static bool name_char(int ch)
    {
        return isalnum(ch) || (ch > 0 && strchr("!\"#$%&()/,.;?@_`'{}|~", ch) != nullptr);
    }

    static bool same(const std::string &a, const char *b)
    {
        size_t n = strlen(b);
        if (a.size() != n) return false;
        for (size_t i = 0; i < n; i++)
            if (tolower((unsigned char)a[i]) != b[i]) return false;
        return true;
    }

    void skip_blank() { while (c == ' ') read_char(); }

    void scan_token()
    {
        bool at_bol;
        token = T_EOF; image.clear(); value = 0.0;
        for (;;) {
            at_bol = false;
            while (c == ' ') read_char();
            if (c == -1) { token = T_EOF; return; }
            if (c == '\n') {
                read_char();
                /* a name starting a line may be a keyword */
                while (c == ' ') read_char();
                if (isalpha(c)) { at_bol = true; break; }
                continue;
            }
            if (c == '\\') { while (c != '\n' && c != -1) read_char(); continue; }
            break;
        }
        if (c == -1) { token = T_EOF; return; }
        if (isalpha(c) || (c != '.' && name_char(c) && !isdigit(c))) {
            while (name_char(c)) { image.push_back((char)c); read_char(); }
            token = T_NAME;
            if (at_bol) {
                if (same(image, "minimize") || same(image, "minimum") || same(image, "min")) token = T_MINIMIZE;
                else if (same(image, "maximize") || same(image, "maximum") || same(image, "max")) token = T_MAXIMIZE;
                else if (same(image, "subject") || same(image, "such")) {
                    size_t save_pos = pos; int save_c = c, save_line = line;
                    skip_blank();
                    std::string w;
                    while (isalpha(c)) { w.push_back((char)c); read_char(); }
                    if ((same(image, "subject") && same(w, "to")) || (same(image, "such") && same(w, "that")))
                        token = T_SUBJECT_TO;
                    else { pos = save_pos; c = save_c; line = save_line; }
                }
                else if (same(image, "st") || same(image, "s.t.") || same(image, "st.")) token = T_SUBJECT_TO;
                else if (same(image, "bounds") || same(image, "bound")) token = T_BOUNDS;
                else if (same(image, "general") || same(image, "generals") || same(image, "gen")) token = T_GENERAL;
                else if (same(image, "integer") || same(image, "integers") || same(image, "int")) token = T_INTEGER;
                else if (same(image, "binary") || same(image, "binaries") || same(image, "bin")) token = T_BINARY;
                else if (same(image, "end")) token = T_END;
            }
            return;
        }
        if (isdigit(c) || c == '.') {
            while (isdigit(c)) { image.push_back((char)c); read_char(); }
            if (c == '.') {
                image.push_back('.'); read_char();
                while (isdigit(c)) { image.push_back((char)c); read_char(); }
            }
            if (c == 'e' || c == 'E') {
                image.push_back('e'); read_char();
                if (c == '+' || c == '-') { image.push_back((char)c); read_char(); }
                if (!isdigit(c)) fail("numeric constant `" + image + "' incomplete");
                while (isdigit(c)) { image.push_back((char)c); read_char(); }
            }
            value = strtod(image.c_str(), nullptr);
            token = T_NUMBER;
            return;
        }
        if (c == '+') { token = T_PLUS; image = "+"; read_char(); return; }
        if (c == '-') { token = T_MINUS; image = "-"; read_char(); return; }
        if (c == ':') { token = T_COLON; image = ":"; read_char(); return; }
        if (c == '<') { token = T_LE; image = "<"; read_char(); if (c == '=') read_char(); return; }
        if (c == '>') { token = T_GE; image = ">"; read_char(); if (c == '=') read_char(); return; }
        if (c == '=') {
            token = T_EQ; image = "="; read_char();
            if (c == '<') { token = T_LE; read_char(); }
            else if (c == '>') { token = T_GE; read_char(); }
            return;
        }
        fail(std::string("character `") + (char)c + "' not recognized");
    }

    bool colon_follows()
    {
        /* the reference tests csa.c == ':' right after the name */
        return c == ':';
    }

    /* lib/glpcpx.js:260-291 find_col */
    int find_col(const std::string &name)
    {
        auto it = col_index.find(name);
        if (it != col_index.end()) return it->second;
        prob_add_cols(P, 1);
        int j = P.n;
        P.c_name[j] = name;
        col_index[name] = j;
        lb.resize(1 + j, +DBL_MAX); ub.resize(1 + j, -DBL_MAX);
        lb[j] = +DBL_MAX; ub[j] = -DBL_MAX;
        return j;
    }

    /* lib/glpcpx.js:293-343 parse_linear_form */
    int parse_linear_form(std::vector<int> &ind, std::vector<double> &val)
    {
        ind.assign(1, 0); val.assign(1, 0.0);
        std::vector<char> used(1 + P.n + 1, 0);
        for (;;) {
            double sgn = +1.0, coef = 1.0;
            if (token == T_PLUS) { sgn = +1.0; scan_token(); }
            else if (token == T_MINUS) { sgn = -1.0; scan_token(); }
            if (token == T_NUMBER) { coef = value; scan_token(); }
            if (token != T_NAME) fail("missing variable name");
            int j = find_col(image);
            if ((int)used.size() <= j) used.resize(j + 1, 0);
            if (used[j]) fail("multiple use of variable `" + image + "' not allowed");
            used[j] = 1;
            ind.push_back(j); val.push_back(sgn * coef);
            scan_token();
            if (token == T_PLUS || token == T_MINUS) continue;
            break;
        }
        int newlen = 0;
        for (size_t k = 1; k < ind.size(); k++)
            if (val[k] != 0.0) { newlen++; ind[newlen] = ind[k]; val[newlen] = val[k]; }
        return newlen;
    }

    void set_lb(int j, double v) { lb[j] = v; }
    void set_ub(int j, double v) { ub[j] = v; }

    double parse_signed_or_inf(double sgn, bool lower)
    {
        if (token == T_NUMBER) { double v = sgn * value; scan_token(); return v; }
        if (same(image, "infinity") || same(image, "inf")) {
            if (lower && sgn > 0.0) fail("invalid use of `+inf' as lower bound");
            if (!lower && sgn < 0.0) fail("invalid use of `-inf' as upper bound");
            scan_token();
            return lower ? -DBL_MAX : +DBL_MAX;
        }
        fail(lower ? "missing lower bound" : "missing upper bound");
        return 0.0;
    }

    /* lib/glpcpx.js:425-596 parse_bounds */
    void parse_bounds()
    {
        scan_token();
        for (;;) {
            if (!(token == T_PLUS || token == T_MINUS || token == T_NUMBER || token == T_NAME)) return;
            bool lb_flag = false;
            double lbv = 0.0;
            if (token == T_PLUS || token == T_MINUS) {
                lb_flag = true;
                double sgn = (token == T_PLUS ? +1.0 : -1.0);
                scan_token();
                lbv = parse_signed_or_inf(sgn, true);
            } else if (token == T_NUMBER) {
                lb_flag = true; lbv = value; scan_token();
            }
            if (lb_flag) {
                if (token != T_LE) fail("missing `<', `<=', or `=<' after lower bound");
                scan_token();
            }
            if (token != T_NAME) fail("missing variable name");
            int j = find_col(image);
            if (lb_flag) set_lb(j, lbv);
            scan_token();
            if (token == T_LE) {
                scan_token();
                if (token == T_PLUS || token == T_MINUS) {
                    double sgn = (token == T_PLUS ? +1.0 : -1.0);
                    scan_token();
                    set_ub(j, parse_signed_or_inf(sgn, false));
                } else if (token == T_NUMBER) { set_ub(j, value); scan_token(); }
                else fail("missing upper bound");
            } else if (token == T_GE) {
                if (lb_flag) fail("invalid bound definition");
                scan_token();
                if (token == T_PLUS || token == T_MINUS) {
                    double sgn = (token == T_PLUS ? +1.0 : -1.0);
                    scan_token();
                    set_lb(j, parse_signed_or_inf(sgn, true));
                } else if (token == T_NUMBER) { set_lb(j, value); scan_token(); }
                else fail("missing lower bound");
            } else if (token == T_EQ) {
                if (lb_flag) fail("invalid bound definition");
                scan_token();
                double sgn = +1.0;
                if (token == T_PLUS || token == T_MINUS) { sgn = (token == T_PLUS ? +1.0 : -1.0); scan_token(); }
                if (token != T_NUMBER) fail("missing fixed value");
                set_lb(j, sgn * value); set_ub(j, sgn * value);
                scan_token();
            } else if (token == T_NAME && same(image, "free")) {
                if (lb_flag) fail("invalid bound definition");
                set_lb(j, -DBL_MAX); set_ub(j, +DBL_MAX);
                scan_token();
            } else if (!lb_flag)
                fail("invalid bound definition");
        }
    }

    /* lib/glpcpx.js:598-630 parse_integer */
    void parse_integer()
    {
        bool binary = (token == T_BINARY);
        scan_token();
        while (token == T_NAME) {
            int j = find_col(image);
            P.c_kind[j] = GLP_IV;
            if (binary) { set_lb(j, 0.0); set_ub(j, 1.0); }
            scan_token();
        }
    }

    void run()
    {
        std::vector<int> ind; std::vector<double> val;
        scan_token();
        if (!(token == T_MINIMIZE || token == T_MAXIMIZE)) fail("`minimize' or `maximize' keyword missing");
        /* parse_objective: lib/glpcpx.js:345-372 */
        P.dir = (token == T_MINIMIZE ? GLP_MIN : GLP_MAX);
        scan_token();
        if (token == T_NAME && colon_follows()) {
            P.obj_name = image; scan_token(); scan_token();
        } else
            P.obj_name = "obj";
        int len = parse_linear_form(ind, val);
        for (int k = 1; k <= len; k++) P.c_coef[ind[k]] = val[k];
        if (token != T_SUBJECT_TO) fail("constraints section missing");
        /* parse_constraints: lib/glpcpx.js:374-436 */
        scan_token();
        for (;;) {
            prob_add_rows(P, 1);
            int i = P.m;
            if (token == T_NAME && colon_follows()) {
                if (row_index.count(image)) fail("constraint `" + image + "' multiply defined");
                P.r_name[i] = image; row_index[image] = i;
                scan_token(); scan_token();
            } else {
                std::ostringstream o; o << "r." << line; P.r_name[i] = o.str();
            }
            len = parse_linear_form(ind, val);
            prob_set_mat_row(P, i, len, ind.data(), val.data());
            int type;
            if (token == T_LE) type = GLP_UP;
            else if (token == T_GE) type = GLP_LO;
            else if (token == T_EQ) type = GLP_FX;
            else { fail("missing constraint sense"); return; }
            scan_token();
            double sgn = +1.0;
            if (token == T_PLUS) { scan_token(); }
            else if (token == T_MINUS) { sgn = -1.0; scan_token(); }
            if (token != T_NUMBER) fail("missing right-hand side");
            prob_set_row_bnds(P, i, type, sgn * value, sgn * value);
            scan_token();
            if (token == T_PLUS || token == T_MINUS || token == T_NUMBER || token == T_NAME) continue;
            break;
        }
        if (token == T_BOUNDS) parse_bounds();
        while (token == T_GENERAL || token == T_INTEGER || token == T_BINARY) parse_integer();
        if (token == T_END) scan_token();
        else if (token != T_EOF) fail("symbol " + image + " in wrong position");
        if (token != T_EOF) fail("extra symbol(s) detected beyond `end'");
        /* lib/glpcpx.js:698-717 deferred column bounds */
        for (int j = 1; j <= P.n; j++) {
            double l = lb[j], u = ub[j];
            int type;
            if (l == +DBL_MAX) l = 0.0;
            if (u == -DBL_MAX) u = +DBL_MAX;
            if (l == -DBL_MAX && u == +DBL_MAX) type = GLP_FR;
            else if (u == +DBL_MAX) type = GLP_LO;
            else if (l == -DBL_MAX) type = GLP_UP;
            else if (l != u) type = GLP_DB;
            else type = GLP_FX;
            prob_set_col_bnds(P, j, type, l, u);
        }
        prob_sort_matrix(P);
    }
};

} /* anonymous namespace */

int read_lp(Prob &P, const std::string &text, std::string &err)
{
    Reader r(P, text);
    try { r.run(); }
    catch (int) { err = r.err; return 1; }
    return 0;
}

/* lib/glpcpx.js:755-998 glp_write_lp (same section layout; one term per line
   so any reader takes it back) */
std::string write_lp(const Prob &P)
{
    std::ostringstream o;
    char buf[64];
    auto num = [&](double v) { snprintf(buf, sizeof buf, "%.17g", v); return std::string(buf); };
    auto cname = [&](int j) { if (!P.c_name[j].empty()) return P.c_name[j]; return "x_" + std::to_string(j); };
    auto rname = [&](int i) { if (!P.r_name[i].empty()) return P.r_name[i]; return "r_" + std::to_string(i); };
    o << "\\* Problem: " << (P.name.empty() ? "Unknown" : P.name) << " *\\\n\n";
    o << (P.dir == GLP_MIN ? "Minimize\n" : "Maximize\n");
    o << " " << (P.obj_name.empty() ? "obj" : P.obj_name) << ":";
    int cnt = 0;
    for (int j = 1; j <= P.n; j++)
        if (P.c_coef[j] != 0.0) {
            o << (P.c_coef[j] < 0 ? " - " : " + ") << num(fabs(P.c_coef[j])) << " " << cname(j) << "\n";
            cnt++;
        }
    if (cnt == 0 && P.n > 0) o << " 0 " << cname(1) << "\n";
    o << "\nSubject To\n";
    for (int i = 1; i <= P.m; i++) {
        o << " " << rname(i) << ":";
        for (const Elem &e : P.row_list[i])
            o << (e.val < 0 ? " - " : " + ") << num(fabs(e.val)) << " " << cname(e.idx) << "\n";
        if (P.row_list[i].empty()) o << " 0 " << cname(1) << "\n";
        switch (P.r_type[i]) {
        case GLP_LO: o << " >= " << num(P.r_lb[i]) << "\n"; break;
        case GLP_UP: o << " <= " << num(P.r_ub[i]) << "\n"; break;
        case GLP_FX: o << " = " << num(P.r_lb[i]) << "\n"; break;
        default: assert(!"row type not representable here");
        }
    }
    o << "\nBounds\n";
    for (int j = 1; j <= P.n; j++) {
        switch (P.c_type[j]) {
        case GLP_FR: o << " " << cname(j) << " free\n"; break;
        case GLP_LO: if (P.c_lb[j] != 0.0) o << " " << cname(j) << " >= " << num(P.c_lb[j]) << "\n"; break;
        case GLP_UP: o << " -inf <= " << cname(j) << " <= " << num(P.c_ub[j]) << "\n"; break;
        case GLP_DB: o << " " << num(P.c_lb[j]) << " <= " << cname(j) << " <= " << num(P.c_ub[j]) << "\n"; break;
        case GLP_FX: o << " " << cname(j) << " = " << num(P.c_lb[j]) << "\n"; break;
        }
    }
    bool any = false;
    for (int j = 1; j <= P.n; j++)
        if (P.c_kind[j] == GLP_IV) { if (!any) { o << "\nGenerals\n"; any = true; } o << " " << cname(j) << "\n"; }
    o << "\nEnd\n";
    return o.str();
}

} /* namespace glpo */
