/* lpformat.cpp -- native reader of the CPLEX LP format.
 *
 * glp_read_lp (lib/glpcpx.js:10-753) feeds the path its inputs; the reference
 * scans the text one character per callback.  This reader takes the whole
 * text, tokenises it in one pass and emits exactly what the reference leaves
 * in the glp_prob after its final glp_sort_matrix: rows and columns in order
 * of first appearance, column types from the collected bounds
 * (lib/glpcpx.js:697-715), the matrix by columns with ascending row indices.
 * Host code; O(length of the text).
 *
 * Grammar restated from the reference's scanner and section parsers:
 *   objective  = (min|max keyword) [name ':'] linear-form
 *   constraint = [name ':'] linear-form (<|<=|=<|>|>=|=>|=) [sign] number
 *   bounds     = [sign] (number|inf) <= name [<= [sign] (number|inf)]
 *              | name (<=|>=|=) value | name free
 *   generals / integers / binaries = lists of names
 * Keywords are recognised only when a letter stands in column 0 of a line (the
 * reference looks for them right after a newline, lib/glpcpx.js:187-192); a name
 * is followed by a colon if ':' is the next non-blank character of its line; the
 * right-hand side must end its line (not even a comment may follow); unnamed rows
 * are called "r.<line number>"; '\' starts a comment.
 */
#include "../../include/glpb200.h"
#include <cctype>
#include <cfloat>
#include <charconv>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <new>
#include <vector>

void glpb_set_error(const char *fmt, ...);

namespace {

enum Kind { T_EOF, T_MIN, T_MAX, T_ST, T_BOUNDS, T_GEN, T_INT, T_BIN, T_END, T_NAME, T_NUM, T_PLUS, T_MINUS,
            T_COLON, T_LE, T_GE, T_EQ };

struct Token {
    Kind kind;
    std::string image;
    double value;
    bool colon; /* a ':' is the next significant character on the same line (the reference tests
                   csa.c after scan_token has skipped the blanks behind the token, lib/glpcpx.js:256-257) */
    int line;
    bool eol;   /* nothing but blanks follows on the line -- not even a comment (lib/glpcpx.js:431-433) */
};

struct SyntaxError {
    std::string msg;
    int line;
};

const char *NAME_EXTRA = "!\"#$%&()/,.;?@_`'{}|~";

bool name_char(unsigned char c) { return isalnum(c) || (c && strchr(NAME_EXTRA, c)); }
bool name_start(unsigned char c) { return isalpha(c) || (c && c != '.' && strchr(NAME_EXTRA, c)); }

std::string lower(const std::string &s)
{
    std::string r(s);
    for (char &c : r) c = (char)tolower((unsigned char)c);
    return r;
}

Kind keyword(const std::string &low)
{
    static const std::unordered_map<std::string, Kind> kw = {
        {"minimize", T_MIN}, {"minimum", T_MIN}, {"min", T_MIN}, {"maximize", T_MAX}, {"maximum", T_MAX},
        {"max", T_MAX}, {"st", T_ST}, {"s.t.", T_ST}, {"st.", T_ST}, {"bounds", T_BOUNDS}, {"bound", T_BOUNDS},
        {"general", T_GEN}, {"generals", T_GEN}, {"gen", T_GEN}, {"integer", T_INT}, {"integers", T_INT},
        {"int", T_INT}, {"binary", T_BIN}, {"binaries", T_BIN}, {"bin", T_BIN}, {"end", T_END}};
    auto it = kw.find(low);
    return it == kw.end() ? T_NAME : it->second;
}

std::vector<Token> tokenize(const char *text, long len)
{
    std::vector<Token> toks;
    long pos = 0;
    int lineno = 0;
    while (pos < len) {
        long eol = pos;
        while (eol < len && text[eol] != '\n') eol++;
        lineno++;
        std::string line(text + pos, (size_t)(eol - pos));
        pos = eol + 1;
        for (char &c : line)
            if (c == '\t' || c == '\r' || c == '\v' || c == '\f') c = ' ';
        const size_t full = line.size();                 /* the comment stays for the end-of-line tests */
        size_t cut = line.find('\\');
        const size_t n = (cut != std::string::npos) ? cut : full;
        /* next significant position on the RAW line (a comment counts: after it the reference's
           current character is '\\', not the end of the line) */
        auto rest = [&](size_t j) { while (j < full && line[j] == ' ') j++; return j; };
        size_t i = 0;
        while (i < n) {
            unsigned char ch = (unsigned char)line[i];
            if (ch == ' ') { i++; continue; }
            if (name_start(ch)) {
                size_t j = i;
                while (j < n && name_char((unsigned char)line[j])) j++;
                std::string image = line.substr(i, j - i);
                Kind kind = T_NAME;
                if (i == 0 && isalpha(ch)) {             /* keywords: a letter in column 0 (lib/glpcpx.js:187-192) */
                    std::string low = lower(image);
                    if (low == "subject" || low == "such") {
                        /* "subject to" / "such that": the second word must follow on the same line */
                        size_t r = j;
                        while (r < n && line[r] == ' ') r++;
                        const std::string want = (low == "subject") ? "to" : "that";
                        if (n - r >= want.size() && lower(line.substr(r, want.size())) == want &&
                            (r + want.size() == n || !isalnum((unsigned char)line[r + want.size()]))) {
                            kind = T_ST;
                            j = r + want.size();
                        }
                    } else
                        kind = keyword(low);
                }
                { size_t r = rest(j); toks.push_back({kind, image, 0.0, r < full && line[r] == ':', lineno, r >= full}); }
                i = j;
            } else if (isdigit(ch) || ch == '.') {
                size_t j = i;
                while (j < n && isdigit((unsigned char)line[j])) j++;
                if (j < n && line[j] == '.') {
                    j++;
                    if (j - i == 1 && !(j < n && isdigit((unsigned char)line[j])))
                        throw SyntaxError{"invalid use of decimal point", lineno};
                    while (j < n && isdigit((unsigned char)line[j])) j++;
                }
                if (j < n && (line[j] == 'e' || line[j] == 'E')) {
                    j++;
                    if (j < n && (line[j] == '+' || line[j] == '-')) j++;
                    if (!(j < n && isdigit((unsigned char)line[j])))       /* lib/glpcpx.js:214-218 */
                        throw SyntaxError{"numeric constant `" + line.substr(i, j - i) + "' incomplete", lineno};
                    while (j < n && isdigit((unsigned char)line[j])) j++;
                }
                std::string image = line.substr(i, j - i);
                char *endp = nullptr;
                double v = strtod(image.c_str(), &endp);
                if (endp == image.c_str() || *endp != '\0')
                    throw SyntaxError{"numeric constant `" + image + "' not recognized", lineno};
                toks.push_back({T_NUM, image, v, false, lineno, rest(j) >= full});
                i = j;
            } else if (ch == '+' || ch == '-' || ch == ':') {
                toks.push_back({ch == '+' ? T_PLUS : (ch == '-' ? T_MINUS : T_COLON), std::string(1, (char)ch), 0.0, false, lineno, rest(i + 1) >= full});
                i++;
            } else if (ch == '<' || ch == '>' || ch == '=') {
                size_t j = i + 1;
                if (j < n && (line[j] == '<' || line[j] == '>' || line[j] == '=')) j++;
                std::string op = line.substr(i, j - i);
                Kind kind = op.find('<') != std::string::npos ? T_LE : (op.find('>') != std::string::npos ? T_GE : T_EQ);
                toks.push_back({kind, op, 0.0, false, lineno, rest(j) >= full});
                i = j;
            } else
                throw SyntaxError{std::string("character `") + (char)ch + "' not recognized", lineno};
        }
    }
    toks.push_back({T_EOF, "", 0.0, false, lineno, true});
    return toks;
}

struct Reader {
    std::vector<Token> toks;
    size_t pos = 0;
    int dir = 1; /* GLP_MIN */
    std::string obj_name;
    std::vector<std::string> row_name, col_name;
    std::unordered_map<std::string, int> row_of, col_of;
    std::vector<double> coef;                         /* by column */
    std::vector<int> kind;                            /* GLP_CV = 1 / GLP_IV = 2 */
    std::vector<double> lb, ub;                       /* +DBL_MAX / -DBL_MAX = not given */
    std::vector<int> r_type;
    std::vector<double> r_bnd;
    std::vector<std::vector<std::pair<int, double>>> rows; /* (column, value) */

    const Token &tok() const { return toks[pos]; }
    void adv() { if (pos + 1 < toks.size()) pos++; }
    [[noreturn]] void fail(const std::string &msg) const { throw SyntaxError{msg, tok().line}; }

    int find_col(const std::string &name)
    {
        auto it = col_of.find(name);
        if (it != col_of.end()) return it->second;
        int j = (int)col_name.size();
        col_of.emplace(name, j);
        col_name.push_back(name);
        coef.push_back(0.0); kind.push_back(1);
        lb.push_back(+DBL_MAX); ub.push_back(-DBL_MAX);
        return j;
    }
    bool is_sign() const { return tok().kind == T_PLUS || tok().kind == T_MINUS; }

    std::vector<std::pair<int, double>> linear_form()
    {
        std::vector<std::pair<int, double>> form;
        std::vector<char> used;
        for (;;) {
            double s = 1.0, c = 1.0;
            if (is_sign()) { s = tok().kind == T_PLUS ? 1.0 : -1.0; adv(); }
            if (tok().kind == T_NUM) { c = tok().value; adv(); }
            if (tok().kind != T_NAME) fail("missing variable name");
            int j = find_col(tok().image);
            if ((size_t)j >= used.size()) used.resize(col_name.size(), 0);
            if (used[j]) fail("multiple use of variable `" + tok().image + "' not allowed");
            used[j] = 1;
            form.emplace_back(j, s * c);
            adv();
            if (!is_sign()) break;
        }
        std::vector<std::pair<int, double>> keep;
        for (auto &e : form)
            if (e.second != 0.0) keep.push_back(e);
        return keep;
    }
    double signed_number(const char *what)
    {
        double s = 1.0;
        if (is_sign()) { s = tok().kind == T_PLUS ? 1.0 : -1.0; adv(); }
        if (tok().kind != T_NUM) fail(std::string("missing ") + what);
        double v = s * tok().value;
        adv();
        return v;
    }
    double bound_value(bool is_lower)
    {
        double s = 1.0;
        bool sgn = false;
        if (is_sign()) { s = tok().kind == T_PLUS ? 1.0 : -1.0; sgn = true; adv(); }
        if (tok().kind == T_NUM) { double v = s * tok().value; adv(); return v; }
        if (sgn && tok().kind == T_NAME) {
            std::string low = lower(tok().image);
            if (low == "infinity" || low == "inf") {
                if (is_lower && s > 0) fail("invalid use of `+inf' as lower bound");
                if (!is_lower && s < 0) fail("invalid use of `-inf' as upper bound");
                adv();
                return is_lower ? -DBL_MAX : +DBL_MAX;
            }
        }
        fail(is_lower ? "missing lower bound" : "missing upper bound");
    }

    void parse()
    {
        if (tok().kind != T_MIN && tok().kind != T_MAX) fail("`minimize' or `maximize' keyword missing");
        dir = tok().kind == T_MIN ? 1 : 2;
        adv();
        if (tok().kind == T_NAME && tok().colon) { obj_name = tok().image; adv(); adv(); }
        else obj_name = "obj";
        for (auto &e : linear_form()) coef[e.first] = e.second;
        if (tok().kind != T_ST) fail("constraints section missing");
        adv();
        for (;;) {
            int i = (int)rows.size();
            if (tok().kind == T_NAME && tok().colon) {
                if (row_of.count(tok().image)) fail("constraint `" + tok().image + "' multiply defined");
                row_of.emplace(tok().image, i);
                row_name.push_back(tok().image);
                adv(); adv();
            } else
                row_name.push_back("r." + std::to_string(tok().line));     /* "r." + csa.count: the LINE number, lib/glpcpx.js:395 */
            rows.push_back(linear_form());
            Kind sense = tok().kind;
            if (sense != T_LE && sense != T_GE && sense != T_EQ) fail("missing constraint sense");
            adv();
            {
                double sgn = 1.0;
                if (is_sign()) { sgn = tok().kind == T_PLUS ? 1.0 : -1.0; adv(); }
                if (tok().kind != T_NUM) fail("missing right-hand side");
                r_bnd.push_back(sgn * tok().value);
                if (!tok().eol) fail("invalid symbol(s) beyond right-hand side");   /* lib/glpcpx.js:431-433 */
                adv();
            }
            r_type.push_back(sense == T_LE ? 3 : (sense == T_GE ? 2 : 5)); /* GLP_UP / GLP_LO / GLP_FX */
            if (!(is_sign() || tok().kind == T_NUM || tok().kind == T_NAME)) break;
        }
        if (tok().kind == T_BOUNDS) {
            adv();
            while (is_sign() || tok().kind == T_NUM || tok().kind == T_NAME) {
                const bool lb_flag = tok().kind != T_NAME;
                double lbv = 0.0;
                if (lb_flag) {
                    lbv = bound_value(true);
                    if (tok().kind != T_LE) fail("missing `<', `<=', or `=<' after lower bound");
                    adv();
                }
                if (tok().kind != T_NAME) fail("missing variable name");
                int j = find_col(tok().image);
                if (lb_flag) lb[j] = lbv;
                adv();
                if (tok().kind == T_LE) { adv(); ub[j] = bound_value(false); }
                else if (tok().kind == T_GE) {
                    if (lb_flag) fail("invalid bound definition");
                    adv(); lb[j] = bound_value(true);
                } else if (tok().kind == T_EQ) {
                    if (lb_flag) fail("invalid bound definition");
                    adv(); lb[j] = ub[j] = signed_number("fixed value");
                } else if (tok().kind == T_NAME && lower(tok().image) == "free") {
                    if (lb_flag) fail("invalid bound definition");
                    lb[j] = -DBL_MAX; ub[j] = +DBL_MAX;
                    adv();
                } else if (!lb_flag)
                    fail("invalid bound definition");
            }
        }
        while (tok().kind == T_GEN || tok().kind == T_INT || tok().kind == T_BIN) {
            const bool binary = tok().kind == T_BIN;
            adv();
            while (tok().kind == T_NAME) {
                int j = find_col(tok().image);
                kind[j] = 2;
                if (binary) { lb[j] = 0.0; ub[j] = 1.0; }
                adv();
            }
        }
        if (tok().kind == T_END) adv();
        else if (tok().kind != T_EOF) fail("symbol " + tok().image + " in wrong position");
        if (tok().kind != T_EOF) fail("extra symbol(s) detected beyond `end'");
    }
};

template <class T> T *dup(const std::vector<T> &v)
{
    T *p = (T *)malloc((v.size() ? v.size() : 1) * sizeof(T));
    if (p && !v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}

} // namespace

extern "C" {

/* replaces glp_read_lp (lib/glpcpx.js:10-753) */
int glpb_read_lp(const char *text, long len, glpb_problem_data *out, char **names, long *names_len)
try {
    if (!text || len < 0 || !out) return GLPB_EINVAL;
    memset(out, 0, sizeof *out);
    if (names) *names = nullptr;
    if (names_len) *names_len = 0;
    Reader R;
    try {
        R.toks = tokenize(text, len);
        R.parse();
    } catch (const SyntaxError &e) {
        glpb_set_error("glp_read_lp: line %d: %s", e.line, e.msg.c_str());
        return 1;
    }
    const int m = (int)R.rows.size(), n = (int)R.col_name.size();
    std::vector<int> type(m + n), ptr(n + 1, 0), ind;
    std::vector<double> lo(m + n, 0.0), hi(m + n, 0.0), val;
    for (int i = 0; i < m; i++) { /* glp_set_row_bnds(type, rhs, rhs) keeps only the bounds the type has */
        type[i] = R.r_type[i];
        lo[i] = (type[i] == 2 || type[i] == 5) ? R.r_bnd[i] : 0.0;
        hi[i] = (type[i] == 3 || type[i] == 5) ? R.r_bnd[i] : 0.0;
    }
    for (int j = 0; j < n; j++) { /* lib/glpcpx.js:697-715 */
        double l = R.lb[j], u = R.ub[j];
        if (l == +DBL_MAX) l = 0.0;
        if (u == -DBL_MAX) u = +DBL_MAX;
        int t;
        if (l == -DBL_MAX && u == +DBL_MAX) t = 1;      /* GLP_FR */
        else if (u == +DBL_MAX) t = 2;                  /* GLP_LO */
        else if (l == -DBL_MAX) t = 3;                  /* GLP_UP */
        else if (l != u) t = 4;                         /* GLP_DB */
        else t = 5;                                     /* GLP_FX */
        type[m + j] = t;
        /* what glp_set_col_bnds keeps (lib/glpapi01.js:249-281) */
        lo[m + j] = (t == 2 || t == 4 || t == 5) ? l : 0.0;
        hi[m + j] = (t == 3 || t == 4) ? u : (t == 5 ? l : 0.0);
    }
    /* rows arrive in order: a counting pass gives the columns with ascending row indices */
    size_t nnz = 0;
    for (auto &r : R.rows) { for (auto &e : r) ptr[e.first + 1]++; nnz += r.size(); }
    if (nnz > 0x7fffffffu) { glpb_set_error("glp_read_lp: too many constraint coefficients"); return 1; }
    for (int j = 0; j < n; j++) ptr[j + 1] += ptr[j];
    ind.resize(nnz); val.resize(nnz);
    {
        std::vector<int> fill(ptr.begin(), ptr.end() - 1);
        for (int i = 0; i < m; i++)
            for (auto &e : R.rows[i]) { int p = fill[e.first]++; ind[p] = i; val[p] = e.second; }
    }
    out->m = m; out->n = n; out->nnz = (int)nnz; out->dir = R.dir; out->c0 = 0.0;
    out->type = dup(type); out->lb = dup(lo); out->ub = dup(hi);
    out->coef = dup(R.coef); out->kind = dup(R.kind);
    out->A_ptr = dup(ptr); out->A_ind = dup(ind); out->A_val = dup(val);
    if (!out->type || !out->lb || !out->ub || !out->coef || !out->kind || !out->A_ptr || !out->A_ind || !out->A_val) {
        glpb_free_problem(out);
        return GLPB_ENOMEM;
    }
    if (names) {
        /* NUL-separated: objective name, m row names, n column names */
        std::string blob = R.obj_name;
        blob.push_back('\0');
        for (auto &s : R.row_name) { blob += s; blob.push_back('\0'); }
        for (auto &s : R.col_name) { blob += s; blob.push_back('\0'); }
        char *p = (char *)malloc(blob.size() ? blob.size() : 1);
        if (!p) { glpb_free_problem(out); return GLPB_ENOMEM; }
        memcpy(p, blob.data(), blob.size());
        *names = p;
        if (names_len) *names_len = (long)blob.size();
    }
    return 0;
} catch (const std::bad_alloc &) {
    return GLPB_ENOMEM;   /* nothing crosses the C ABI as an exception */
}

void glpb_free_names(char *names) { free(names); }

} // extern "C"

/* ------------------------------------------------------------------ writer
 * glp_write_lp (lib/glpcpx.js:755-999).  The text is assembled in one buffer,
 * lines separated by '\n'; a binding replays it through the caller's callback.
 * Numbers are written the way the reference's string concatenation shows them:
 * ECMAScript Number::toString(10) -- the shortest digits that round-trip
 * (std::to_chars yields the same ones), laid out by the position n of the
 * decimal point. */
namespace {

std::string js_number(double v)
{
    if (v != v) return "NaN";
    if (std::isinf(v)) return v > 0 ? "Infinity" : "-Infinity";
    if (v == 0.0) return "0";
    if (v < 0) return "-" + js_number(-v);
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);
    std::string sci(buf, r.ptr);                       /* d[.ddd]e[+-]xx, shortest round-trip */
    const size_t epos = sci.find('e');
    std::string digits;
    for (size_t i = 0; i < epos; i++) if (sci[i] != '.') digits.push_back(sci[i]);
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    const int n = atoi(sci.c_str() + epos + 1) + 1;
    const int k = (int)digits.size();
    if (k <= n && n <= 21) return digits + std::string(n - k, '0');
    if (0 < n && n <= 21) return digits.substr(0, n) + "." + digits.substr(n);
    if (-6 < n && n <= 0) return "0." + std::string(-n, '0') + digits;
    const int e = n - 1;
    std::string tail = std::string("e") + (e > 0 ? "+" : "-") + std::to_string(e < 0 ? -e : e);
    return (k == 1 ? digits : digits.substr(0, 1) + "." + digits.substr(1)) + tail;
}

/* check_name (lib/glpcpx.js:757-766); adjust_name (:768-780) assigns into an immutable
   string and therefore changes nothing, so a name with a blank or a dash is simply invalid */
bool name_ok(const char *s)
{
    if (!s || !*s) return false;
    if (s[0] == '.' || isdigit((unsigned char)s[0])) return false;
    for (const unsigned char *p = (const unsigned char *)s; *p; p++)
        if (!(*p < 128 && isalnum(*p)) && !strchr(NAME_EXTRA, *p)) return false;
    return true;
}

struct LpText {
    std::string text, line;
    int count = 0;
    void out(const std::string &l) { text += l; text.push_back('\n'); count++; }
    void begin(const std::string &l) { line = l; }
    void add(const std::string &term)                   /* lines are broken before column 73 */
    {
        if (line.size() + term.size() > 72) { out(line); line.clear(); }
        line += term;
    }
    void end() { out(line); }
};

std::string signed_term(double v, const std::string &name)
{
    if (v == +1.0) return " + " + name;
    if (v == -1.0) return " - " + name;
    if (v > 0.0) return " + " + js_number(v) + " " + name;
    return " - " + js_number(-v) + " " + name;
}

} // namespace

extern "C" int glpb_write_lp(int m, int n, int dir, double c0, const int *type, const double *lb, const double *ub,
                             const double *coef, const int *kind, const int *col_len, const int *R_ptr,
                             const int *R_ind, const double *R_val, const char *prob_name, const char *names,
                             char **text, long *text_len, int *lines)
{
    if (m < 0 || n < 0 || !text || !(dir == 1 || dir == 2)) return GLPB_EINVAL;
    if (m > 0 && n > 0 && (!type || !lb || !ub || !coef || !col_len || !R_ptr || (R_ptr[m] > 0 && (!R_ind || !R_val))))
        return GLPB_EINVAL;
    try {
        /* names block as glpb_read_lp returns it: objective, m rows, n columns, NUL-terminated; an empty
           string stands for "no name" */
        std::vector<const char *> nm(1 + m + n, nullptr);
        if (names) {
            const char *p = names;
            for (int k = 0; k < 1 + m + n; k++) { nm[k] = p; p += strlen(p) + 1; }
        }
        auto row_name = [&](int i) -> std::string {     /* i = 0: the objective */
            if (name_ok(nm[i])) return nm[i];
            return i == 0 ? std::string("obj") : "r_" + std::to_string(i);
        };
        auto col_name = [&](int j) -> std::string {
            if (name_ok(nm[m + j])) return nm[m + j];
            return "x_" + std::to_string(j);
        };
        LpText T;
        T.out(std::string("\\* Problem: ") + (prob_name ? prob_name : "Unknown") + " *\\");
        T.out("");
        if (!(m > 0 && n > 0)) {
            T.out("\\* WARNING: PROBLEM HAS NO ROWS/COLUMNS *\\");
            T.out("");
        } else {
            T.out(dir == 1 ? "Minimize" : "Maximize");
            T.begin(" " + row_name(0) + ":");
            int terms = 0;
            for (int j = 1; j <= n; j++) {
                const double c = coef[j - 1];
                if (c != 0.0 || col_len[j - 1] == 0) {
                    terms++;
                    T.add(c == 0.0 ? " + 0 " + col_name(j) : signed_term(c, col_name(j)));
                }
            }
            if (terms == 0) T.line += " 0 " + col_name(1);
            T.end();
            if (c0 != 0.0) T.out("\\* constant term = " + js_number(c0) + " *\\");
            T.out("");
            T.out("Subject To");
            for (int i = 1; i <= m; i++) {
                const int t = type[i - 1];
                if (t == 1) continue;                    /* GLP_FR */
                T.begin(" " + row_name(i) + ":");
                for (int e = R_ptr[i - 1]; e < R_ptr[i]; e++) {
                    if (R_ind[e] < 0 || R_ind[e] >= n) return GLPB_EINVAL;
                    T.add(signed_term(R_val[e], col_name(R_ind[e] + 1)));
                }
                if (t == 4) T.add(" - ~r_" + std::to_string(i));
                else if (R_ptr[i] == R_ptr[i - 1]) T.line += " 0 " + col_name(1);
                if (t == 2) T.add(" >= " + js_number(lb[i - 1]));
                else if (t == 3) T.add(" <= " + js_number(ub[i - 1]));
                else T.add(" = " + js_number(lb[i - 1]));
                T.end();
            }
            T.out("");
            bool flag = false;
            for (int i = 1; i <= m; i++) {
                if (type[i - 1] != 4) continue;
                if (!flag) { T.out("Bounds"); flag = true; }
                T.out(" 0 <= ~r_" + std::to_string(i) + " <= " + js_number(ub[i - 1] - lb[i - 1]));
            }
            for (int j = 1; j <= n; j++) {
                const int k = m + j - 1, t = type[k];
                if (t == 2 && lb[k] == 0.0) continue;
                if (!flag) { T.out("Bounds"); flag = true; }
                const std::string name = col_name(j);
                if (t == 1) T.out(" " + name + " free");
                else if (t == 2) T.out(" " + name + " >= " + js_number(lb[k]));
                else if (t == 3) T.out(" -Inf <= " + name + " <= " + js_number(ub[k]));
                else if (t == 4) T.out(" " + js_number(lb[k]) + " <= " + name + " <= " + js_number(ub[k]));
                else T.out(" " + name + " = " + js_number(lb[k]));
            }
            if (flag) { T.text.push_back('\n'); }
            T.count++;                                   /* the reference counts this line even when it does not write it (:973) */
            flag = false;
            for (int j = 1; j <= n; j++) {
                if (!kind || kind[j - 1] == 1) continue; /* GLP_CV */
                if (!flag) { T.out("Generals"); flag = true; }
                T.out(" " + col_name(j));
            }
            if (flag) T.out("");
        }
        T.out("End");
        char *p = (char *)malloc(T.text.size() + 1);
        if (!p) return GLPB_ENOMEM;
        memcpy(p, T.text.data(), T.text.size());
        p[T.text.size()] = '\0';
        *text = p;
        if (text_len) *text_len = (long)T.text.size();
        if (lines) *lines = T.count;
        return 0;
    } catch (const std::bad_alloc &) {
        return GLPB_ENOMEM;
    }
}
