/* glpb200_napi.c -- thin N-API addon over libglpb200.so (include/glpb200.h).
 *
 * Not loadable in this repository's image (no Node.js, no node_api.h); it is the
 * binding a glpk.js maintainer would add.  It is written against the stable
 * N-API C signatures only; tests/ compile it against a declarations-only stub of
 * node_api.h (tests/napi_stub) so that it at least stays well-formed C.  Its only
 * logic is validation: every typed array is checked for element type and length
 * against m, n, nnz BEFORE a pointer crosses the C ABI (a short buffer would be a
 * heap overflow inside the library), arguments of the wrong kind throw TypeError.
 *
 *   build:  cc -shared -fPIC glpb200_napi.c -I../include -I$NODE/include/node \
 *              -L../glpk.js_b200 -lglpb200 -o glpb200.node
 */
#include <node_api.h>
#include <stdlib.h>
#include <string.h>
#include "glpb200.h"

#define NAPI_OK(call) do { if ((call) != napi_ok) { napi_throw_error(env, NULL, #call); return NULL; } } while (0)

/* A typed array argument of the given element type with at least `need` elements.
   Returns its data pointer, or NULL after throwing a TypeError: nothing reaches the
   C ABI unchecked (the library reads type/lb/ub[m+n], A_ptr[n+1], ... and writes
   stat/prim/dual[m+n] into the caller's buffers). */
static void *typed_n(napi_env env, napi_value v, napi_typedarray_type want, size_t need, const char *what, size_t *len_out)
{
    napi_typedarray_type t; void *data = NULL; napi_value ab; size_t off = 0, len = 0;
    bool is = false;
    if (napi_is_typedarray(env, v, &is) != napi_ok || !is ||
        napi_get_typedarray_info(env, v, &t, &len, &data, &ab, &off) != napi_ok || t != want || len < need ||
        (data == NULL && need > 0)) {
        napi_throw_type_error(env, NULL, what);
        return NULL;
    }
    if (len_out) *len_out = len;
    return data;
}
#define I32(v, need, what) ((int *)typed_n(env, (v), napi_int32_array, (need), "glpb200: " what ": Int32Array too short or of the wrong type", NULL))
#define F64(v, need, what) ((double *)typed_n(env, (v), napi_float64_array, (need), "glpb200: " what ": Float64Array too short or of the wrong type", NULL))

static bool is_nullish(napi_env env, napi_value v)
{
    napi_valuetype t;
    return napi_typeof(env, v, &t) == napi_ok && (t == napi_undefined || t == napi_null);
}

static void finalize_handle(napi_env env, void *data, void *hint) { (void)env; (void)hint; glpb_destroy((glpb_prob *)data); }

/* the handle and its dimensions (kept next to it so that every later call can size-check its arrays) */
typedef struct { glpb_prob *P; int m, n; } shim_handle;
static void finalize_shim(napi_env env, void *data, void *hint) { shim_handle *h = (shim_handle *)data; (void)env; (void)hint; glpb_destroy(h->P); free(h); }

static shim_handle *handle(napi_env env, napi_value v)
{
    void *p = NULL;
    if (napi_get_value_external(env, v, &p) != napi_ok || !p) { napi_throw_type_error(env, NULL, "glpb200: not a problem handle"); return NULL; }
    return (shim_handle *)p;
}

static napi_value ret_int(napi_env env, int v) { napi_value r; napi_create_int32(env, v, &r); return r; }

/* create(m, n, dir, c0, Int32 type, F64 lb, F64 ub, F64 coef, Int32 kind|null, F64 rii|null, F64 sjj|null,
          Int32 A_ptr, Int32 A_ind, F64 A_val, device) -> external handle */
static napi_value Create(napi_env env, napi_callback_info info)
{
    size_t argc = 15; napi_value a[15];
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    if (argc < 15) { napi_throw_type_error(env, NULL, "glpb200.create: 15 arguments expected"); return NULL; }
    int32_t m, n, dir, device; double c0;
    NAPI_OK(napi_get_value_int32(env, a[0], &m));
    NAPI_OK(napi_get_value_int32(env, a[1], &n));
    NAPI_OK(napi_get_value_int32(env, a[2], &dir));
    NAPI_OK(napi_get_value_double(env, a[3], &c0));
    NAPI_OK(napi_get_value_int32(env, a[14], &device));
    if (m < 1 || n < 1) { napi_throw_range_error(env, NULL, "glpb200.create: m, n must be positive"); return NULL; }
    const size_t mn = (size_t)m + (size_t)n;
    int *type = I32(a[4], mn, "type"); if (!type) return NULL;
    double *lb = F64(a[5], mn, "lb"); if (!lb) return NULL;
    double *ub = F64(a[6], mn, "ub"); if (!ub) return NULL;
    double *coef = F64(a[7], (size_t)n, "coef"); if (!coef) return NULL;
    int *kind = NULL; double *rii = NULL, *sjj = NULL;
    if (!is_nullish(env, a[8])) { kind = I32(a[8], (size_t)n, "kind"); if (!kind) return NULL; }
    if (!is_nullish(env, a[9])) { rii = F64(a[9], (size_t)m, "rii"); if (!rii) return NULL; }
    if (!is_nullish(env, a[10])) { sjj = F64(a[10], (size_t)n, "sjj"); if (!sjj) return NULL; }
    int *ptr = I32(a[11], (size_t)n + 1, "A_ptr"); if (!ptr) return NULL;
    /* A_ptr must be a monotone 0-based column pointer; its last entry is nnz */
    if (ptr[0] != 0) { napi_throw_range_error(env, NULL, "glpb200.create: A_ptr[0] must be 0"); return NULL; }
    for (int j = 0; j < n; j++)
        if (ptr[j + 1] < ptr[j]) { napi_throw_range_error(env, NULL, "glpb200.create: A_ptr is not monotone"); return NULL; }
    const size_t nnz = (size_t)ptr[n];
    int *ind = I32(a[12], nnz, "A_ind"); if (!ind && nnz) return NULL;
    double *val = F64(a[13], nnz, "A_val"); if (!val && nnz) return NULL;
    for (size_t t = 0; t < nnz; t++)
        if (ind[t] < 0 || ind[t] >= m) { napi_throw_range_error(env, NULL, "glpb200.create: row index out of range"); return NULL; }
    shim_handle *h = (shim_handle *)calloc(1, sizeof *h);
    if (!h) { napi_throw_error(env, NULL, "glpb200.create: out of memory"); return NULL; }
    h->P = glpb_create(m, n, (int)nnz, dir, c0, type, lb, ub, coef, kind, rii, sjj, ptr, ind, val, device);
    h->m = m; h->n = n;
    if (!h->P) { free(h); napi_throw_error(env, NULL, glpb_last_error()); return NULL; }
    napi_value ext;
    if (napi_create_external(env, h, finalize_shim, NULL, &ext) != napi_ok) { finalize_shim(env, h, NULL); napi_throw_error(env, NULL, "napi_create_external"); return NULL; }
    return ext;
}

/* setBasis(h, Int32 stat[m+n]) ; setBounds(h, Int32 k[cnt], Int32 type[cnt], F64 lb[cnt], F64 ub[cnt]) */
static napi_value SetBasis(napi_env env, napi_callback_info info)
{
    size_t argc = 2; napi_value a[2];
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    shim_handle *h = handle(env, a[0]); if (!h) return NULL;
    int *stat = I32(a[1], (size_t)h->m + h->n, "stat"); if (!stat) return NULL;
    return ret_int(env, glpb_set_basis(h->P, stat));
}

static napi_value SetBounds(napi_env env, napi_callback_info info)
{
    size_t argc = 5, cnt = 0; napi_value a[5];
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    shim_handle *h = handle(env, a[0]); if (!h) return NULL;
    int *k = (int *)typed_n(env, a[1], napi_int32_array, 0, "glpb200: k: Int32Array expected", &cnt);
    if (!k && cnt) return NULL;
    if (cnt == 0) return ret_int(env, 0);
    int *type = I32(a[2], cnt, "type"); if (!type) return NULL;
    double *lb = F64(a[3], cnt, "lb"); if (!lb) return NULL;
    double *ub = F64(a[4], cnt, "ub"); if (!ub) return NULL;
    return ret_int(env, glpb_set_bounds(h->P, (int)cnt, k, type, lb, ub));
}

/* simplex(h, Int32 ip[9], F64 dp[5]) : the SMCP fields packed by the facade */
static napi_value Simplex(napi_env env, napi_callback_info info)
{
    size_t argc = 3; napi_value a[3];
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    shim_handle *h = handle(env, a[0]); if (!h) return NULL;
    int *ip = I32(a[1], 9, "smcp integers"); if (!ip) return NULL;
    double *dp = F64(a[2], 5, "smcp doubles"); if (!dp) return NULL;
    glpb_smcp s = { ip[0], ip[1], ip[2], ip[3], dp[0], dp[1], dp[2], dp[3], dp[4], ip[4], ip[5], ip[6], ip[7], ip[8] };
    int rc = glpb_simplex(h->P, &s);
    if (rc < 0) { napi_throw_error(env, NULL, glpb_last_error()); return NULL; }
    return ret_int(env, rc);
}

/* intopt(h, Int32 ip[8], F64 dp[3]) */
static napi_value Intopt(napi_env env, napi_callback_info info)
{
    size_t argc = 3; napi_value a[3];
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    shim_handle *h = handle(env, a[0]); if (!h) return NULL;
    int *ip = I32(a[1], 8, "iocp integers"); if (!ip) return NULL;
    double *dp = F64(a[2], 3, "iocp doubles"); if (!dp) return NULL;
    glpb_iocp s = { ip[0], ip[1], ip[2], dp[0], dp[1], ip[3], ip[4], ip[5], ip[6], dp[2], ip[7], -1 };
    int rc = glpb_intopt(h->P, &s);
    if (rc < 0) { napi_throw_error(env, NULL, glpb_last_error()); return NULL; }
    return ret_int(env, rc);
}

/* getSolution(h, Int32 stat[m+n], F64 prim[m+n], F64 dual[m+n], Int32 head[m], Int32 ints[4], F64 obj[1]) */
static napi_value GetSolution(napi_env env, napi_callback_info info)
{
    size_t argc = 7; napi_value a[7];
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    shim_handle *h = handle(env, a[0]); if (!h) return NULL;
    const size_t mn = (size_t)h->m + h->n;
    int *stat = I32(a[1], mn, "stat"); if (!stat) return NULL;
    double *prim = F64(a[2], mn, "prim"); if (!prim) return NULL;
    double *dual = F64(a[3], mn, "dual"); if (!dual) return NULL;
    int *head = I32(a[4], (size_t)h->m, "head"); if (!head) return NULL;
    int *ints = I32(a[5], 4, "ints"); if (!ints) return NULL;
    double *obj = F64(a[6], 1, "obj"); if (!obj) return NULL;
    return ret_int(env, glpb_get_solution(h->P, stat, prim, dual, head, &ints[0], &ints[1], obj, &ints[2], &ints[3]));
}

/* getMip(h, Int32 st[1], F64 obj[1], F64 mipx[m+n]) */
static napi_value GetMip(napi_env env, napi_callback_info info)
{
    size_t argc = 4; napi_value a[4];
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    shim_handle *h = handle(env, a[0]); if (!h) return NULL;
    int *st = I32(a[1], 1, "mip status"); if (!st) return NULL;
    double *obj = F64(a[2], 1, "mip objective"); if (!obj) return NULL;
    double *x = F64(a[3], (size_t)h->m + h->n, "mipx"); if (!x) return NULL;
    return ret_int(env, glpb_get_mip(h->P, st, obj, x, NULL));
}

/* scaleProb(m, n, Int32 A_ptr, Int32 A_ind, F64 A_val, flags, F64 rii[m], F64 sjj[n], F64 report[13])
   -- glp_scale_prob (lib/glpscl.js) on flat arrays; host only */
static napi_value ScaleProb(napi_env env, napi_callback_info info)
{
    size_t argc = 9, len; napi_value a[9]; int m, n, flags;
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    NAPI_OK(napi_get_value_int32(env, a[0], &m)); NAPI_OK(napi_get_value_int32(env, a[1], &n));
    NAPI_OK(napi_get_value_int32(env, a[5], &flags));
    if (m < 1 || n < 1) { napi_throw_range_error(env, NULL, "glpb200.scaleProb: m, n must be positive"); return NULL; }
    int *ptr = I32(a[2], (size_t)n + 1, "A_ptr"); if (!ptr) return NULL;
    if (ptr[0] != 0 || ptr[n] < 0) { napi_throw_range_error(env, NULL, "glpb200.scaleProb: bad A_ptr"); return NULL; }
    int *ind = I32(a[3], (size_t)ptr[n], "A_ind"); if (!ind && ptr[n]) return NULL;
    double *val = F64(a[4], (size_t)ptr[n], "A_val"); if (!val && ptr[n]) return NULL;
    double *rii = F64(a[6], (size_t)m, "rii"); if (!rii) return NULL;
    double *sjj = F64(a[7], (size_t)n, "sjj"); if (!sjj) return NULL;
    double *rep = NULL;
    if (!is_nullish(env, a[8])) { rep = F64(a[8], 13, "report"); if (!rep) return NULL; }
    (void)len;
    return ret_int(env, glpb_scale_prob(m, n, ptr, ind, val, flags, rii, sjj, rep));
}

/* advBasis(m, n, Int32 A_ptr, Int32 A_ind, Int32 R_ptr, Int32 R_ind, Int32 type, F64 lb, F64 ub,
            Int32 stat[m+n], Int32 size[1]) -- glp_adv_basis (lib/glpini01.js); host only */
static napi_value AdvBasis(napi_env env, napi_callback_info info)
{
    size_t argc = 11, len; napi_value a[11]; int m, n;
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    NAPI_OK(napi_get_value_int32(env, a[0], &m)); NAPI_OK(napi_get_value_int32(env, a[1], &n));
    if (m < 1 || n < 1) { napi_throw_range_error(env, NULL, "glpb200.advBasis: m, n must be positive"); return NULL; }
    const size_t mn = (size_t)m + (size_t)n;
    int *ptr = I32(a[2], (size_t)n + 1, "A_ptr"); if (!ptr) return NULL;
    int *rptr = I32(a[4], (size_t)m + 1, "R_ptr"); if (!rptr) return NULL;
    if (ptr[0] != 0 || rptr[0] != 0 || ptr[n] < 0 || rptr[m] != ptr[n]) { napi_throw_range_error(env, NULL, "glpb200.advBasis: bad A_ptr / R_ptr"); return NULL; }
    int *ind = I32(a[3], (size_t)ptr[n], "A_ind"); if (!ind && ptr[n]) return NULL;
    int *rind = I32(a[5], (size_t)ptr[n], "R_ind"); if (!rind && ptr[n]) return NULL;
    int *type = I32(a[6], mn, "type"); if (!type) return NULL;
    double *lb = F64(a[7], mn, "lb"); if (!lb) return NULL;
    double *ub = F64(a[8], mn, "ub"); if (!ub) return NULL;
    int *stat = I32(a[9], mn, "stat"); if (!stat) return NULL;
    int *size = NULL;
    if (!is_nullish(env, a[10])) { size = I32(a[10], 1, "size"); if (!size) return NULL; }
    (void)len;
    return ret_int(env, glpb_adv_basis(m, n, ptr, ind, rptr, rind, type, lb, ub, stat, size));
}

/* readLp(text) -> { m, n, dir, type, lb, ub, coef, kind, ptr, ind, val, names } -- glp_read_lp
   (lib/glpcpx.js) on a string held in memory; throws the reader's message on a syntax error */
static napi_value ReadLp(napi_env env, napi_callback_info info)
{
    size_t argc = 1, len = 0; napi_value a[1], r, v; glpb_problem_data d; char *names = NULL; long nlen = 0; void *p;
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    NAPI_OK(napi_get_value_string_utf8(env, a[0], NULL, 0, &len));
    char *text = (char *)malloc(len + 1);
    if (!text) { napi_throw_error(env, NULL, "glpb200.readLp: out of memory"); return NULL; }
    if (napi_get_value_string_utf8(env, a[0], text, len + 1, &len) != napi_ok) {
        free(text);                                  /* no leak on the error path */
        napi_throw_error(env, NULL, "glpb200.readLp: argument is not a string");
        return NULL;
    }
    int rc = glpb_read_lp(text, (long)len, &d, &names, &nlen);
    free(text);
    if (rc != 0) { napi_throw_error(env, NULL, glpb_last_error()); return NULL; }
    if (napi_create_object(env, &r) != napi_ok) { glpb_free_names(names); glpb_free_problem(&d); napi_throw_error(env, NULL, "napi_create_object"); return NULL; }
#define PUT_INT(key, x) do { napi_create_int32(env, (x), &v); napi_set_named_property(env, r, key, v); } while (0)
#define PUT_ARR(key, type, src, count, bytes) do { napi_value ab; napi_create_arraybuffer(env, (count) * (bytes), &p, &ab); \
        memcpy(p, (src), (count) * (bytes)); napi_create_typedarray(env, type, (count), ab, 0, &v); \
        napi_set_named_property(env, r, key, v); } while (0)
    PUT_INT("m", d.m); PUT_INT("n", d.n); PUT_INT("dir", d.dir);
    PUT_ARR("type", napi_int32_array, d.type, (size_t)(d.m + d.n), 4); PUT_ARR("lb", napi_float64_array, d.lb, (size_t)(d.m + d.n), 8);
    PUT_ARR("ub", napi_float64_array, d.ub, (size_t)(d.m + d.n), 8); PUT_ARR("coef", napi_float64_array, d.coef, (size_t)d.n, 8);
    PUT_ARR("kind", napi_int32_array, d.kind, (size_t)d.n, 4); PUT_ARR("ptr", napi_int32_array, d.A_ptr, (size_t)d.n + 1, 4);
    PUT_ARR("ind", napi_int32_array, d.A_ind, (size_t)d.nnz, 4); PUT_ARR("val", napi_float64_array, d.A_val, (size_t)d.nnz, 8);
    napi_create_string_utf8(env, names, (size_t)nlen, &v);       /* NUL-separated: obj, rows, columns */
    napi_set_named_property(env, r, "names", v);
    glpb_free_names(names); glpb_free_problem(&d);
    return r;
}

/* writeLp(m, n, dir, c0, Int32 type, F64 lb, F64 ub, F64 coef, Int32 kind|null, Int32 col_len, Int32 R_ptr, Int32 R_ind,
           F64 R_val, probName|null, names|null) -> { text, lines } -- glp_write_lp (lib/glpcpx.js:755-999); `names` is
   one string of 1+m+n NUL-terminated names (objective, rows, columns; "" = none), rows in list order */
static napi_value WriteLp(napi_env env, napi_callback_info info)
{
    size_t argc = 15, len = 0; napi_value a[15], r, v;
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    if (argc < 15) { napi_throw_type_error(env, NULL, "glpb200.writeLp: 15 arguments expected"); return NULL; }
    int32_t m, n, dir; double c0;
    NAPI_OK(napi_get_value_int32(env, a[0], &m));
    NAPI_OK(napi_get_value_int32(env, a[1], &n));
    NAPI_OK(napi_get_value_int32(env, a[2], &dir));
    NAPI_OK(napi_get_value_double(env, a[3], &c0));
    if (m < 0 || n < 0) { napi_throw_range_error(env, NULL, "glpb200.writeLp: m, n must not be negative"); return NULL; }
    const size_t mn = (size_t)m + (size_t)n;
    int *type = I32(a[4], mn, "type"); if (!type && mn) return NULL;
    double *lb = F64(a[5], mn, "lb"); if (!lb && mn) return NULL;
    double *ub = F64(a[6], mn, "ub"); if (!ub && mn) return NULL;
    double *coef = F64(a[7], (size_t)n, "coef"); if (!coef && n) return NULL;
    int *kind = NULL;
    if (!is_nullish(env, a[8])) { kind = I32(a[8], (size_t)n, "kind"); if (!kind && n) return NULL; }
    int *col_len = I32(a[9], (size_t)n, "col_len"); if (!col_len && n) return NULL;
    int *rptr = I32(a[10], (size_t)m + 1, "R_ptr"); if (!rptr) return NULL;
    if (rptr[m] < 0) { napi_throw_range_error(env, NULL, "glpb200.writeLp: R_ptr[m] < 0"); return NULL; }
    int *rind = I32(a[11], (size_t)rptr[m], "R_ind"); if (!rind && rptr[m]) return NULL;
    double *rval = F64(a[12], (size_t)rptr[m], "R_val"); if (!rval && rptr[m]) return NULL;
    char *pname = NULL, *names = NULL, *text = NULL; long tlen = 0; int lines = 0;
    if (!is_nullish(env, a[13])) {
        NAPI_OK(napi_get_value_string_utf8(env, a[13], NULL, 0, &len));
        if (!(pname = (char *)malloc(len + 1))) { napi_throw_error(env, NULL, "glpb200.writeLp: out of memory"); return NULL; }
        napi_get_value_string_utf8(env, a[13], pname, len + 1, &len);
    }
    if (!is_nullish(env, a[14])) {
        NAPI_OK(napi_get_value_string_utf8(env, a[14], NULL, 0, &len));
        if (!(names = (char *)malloc(len + 2))) { free(pname); napi_throw_error(env, NULL, "glpb200.writeLp: out of memory"); return NULL; }
        napi_get_value_string_utf8(env, a[14], names, len + 1, &len);
        names[len + 1] = '\0';
        size_t count = 0;                            /* the block must hold 1+m+n terminated names */
        for (size_t i = 0; i < len; i++) count += names[i] == '\0';
        if (count < 1 + mn) { free(pname); free(names); napi_throw_range_error(env, NULL, "glpb200.writeLp: names block too short"); return NULL; }
    }
    int rc = glpb_write_lp(m, n, dir, c0, type, lb, ub, coef, kind, col_len, rptr, rind, rval, pname, names, &text, &tlen, &lines);
    free(pname); free(names);
    if (rc != 0) { napi_throw_range_error(env, NULL, "glpb200.writeLp: invalid problem data"); return NULL; }
    if (napi_create_object(env, &r) != napi_ok) { glpb_free_names(text); napi_throw_error(env, NULL, "napi_create_object"); return NULL; }
    napi_create_string_utf8(env, text, (size_t)tlen, &v); napi_set_named_property(env, r, "text", v);
    napi_create_int32(env, lines, &v); napi_set_named_property(env, r, "lines", v);
    glpb_free_names(text);
    return r;
}

/* ---- presolver workspace (glpb_npp_*): the handle remembers the sizes every later array is checked against ---- */
typedef struct { glpb_npp *npp; int om, on, sol, rm, rn, rnz, loaded, built; } shim_npp;
static void finalize_npp(napi_env env, void *data, void *hint) { shim_npp *h = (shim_npp *)data; (void)env; (void)hint; glpb_npp_destroy(h->npp); free(h); }

static shim_npp *npp_handle(napi_env env, napi_value v)
{
    void *p = NULL;
    if (napi_get_value_external(env, v, &p) != napi_ok || !p) { napi_throw_type_error(env, NULL, "glpb200: not a presolver workspace"); return NULL; }
    return (shim_npp *)p;
}

/* nppCreate() -> external workspace (npp_create_wksp, lib/glpnpp01.js:2-22); freed by the finalizer */
static napi_value NppCreate(napi_env env, napi_callback_info info)
{
    napi_value r; (void)info;
    shim_npp *h = (shim_npp *)calloc(1, sizeof *h);
    if (!h || !(h->npp = glpb_npp_create())) { free(h); napi_throw_error(env, NULL, "glpb200.nppCreate: out of memory"); return NULL; }
    if (napi_create_external(env, h, finalize_npp, NULL, &r) != napi_ok) { glpb_npp_destroy(h->npp); free(h); napi_throw_error(env, NULL, "napi_create_external"); return NULL; }
    return r;
}

/* nppLoadProb(npp, m, n, dir, c0, Int32 type, F64 lb, F64 ub, F64 coef, Int32 kind|null, Int32 A_ptr, Int32 A_ind, F64 A_val, sol)
   (npp_load_prob(npp, P, GLP_OFF, sol, GLP_OFF), lib/glpnpp01.js:262-394; columns in list order, unscaled) */
static napi_value NppLoadProb(napi_env env, napi_callback_info info)
{
    size_t argc = 14; napi_value a[14];
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    if (argc < 14) { napi_throw_type_error(env, NULL, "glpb200.nppLoadProb: 14 arguments expected"); return NULL; }
    shim_npp *h = npp_handle(env, a[0]); if (!h) return NULL;
    int32_t m, n, dir, sol; double c0;
    NAPI_OK(napi_get_value_int32(env, a[1], &m));
    NAPI_OK(napi_get_value_int32(env, a[2], &n));
    NAPI_OK(napi_get_value_int32(env, a[3], &dir));
    NAPI_OK(napi_get_value_double(env, a[4], &c0));
    NAPI_OK(napi_get_value_int32(env, a[13], &sol));
    if (m < 0 || n < 0 || h->loaded) { napi_throw_range_error(env, NULL, "glpb200.nppLoadProb: bad sizes or workspace already loaded"); return NULL; }
    const size_t mn = (size_t)m + (size_t)n;
    int *type = I32(a[5], mn, "type"); if (!type && mn) return NULL;
    double *lb = F64(a[6], mn, "lb"); if (!lb && mn) return NULL;
    double *ub = F64(a[7], mn, "ub"); if (!ub && mn) return NULL;
    double *coef = F64(a[8], (size_t)n, "coef"); if (!coef && n) return NULL;
    int *kind = NULL;
    if (!is_nullish(env, a[9])) { kind = I32(a[9], (size_t)n, "kind"); if (!kind && n) return NULL; }
    int *ptr = I32(a[10], (size_t)n + 1, "A_ptr"); if (!ptr) return NULL;
    if (ptr[n] < 0) { napi_throw_range_error(env, NULL, "glpb200.nppLoadProb: A_ptr[n] < 0"); return NULL; }
    int *ind = I32(a[11], (size_t)ptr[n], "A_ind"); if (!ind && ptr[n]) return NULL;
    double *val = F64(a[12], (size_t)ptr[n], "A_val"); if (!val && ptr[n]) return NULL;
    int rc = glpb_npp_load_prob(h->npp, m, n, dir, c0, type, lb, ub, coef, kind, ptr, ind, val, sol);
    if (rc != 0) { napi_throw_range_error(env, NULL, "glpb200.nppLoadProb: invalid problem data"); return NULL; }
    h->om = m; h->on = n; h->sol = sol; h->loaded = 1;
    return ret_int(env, 0);
}

/* nppSimplex(npp) / nppInteger(npp, binarize) -> 0 | GLP_ENOPFS | GLP_ENODFS (lib/glpnpp05.js:430-521) */
static napi_value NppSimplex(napi_env env, napi_callback_info info)
{
    size_t argc = 1; napi_value a[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    shim_npp *h = argc >= 1 ? npp_handle(env, a[0]) : NULL; if (!h) return NULL;
    int rc = glpb_npp_simplex(h->npp);
    if (rc < 0) { napi_throw_error(env, NULL, "glpb200.nppSimplex: workspace not loaded for a basic solution"); return NULL; }
    return ret_int(env, rc);
}

static napi_value NppInteger(napi_env env, napi_callback_info info)
{
    size_t argc = 2; napi_value a[2]; int32_t binarize = 0;
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    shim_npp *h = argc >= 1 ? npp_handle(env, a[0]) : NULL; if (!h) return NULL;
    if (argc >= 2 && !is_nullish(env, a[1])) NAPI_OK(napi_get_value_int32(env, a[1], &binarize));
    int rc = glpb_npp_integer(h->npp, binarize);
    if (rc < 0) { napi_throw_error(env, NULL, "glpb200.nppInteger: workspace not loaded for a MIP solution"); return NULL; }
    return ret_int(env, rc);
}

/* nppBuildProb(npp) -> { m, n, c0, type, lb, ub, coef, kind, ptr, ind, val, rowRef, colRef } (npp_build_prob,
   lib/glpnpp01.js:396-472): the reduced problem; the arrays are allocated here with the sizes the library reports */
static napi_value NppBuildProb(napi_env env, napi_callback_info info)
{
    size_t argc = 1; napi_value a[1], r, v, ab; void *p;
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    shim_npp *h = argc >= 1 ? npp_handle(env, a[0]) : NULL; if (!h) return NULL;
    if (!h->loaded || h->built) { napi_throw_error(env, NULL, "glpb200.nppBuildProb: workspace not loaded or already built"); return NULL; }
    int m = 0, n = 0, nz = 0; double c0 = 0.0;
    if (glpb_npp_get_size(h->npp, &m, &n, &nz) != 0) { napi_throw_error(env, NULL, "glpb_npp_get_size"); return NULL; }
    NAPI_OK(napi_create_object(env, &r));
    void *buf[10]; const size_t cnt[10] = { (size_t)m + n, (size_t)m + n, (size_t)m + n, (size_t)n, (size_t)n, (size_t)n + 1, (size_t)nz, (size_t)nz, (size_t)m, (size_t)n };
    static const char *key[10] = { "type", "lb", "ub", "coef", "kind", "ptr", "ind", "val", "rowRef", "colRef" };
    static const int is_f64[10] = { 0, 1, 1, 1, 0, 0, 0, 1, 0, 0 };
    for (int k = 0; k < 10; k++) {
        NAPI_OK(napi_create_arraybuffer(env, cnt[k] * (is_f64[k] ? 8 : 4), &p, &ab));
        NAPI_OK(napi_create_typedarray(env, is_f64[k] ? napi_float64_array : napi_int32_array, cnt[k], ab, 0, &v));
        NAPI_OK(napi_set_named_property(env, r, key[k], v));
        buf[k] = p;
    }
    int rc = glpb_npp_build_prob(h->npp, &c0, (int *)buf[0], (double *)buf[1], (double *)buf[2], (double *)buf[3], (int *)buf[4],
                                 (int *)buf[5], (int *)buf[6], (double *)buf[7], (int *)buf[8], (int *)buf[9]);
    if (rc != 0) { napi_throw_error(env, NULL, "glpb_npp_build_prob"); return NULL; }
    h->rm = m; h->rn = n; h->rnz = nz; h->built = 1;
    napi_create_int32(env, m, &v); napi_set_named_property(env, r, "m", v);
    napi_create_int32(env, n, &v); napi_set_named_property(env, r, "n", v);
    napi_create_double(env, c0, &v); napi_set_named_property(env, r, "c0", v);
    return r;
}

/* nppPostprocess(npp, Int32 r_stat|null, F64 r_dual|null, Int32 c_stat|null, F64 c_value) ->
   { rowStat, rowDual, colStat, colValue } of the ORIGINAL problem (npp_postprocess, lib/glpnpp01.js:474-570);
   inputs sized by the REDUCED problem, outputs by the original one, both known from the handle */
static napi_value NppPostprocess(napi_env env, napi_callback_info info)
{
    size_t argc = 5; napi_value a[5], r, v, ab; void *p;
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    if (argc < 5) { napi_throw_type_error(env, NULL, "glpb200.nppPostprocess: 5 arguments expected"); return NULL; }
    shim_npp *h = npp_handle(env, a[0]); if (!h) return NULL;
    if (!h->built) { napi_throw_error(env, NULL, "glpb200.nppPostprocess: reduced problem not built"); return NULL; }
    const int basic = h->sol == 1;   /* GLP_SOL */
    int *rs = NULL, *cs = NULL; double *rd = NULL;
    if (basic) {
        rs = I32(a[1], (size_t)h->rm, "r_stat"); if (!rs && h->rm) return NULL;
        rd = F64(a[2], (size_t)h->rm, "r_dual"); if (!rd && h->rm) return NULL;
        cs = I32(a[3], (size_t)h->rn, "c_stat"); if (!cs && h->rn) return NULL;
    }
    double *cv = F64(a[4], (size_t)h->rn, "c_value"); if (!cv && h->rn) return NULL;
    NAPI_OK(napi_create_object(env, &r));
    void *out[4] = { NULL, NULL, NULL, NULL };
    const size_t cnt[4] = { (size_t)h->om, (size_t)h->om, (size_t)h->on, (size_t)h->on };
    static const char *key[4] = { "rowStat", "rowDual", "colStat", "colValue" };
    static const int is_f64[4] = { 0, 1, 0, 1 };
    for (int k = 0; k < 4; k++) {
        if (!basic && k < 3) continue;
        NAPI_OK(napi_create_arraybuffer(env, cnt[k] * (is_f64[k] ? 8 : 4), &p, &ab));
        NAPI_OK(napi_create_typedarray(env, is_f64[k] ? napi_float64_array : napi_int32_array, cnt[k], ab, 0, &v));
        NAPI_OK(napi_set_named_property(env, r, key[k], v));
        out[k] = p;
    }
    int rc = glpb_npp_postprocess(h->npp, rs, rd, cs, cv, (int *)out[0], (double *)out[1], (int *)out[2], (double *)out[3]);
    if (rc != 0) { napi_throw_error(env, NULL, "glpb200.nppPostprocess: solution cannot be recovered"); return NULL; }   /* the reference's xassert */
    return r;
}

static napi_value Init(napi_env env, napi_value exports)
{
    napi_property_descriptor d[] = {
        { "create", 0, Create, 0, 0, 0, napi_default, 0 }, { "setBasis", 0, SetBasis, 0, 0, 0, napi_default, 0 },
        { "setBounds", 0, SetBounds, 0, 0, 0, napi_default, 0 }, { "simplex", 0, Simplex, 0, 0, 0, napi_default, 0 },
        { "intopt", 0, Intopt, 0, 0, 0, napi_default, 0 }, { "getSolution", 0, GetSolution, 0, 0, 0, napi_default, 0 },
        { "getMip", 0, GetMip, 0, 0, 0, napi_default, 0 },
        { "scaleProb", 0, ScaleProb, 0, 0, 0, napi_default, 0 }, { "advBasis", 0, AdvBasis, 0, 0, 0, napi_default, 0 },
        { "readLp", 0, ReadLp, 0, 0, 0, napi_default, 0 }, { "writeLp", 0, WriteLp, 0, 0, 0, napi_default, 0 },
        { "nppCreate", 0, NppCreate, 0, 0, 0, napi_default, 0 }, { "nppLoadProb", 0, NppLoadProb, 0, 0, 0, napi_default, 0 },
        { "nppSimplex", 0, NppSimplex, 0, 0, 0, napi_default, 0 }, { "nppInteger", 0, NppInteger, 0, 0, 0, napi_default, 0 },
        { "nppBuildProb", 0, NppBuildProb, 0, 0, 0, napi_default, 0 }, { "nppPostprocess", 0, NppPostprocess, 0, 0, 0, napi_default, 0 },
    };
    napi_define_properties(env, exports, sizeof d / sizeof d[0], d);
    return exports;
}

NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
