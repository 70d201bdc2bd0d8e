/* glpb200_napi.c -- thin N-API addon over libglpb200.so (include/glpb200.h).
 *
 * NOT built or tested in this repository's image (no Node.js, no node_api.h);
 * it is the binding a glpk.js maintainer would add.  It is written against the
 * stable N-API C signatures only and contains no logic: typed arrays in, typed
 * arrays out, one exported function per C-ABI entry point.
 *
 *   build:  cc -shared -fPIC glpb200_napi.c -I../include -I$NODE/include/node \
 *              -L../glpk.js_b200 -lglpb200 -o glpb200.node
 */
#include <node_api.h>
#include <stdlib.h>
#include <string.h>
#include "glpb200.h"

#define NAPI_OK(call) do { if ((call) != napi_ok) { napi_throw_error(env, NULL, #call); return NULL; } } while (0)

static void *typed(napi_env env, napi_value v, size_t *len)
{
    napi_typedarray_type t; void *data = NULL; napi_value ab; size_t off;
    bool is = false;
    if (napi_is_typedarray(env, v, &is) != napi_ok || !is) return NULL;
    if (napi_get_typedarray_info(env, v, &t, len, &data, &ab, &off) != napi_ok) return NULL;
    return data;
}

static void finalize_handle(napi_env env, void *data, void *hint) { (void)env; (void)hint; glpb_destroy((glpb_prob *)data); }

/* create(m, n, dir, c0, Int32 type, F64 lb, F64 ub, F64 coef, Int32 kind, F64 rii, F64 sjj,
          Int32 A_ptr, Int32 A_ind, F64 A_val, device) -> external handle */
static napi_value Create(napi_env env, napi_callback_info info)
{
    size_t argc = 15; napi_value a[15];
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    int32_t m, n, dir, device; double c0; size_t len, nnz;
    NAPI_OK(napi_get_value_int32(env, a[0], &m));
    NAPI_OK(napi_get_value_int32(env, a[1], &n));
    NAPI_OK(napi_get_value_int32(env, a[2], &dir));
    NAPI_OK(napi_get_value_double(env, a[3], &c0));
    NAPI_OK(napi_get_value_int32(env, a[14], &device));
    int *type = typed(env, a[4], &len); double *lb = typed(env, a[5], &len), *ub = typed(env, a[6], &len);
    double *coef = typed(env, a[7], &len); int *kind = typed(env, a[8], &len);
    double *rii = typed(env, a[9], &len), *sjj = typed(env, a[10], &len);
    int *ptr = typed(env, a[11], &len), *ind = typed(env, a[12], &nnz); double *val = typed(env, a[13], &nnz);
    glpb_prob *P = glpb_create(m, n, (int)nnz, dir, c0, type, lb, ub, coef, kind, rii, sjj, ptr, ind, val, device);
    if (!P) { napi_throw_error(env, NULL, glpb_last_error()); return NULL; }
    napi_value ext;
    NAPI_OK(napi_create_external(env, P, finalize_handle, NULL, &ext));
    return ext;
}

static glpb_prob *handle(napi_env env, napi_value v) { void *p = NULL; napi_get_value_external(env, v, &p); return (glpb_prob *)p; }

static napi_value ret_int(napi_env env, int v) { napi_value r; napi_create_int32(env, v, &r); return r; }

/* setBasis(h, Int32 stat[m+n]) ; setBounds(h, Int32 k, Int32 type, F64 lb, F64 ub) */
static napi_value SetBasis(napi_env env, napi_callback_info info)
{
    size_t argc = 2, len; napi_value a[2];
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    return ret_int(env, glpb_set_basis(handle(env, a[0]), typed(env, a[1], &len)));
}

static napi_value SetBounds(napi_env env, napi_callback_info info)
{
    size_t argc = 5, cnt, len; napi_value a[5];
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    int *k = typed(env, a[1], &cnt);
    return ret_int(env, glpb_set_bounds(handle(env, a[0]), (int)cnt, k, typed(env, a[2], &len),
                                        typed(env, a[3], &len), typed(env, a[4], &len)));
}

/* simplex(h, Int32 ip[9], F64 dp[5]) : the SMCP fields packed by the facade */
static napi_value Simplex(napi_env env, napi_callback_info info)
{
    size_t argc = 3, len; napi_value a[3];
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    int *ip = typed(env, a[1], &len); double *dp = typed(env, a[2], &len);
    glpb_smcp s = { ip[0], ip[1], ip[2], ip[3], dp[0], dp[1], dp[2], dp[3], dp[4], ip[4], ip[5], ip[6], ip[7], ip[8] };
    int rc = glpb_simplex(handle(env, a[0]), &s);
    if (rc < 0) { napi_throw_error(env, NULL, glpb_last_error()); return NULL; }
    return ret_int(env, rc);
}

/* intopt(h, Int32 ip[8], F64 dp[3]) */
static napi_value Intopt(napi_env env, napi_callback_info info)
{
    size_t argc = 3, len; napi_value a[3];
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    int *ip = typed(env, a[1], &len); double *dp = typed(env, a[2], &len);
    glpb_iocp s = { ip[0], ip[1], ip[2], dp[0], dp[1], ip[3], ip[4], ip[5], ip[6], dp[2], ip[7], -1 };
    int rc = glpb_intopt(handle(env, a[0]), &s);
    if (rc < 0) { napi_throw_error(env, NULL, glpb_last_error()); return NULL; }
    return ret_int(env, rc);
}

/* getSolution(h, Int32 stat, F64 prim, F64 dual, Int32 head, Int32 ints[4], F64 obj[1]) */
static napi_value GetSolution(napi_env env, napi_callback_info info)
{
    size_t argc = 7, len; napi_value a[7];
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    int *ints = typed(env, a[5], &len); double *obj = typed(env, a[6], &len);
    return ret_int(env, glpb_get_solution(handle(env, a[0]), typed(env, a[1], &len), typed(env, a[2], &len),
                                          typed(env, a[3], &len), typed(env, a[4], &len), &ints[0], &ints[1], obj,
                                          &ints[2], &ints[3]));
}

/* getMip(h, Int32 st[1], F64 obj[1], F64 mipx[m+n]) */
static napi_value GetMip(napi_env env, napi_callback_info info)
{
    size_t argc = 4, len; napi_value a[4];
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    return ret_int(env, glpb_get_mip(handle(env, a[0]), typed(env, a[1], &len), typed(env, a[2], &len),
                                     typed(env, a[3], &len), NULL));
}

/* scaleProb(m, n, Int32 A_ptr, Int32 A_ind, F64 A_val, flags, F64 rii[m], F64 sjj[n], F64 report[13])
   -- glp_scale_prob (lib/glpscl.js) on flat arrays; host only */
static napi_value ScaleProb(napi_env env, napi_callback_info info)
{
    size_t argc = 9, len; napi_value a[9]; int m, n, flags;
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    NAPI_OK(napi_get_value_int32(env, a[0], &m)); NAPI_OK(napi_get_value_int32(env, a[1], &n));
    NAPI_OK(napi_get_value_int32(env, a[5], &flags));
    return ret_int(env, glpb_scale_prob(m, n, typed(env, a[2], &len), typed(env, a[3], &len), typed(env, a[4], &len),
                                        flags, typed(env, a[6], &len), typed(env, a[7], &len), typed(env, a[8], &len)));
}

/* advBasis(m, n, Int32 A_ptr, Int32 A_ind, Int32 R_ptr, Int32 R_ind, Int32 type, F64 lb, F64 ub,
            Int32 stat[m+n], Int32 size[1]) -- glp_adv_basis (lib/glpini01.js); host only */
static napi_value AdvBasis(napi_env env, napi_callback_info info)
{
    size_t argc = 11, len; napi_value a[11]; int m, n;
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    NAPI_OK(napi_get_value_int32(env, a[0], &m)); NAPI_OK(napi_get_value_int32(env, a[1], &n));
    return ret_int(env, glpb_adv_basis(m, n, typed(env, a[2], &len), typed(env, a[3], &len), typed(env, a[4], &len),
                                       typed(env, a[5], &len), typed(env, a[6], &len), typed(env, a[7], &len),
                                       typed(env, a[8], &len), typed(env, a[9], &len), typed(env, a[10], &len)));
}

/* readLp(text) -> { m, n, dir, type, lb, ub, coef, kind, ptr, ind, val, names } -- glp_read_lp
   (lib/glpcpx.js) on a string held in memory; throws the reader's message on a syntax error */
static napi_value ReadLp(napi_env env, napi_callback_info info)
{
    size_t argc = 1, len = 0; napi_value a[1], r, v; glpb_problem_data d; char *names = NULL; long nlen = 0; void *p;
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    NAPI_OK(napi_get_value_string_utf8(env, a[0], NULL, 0, &len));
    char *text = (char *)malloc(len + 1);
    NAPI_OK(napi_get_value_string_utf8(env, a[0], text, len + 1, &len));
    int rc = glpb_read_lp(text, (long)len, &d, &names, &nlen);
    free(text);
    if (rc != 0) { napi_throw_error(env, NULL, glpb_last_error()); return NULL; }
    NAPI_OK(napi_create_object(env, &r));
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

static napi_value Init(napi_env env, napi_value exports)
{
    napi_property_descriptor d[] = {
        { "create", 0, Create, 0, 0, 0, napi_default, 0 }, { "setBasis", 0, SetBasis, 0, 0, 0, napi_default, 0 },
        { "setBounds", 0, SetBounds, 0, 0, 0, napi_default, 0 }, { "simplex", 0, Simplex, 0, 0, 0, napi_default, 0 },
        { "intopt", 0, Intopt, 0, 0, 0, napi_default, 0 }, { "getSolution", 0, GetSolution, 0, 0, 0, napi_default, 0 },
        { "getMip", 0, GetMip, 0, 0, 0, napi_default, 0 },
        { "scaleProb", 0, ScaleProb, 0, 0, 0, napi_default, 0 }, { "advBasis", 0, AdvBasis, 0, 0, 0, napi_default, 0 },
        { "readLp", 0, ReadLp, 0, 0, 0, napi_default, 0 },
    };
    napi_define_properties(env, exports, sizeof d / sizeof d[0], d);
    return exports;
}

NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
