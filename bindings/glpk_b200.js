/* glpk_b200.js -- JS facade: keeps glpk.js's public API and its glp_prob object,
 * and replaces solve_lp / solve_mip (lib/glpapi06.js:3-39, lib/glpapi09.js:62-114)
 * by calls into the N-API addon.  Everything above those two call sites
 * (validation, presolve, scaling, getters) stays the reference's JavaScript.
 * NOT runnable in this repository's image (no Node.js); see INTEGRATION.md.
 *
 *   const glpk = require('glpk');              // the reference
 *   require('./glpk_b200').install(glpk);      // glp_simplex / glp_intopt now run on the GPU
 */
'use strict';
const addon = require('./glpb200.node');

function marshal(glp, P) {
    const m = glp.glp_get_num_rows(P), n = glp.glp_get_num_cols(P);
    const type = new Int32Array(m + n), lb = new Float64Array(m + n), ub = new Float64Array(m + n);
    const coef = new Float64Array(n), kind = new Int32Array(n), rii = new Float64Array(m), sjj = new Float64Array(n);
    const stat = new Int32Array(m + n);
    for (let i = 1; i <= m; i++) {
        type[i - 1] = glp.glp_get_row_type(P, i); lb[i - 1] = P.row[i].lb; ub[i - 1] = P.row[i].ub;
        rii[i - 1] = glp.glp_get_rii(P, i); stat[i - 1] = glp.glp_get_row_stat(P, i);
    }
    const ptr = new Int32Array(n + 1), ind = [], val = [];
    for (let j = 1; j <= n; j++) {
        const k = m + j - 1, col = P.col[j];
        type[k] = col.type; lb[k] = col.lb; ub[k] = col.ub; coef[j - 1] = col.coef;
        kind[j - 1] = col.kind; sjj[j - 1] = col.sjj; stat[k] = col.stat;
        ptr[j - 1] = ind.length;
        for (let a = col.ptr; a != null; a = a.c_next) { ind.push(a.row.i - 1); val.push(a.val); }   // list order
    }
    ptr[n] = ind.length;
    return { m, n, type, lb, ub, coef, kind, rii, sjj, stat, ptr, ind: Int32Array.from(ind), val: Float64Array.from(val) };
}

function device(glp, P) {
    const d = marshal(glp, P);
    if (!P._glpb || P._glpb_nnz !== P.nnz) {          // matrix changed: new handle
        P._glpb = addon.create(d.m, d.n, P.dir, P.c0, d.type, d.lb, d.ub, d.coef, d.kind, d.rii, d.sjj,
                               d.ptr, d.ind, d.val, 0);
        P._glpb_nnz = P.nnz;
    } else {
        const k = Int32Array.from({ length: d.m + d.n }, (_, i) => i + 1);
        addon.setBounds(P._glpb, k, d.type, d.lb, d.ub);
    }
    addon.setBasis(P._glpb, d.stat);
    return d;
}

function pull(glp, P, d) {
    const m = d.m, n = d.n;
    const stat = new Int32Array(m + n), prim = new Float64Array(m + n), dual = new Float64Array(m + n);
    const head = new Int32Array(m), ints = new Int32Array(4), obj = new Float64Array(1);
    addon.getSolution(P._glpb, stat, prim, dual, head, ints, obj);
    P.pbs_stat = ints[0]; P.dbs_stat = ints[1]; P.it_cnt = ints[2]; P.some = ints[3]; P.obj_val = obj[0];
    for (let i = 1; i <= m; i++) { const r = P.row[i]; r.stat = stat[i - 1]; r.prim = prim[i - 1]; r.dual = dual[i - 1]; r.bind = 0; }
    for (let j = 1; j <= n; j++) { const c = P.col[j], k = m + j - 1; c.stat = stat[k]; c.prim = prim[k]; c.dual = dual[k]; c.bind = 0; }
    for (let i = 1; i <= m; i++) { P.head[i] = head[i - 1]; const k = head[i - 1]; if (k <= m) P.row[k].bind = i; else P.col[k - m].bind = i; }
    P.valid = 1;
}

/* rows of A in list order (row.ptr -> r_next), 0-based column indices */
function rowsOf(P, m) {
    const ptr = new Int32Array(m + 1), ind = [];
    for (let i = 1; i <= m; i++) {
        for (let a = P.row[i].ptr; a != null; a = a.r_next) ind.push(a.col.j - 1);
        ptr[i] = ind.length;
    }
    return { ptr, ind: Int32Array.from(ind) };
}

/* preprocess_and_solve_lp (lib/glpapi06.js:41-146) with the native presolver (addon.npp*, csrc/presolve.cpp)
   in place of npp_load_prob / npp_simplex / npp_build_prob / npp_postprocess; npp_unload_sol
   (lib/glpnpp01.js:589-682) stays here because it writes the caller's own glp_prob */
function presolveLp(glp, P, parm) {
    const d = marshal(glp, P), npp = addon.nppCreate();     // marshal: unscaled problem -> rii/sjj unused here
    addon.nppLoadProb(npp, d.m, d.n, P.dir, P.c0, d.type, d.lb, d.ub, d.coef, d.kind, d.ptr, d.ind, d.val, glp.GLP_SOL);
    let ret = addon.nppSimplex(npp);
    if (ret !== 0) return ret;
    const r = addon.nppBuildProb(npp), lp = glp.glp_create_prob();
    glp.glp_set_obj_dir(lp, P.dir); glp.glp_set_obj_coef(lp, 0, r.c0);
    if (r.m) glp.glp_add_rows(lp, r.m);
    for (let i = 1; i <= r.m; i++) glp.glp_set_row_bnds(lp, i, r.type[i - 1], r.lb[i - 1], r.ub[i - 1]);
    if (r.n) glp.glp_add_cols(lp, r.n);
    for (let j = 1; j <= r.n; j++) {
        const k = r.m + j - 1, a = r.ptr[j - 1], len = r.ptr[j] - a;
        glp.glp_set_col_bnds(lp, j, r.type[k], r.lb[k], r.ub[k]); glp.glp_set_obj_coef(lp, j, r.coef[j - 1]);
        const ind = new Int32Array(1 + len), val = new Float64Array(1 + len);
        for (let t = 0; t < len; t++) { ind[1 + t] = r.ind[a + t] + 1; val[1 + t] = r.val[a + t]; }
        glp.glp_set_mat_col(lp, j, len, ind, val);            // prepends: the reference's list state
    }
    if (r.m === 0 && r.n === 0) { lp.pbs_stat = lp.dbs_stat = glp.GLP_FEAS; lp.obj_val = lp.c0; }
    else {
        glp.glp_scale_prob(lp, glp.GLP_SF_AUTO); glp.glp_adv_basis(lp, 0);
        lp.it_cnt = P.it_cnt;
        const inner = Object.assign(new glp.SMCP(), parm, { presolve: glp.GLP_OFF });
        ret = glp.glp_simplex(lp, inner);                     // the device solve of the REDUCED problem
        P.it_cnt = lp.it_cnt;
        if (!(ret === 0 && lp.pbs_stat === glp.GLP_FEAS && lp.dbs_stat === glp.GLP_FEAS)) {
            if (ret === 0) ret = lp.pbs_stat === glp.GLP_NOFEAS ? glp.GLP_ENOPFS : glp.GLP_ENODFS;
            return ret;
        }
    }
    const s = addon.nppPostprocess(npp,
        Int32Array.from({ length: r.m }, (_, i) => lp.row[i + 1].stat), Float64Array.from({ length: r.m }, (_, i) => lp.row[i + 1].dual),
        Int32Array.from({ length: r.n }, (_, j) => lp.col[j + 1].stat), Float64Array.from({ length: r.n }, (_, j) => lp.col[j + 1].prim));
    /* npp_unload_sol, basic solution */
    const bound = (x) => (x.stat === glp.GLP_NU ? x.ub : x.stat === glp.GLP_NF ? 0.0 : x.lb);
    P.valid = 0; P.pbs_stat = lp.pbs_stat; P.dbs_stat = lp.dbs_stat; P.obj_val = P.c0; P.some = 0;
    for (let i = 1; i <= d.m; i++) {
        const row = P.row[i]; row.stat = s.rowStat[i - 1]; row.dual = s.rowDual[i - 1];
        if (row.stat === glp.GLP_BS) row.dual = 0.0; else row.prim = bound(row);
    }
    for (let j = 1; j <= d.n; j++) {
        const col = P.col[j]; col.stat = s.colStat[j - 1]; col.prim = s.colValue[j - 1];
        if (col.stat === glp.GLP_BS) col.dual = 0.0; else col.prim = bound(col);
        P.obj_val += col.coef * col.prim;
    }
    for (let i = 1; i <= d.m; i++) {
        const row = P.row[i];
        if (row.stat === glp.GLP_BS) { let t = 0.0; for (let a = row.ptr; a != null; a = a.r_next) t += a.val * a.col.prim; row.prim = t; }
    }
    for (let j = 1; j <= d.n; j++) {
        const col = P.col[j];
        if (col.stat !== glp.GLP_BS) { let t = col.coef; for (let a = col.ptr; a != null; a = a.c_next) t -= a.val * a.row.dual; col.dual = t; }
    }
    return 0;
}

exports.install = function (glp, opts) {
    opts = opts || {};
    const simplex0 = glp.glp_simplex, intopt0 = glp.glp_intopt;
    /* glp_scale_prob / glp_adv_basis (lib/glpscl.js, lib/glpini01.js): O(nnz) host work in the
       native library instead of linked-list walks; same factors, same statuses */
    glp.glp_scale_prob = function (P, flags) {
        const d = marshal(glp, P), rep = new Float64Array(13);
        if (addon.scaleProb(d.m, d.n, d.ptr, d.ind, d.val, flags | 0, d.rii, d.sjj, rep) !== 0)
            throw new Error('glp_scale_prob: flags = ' + flags + '; invalid scaling options');
        for (let i = 1; i <= d.m; i++) glp.glp_set_rii(P, i, d.rii[i - 1]);
        for (let j = 1; j <= d.n; j++) glp.glp_set_sjj(P, j, d.sjj[j - 1]);
    };
    glp.glp_adv_basis = function (P, flags) {
        if (flags) throw new Error('glp_adv_basis: flags = ' + flags + '; invalid flags');
        const d = marshal(glp, P);
        if (d.m === 0 || d.n === 0) return glp.glp_std_basis(P);
        const r = rowsOf(P, d.m), stat = new Int32Array(d.m + d.n), size = new Int32Array(1);
        addon.advBasis(d.m, d.n, d.ptr, d.ind, r.ptr, r.ind, d.type, d.lb, d.ub, stat, size);
        for (let i = 1; i <= d.m; i++) glp.glp_set_row_stat(P, i, stat[i - 1]);
        for (let j = 1; j <= d.n; j++) glp.glp_set_col_stat(P, j, stat[d.m + j - 1]);
    };
    glp.glp_simplex = function (P, parm) {
        parm = parm || new glp.SMCP();
        if (parm.presolve)      // native presolver on request, else the JS presolver (it calls back into solve_lp)
            return opts.nativePresolve ? presolveLp(glp, P, parm) : simplex0(P, parm);
        const d = device(glp, P);
        const ret = addon.simplex(P._glpb,
            Int32Array.from([parm.msg_lev, parm.meth, parm.pricing, parm.r_test, parm.it_lim, parm.tm_lim, parm.out_frq, parm.out_dly, 0]),
            Float64Array.from([parm.tol_bnd, parm.tol_dj, parm.tol_piv, parm.obj_ll, parm.obj_ul]));
        pull(glp, P, d);
        return ret;
    };
    glp.glp_intopt = function (P, parm) {
        parm = parm || new glp.IOCP();
        if (parm.presolve || parm.cb_func) return intopt0(P, parm);   // callbacks: host-driven reference loop
        const d = device(glp, P);
        const ret = addon.intopt(P._glpb,
            Int32Array.from([parm.msg_lev, parm.br_tech, parm.bt_tech, parm.tm_lim, parm.out_frq, parm.out_dly, parm.pp_tech, 0]),
            Float64Array.from([parm.tol_int, parm.tol_obj, parm.mip_gap]));
        const st = new Int32Array(1), obj = new Float64Array(1), x = new Float64Array(d.m + d.n);
        addon.getMip(P._glpb, st, obj, x);
        P.mip_stat = st[0]; P.mip_obj = obj[0];
        for (let i = 1; i <= d.m; i++) P.row[i].mipx = x[i - 1];
        for (let j = 1; j <= d.n; j++) P.col[j].mipx = x[d.m + j - 1];
        return ret;
    };
};
