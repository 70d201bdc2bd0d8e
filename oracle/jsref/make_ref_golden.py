#!/usr/bin/env python
"""Golden vectors FROM THE REFERENCE ITSELF.

Runs the unmodified sources of /root/reference/lib/*.js in the minijs
interpreter (oracle/jsref/minijs.py) and records, for the reference's own
fixtures and for small generated LPs / MIPs:

  * what test/test.js prints: return codes, status, objective, column values of
    glp_simplex and glp_intopt, presolve OFF and ON;
  * the pivot sequence of spx_primal / spx_dual: after every chuzc / chuzr call
    the selected (q, p, p_stat, teta) resp. (p, delta, q, new_dq)
    (lib/glpspx01.js:646-688, 808-1028; lib/glpspx02.js:572-625, 793-935);
  * live arrays around selected calls -- the inputs the reference handed to
    chuzc / chuzr / the ratio tests / eval_trow / update_gamma and what came
    back -- as inputs of the kernel-level parity tests.

Output (committed): tests/golden/ref_runs.json, tests/golden/ref_vectors.npz.
Run in the build container only:   python oracle/jsref/make_ref_golden.py
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refjs  # noqa: E402
from minijs import NativeFunc, UNDEF  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
GLP_PRIMAL, GLP_DUALP, GLP_DUAL = 1, 2, 3


class Ref:
    def __init__(self):
        self.I = refjs.load()
        self.api, self.g = self.I.api, self.I.globals

    def call(self, name, *a):
        return self.api[name].call(self.g, list(a))

    def problem(self, text):
        lp = self.call("glp_create_prob")
        rc = refjs.read_lp_text(self.I, lp, text)
        assert rc == 0, "glp_read_lp failed"
        return lp

    def problem_from_arrays(self, d):
        """d: oracle layout (r_type/r_lb/r_ub, c_type/c_lb/c_ub/c_coef/c_kind, 0-based CSC with
        ascending rows).  Built through the reference's own API; glp_sort_matrix leaves the element
        lists in the state glp_read_lp leaves them (lib/glpcpx.js:745), which is what the oracle's
        loader reproduces."""
        from minijs import Int32Array, Float64Array
        lp = self.call("glp_create_prob")
        m, n = int(d["m"]), int(d["n"])
        self.call("glp_add_rows", lp, m)
        self.call("glp_add_cols", lp, n)
        self.call("glp_set_obj_dir", lp, int(d["dir"]))
        self.call("glp_set_obj_coef", lp, 0, float(d["c0"]))
        for i in range(m):
            self.call("glp_set_row_bnds", lp, i + 1, int(d["r_type"][i]), float(d["r_lb"][i]), float(d["r_ub"][i]))
        for j in range(n):
            self.call("glp_set_col_bnds", lp, j + 1, int(d["c_type"][j]), float(d["c_lb"][j]), float(d["c_ub"][j]))
            self.call("glp_set_obj_coef", lp, j + 1, float(d["c_coef"][j]))
            if int(d["c_kind"][j]) != 1:
                self.call("glp_set_col_kind", lp, j + 1, int(d["c_kind"][j]))
            a, b = int(d["A_ptr"][j]), int(d["A_ptr"][j + 1])
            ind = Int32Array([0] + [int(x) + 1 for x in d["A_ind"][a:b]])
            val = Float64Array([0.0] + [float(x) for x in d["A_val"][a:b]])
            self.call("glp_set_mat_col", lp, j + 1, b - a, ind, val)
        self.call("glp_sort_matrix", lp)
        return lp

    def make(self, src):
        return self.problem(src) if isinstance(src, str) else self.problem_from_arrays(src)

    def smcp(self, **kw):
        p = {}
        self.api["SMCP"].call(p, [{}])
        p["msg_lev"] = 0
        p.update(kw)
        return p

    def iocp(self, **kw):
        p = {}
        self.api["IOCP"].call(p, [{}])
        p["msg_lev"] = 0
        p.update(kw)
        return p

    def lp_result(self, lp):
        m, n = self.call("glp_get_num_rows", lp), self.call("glp_get_num_cols", lp)
        return dict(status=self.call("glp_get_status", lp), prim_stat=self.call("glp_get_prim_stat", lp),
                    dual_stat=self.call("glp_get_dual_stat", lp), obj=float(self.call("glp_get_obj_val", lp)),
                    it_cnt=int(lp["it_cnt"]),
                    col_prim=[float(self.call("glp_get_col_prim", lp, j)) for j in range(1, n + 1)],
                    col_dual=[float(self.call("glp_get_col_dual", lp, j)) for j in range(1, n + 1)],
                    row_prim=[float(self.call("glp_get_row_prim", lp, i)) for i in range(1, m + 1)],
                    row_dual=[float(self.call("glp_get_row_dual", lp, i)) for i in range(1, m + 1)],
                    col_stat=[int(self.call("glp_get_col_stat", lp, j)) for j in range(1, n + 1)],
                    row_stat=[int(self.call("glp_get_row_stat", lp, i)) for i in range(1, m + 1)])

    def mip_result(self, lp):
        n = self.call("glp_get_num_cols", lp)
        return dict(mip_stat=self.call("glp_mip_status", lp), mip_obj=float(self.call("glp_mip_obj_val", lp)),
                    col_val=[float(self.call("glp_mip_col_val", lp, j)) for j in range(1, n + 1)])


def arr(x, dtype=np.float64):
    return np.array([0 if v is UNDEF else v for v in x], dtype=dtype)


class Tracer:
    """hooks on the nested functions of spx_primal / spx_dual"""

    def __init__(self, ref, capture_at=(), capture_gamma_at=()):
        self.ref, self.seq, self.vec = ref, [], {}
        self.capture_at, self.capture_gamma_at = set(capture_at), set(capture_gamma_at)
        self.cur = None
        self.n_iter = 0
        h = ref.I.hooks
        for q in ("spx_primal.chuzc", "spx_primal.chuzr", "spx_dual.chuzr", "spx_dual.chuzc", "spx_primal.update_gamma",
                  "spx_dual.update_gamma", "spx_dual.eval_trow", "spx_primal.eval_trow", "spx_dual.sort_trow",
                  "spx_primal.sort_tcol"):
            h[q] = self.on

    def close(self):
        self.ref.I.hooks.clear()

    def snap(self, csa, names):
        out = {}
        for nme in names:
            v = csa.get(nme, UNDEF)
            if isinstance(v, list):
                kind = getattr(v, "kind", "f64")
                out[nme] = arr(v, {"f64": np.float64, "i32": np.int32, "i8": np.int8}[kind])
        return out

    def on(self, when, qname, args, ret):
        csa = args[0]
        it = self.n_iter
        if qname == "spx_primal.chuzc":
            if when == "enter":
                if it in self.capture_at:
                    self.vec["p_chuzc_%d" % it] = dict(self.snap(csa, ("stat", "cbar", "gamma")), tol=float(args[1]), n=csa["n"])
            else:
                self.cur = dict(kind="P", q=int(csa["q"]))
                if it in self.capture_at:
                    self.vec["p_chuzc_%d" % it]["q"] = int(csa["q"])
                if csa["q"] == 0:
                    self.seq.append(self.cur)
        elif qname == "spx_primal.chuzr":
            if when == "enter":
                if it in self.capture_at:
                    self.vec["p_chuzr_%d" % it] = dict(
                        self.snap(csa, ("type", "lb", "ub", "coef", "head", "bbar", "cbar", "tcol_ind", "tcol_vec")),
                        rtol=float(args[1]), m=csa["m"], n=csa["n"], q=int(csa["q"]), phase=int(csa["phase"]),
                        tcol_num=int(csa["tcol_num"]), tcol_nnz=int(csa["tcol_nnz"]))
            else:
                self.cur.update(p=int(csa["p"]), p_stat=int(csa["p_stat"]), teta=float(csa["teta"]), phase=int(csa["phase"]))
                self.seq.append(self.cur)
                if it in self.capture_at:
                    self.vec["p_chuzr_%d" % it].update(p=int(csa["p"]), p_stat=int(csa["p_stat"]), teta=float(csa["teta"]))
                if csa["p"] != 0:
                    self.n_iter += 1
        elif qname == "spx_dual.chuzr":
            if when == "enter":
                if it in self.capture_at:
                    self.vec["d_chuzr_%d" % it] = dict(self.snap(csa, ("type", "lb", "ub", "head", "bbar", "gamma")),
                                                       tol=float(args[1]), m=csa["m"], n=csa["n"])
            else:
                self.cur = dict(kind="D", p=int(csa["p"]), delta=float(csa["delta"]), phase=int(csa["phase"]))
                if it in self.capture_at:
                    self.vec["d_chuzr_%d" % it].update(p=int(csa["p"]), delta=float(csa["delta"]))
                if csa["p"] == 0:
                    self.seq.append(self.cur)
        elif qname == "spx_dual.chuzc":
            if when == "enter":
                if it in self.capture_at:
                    self.vec["d_chuzc_%d" % it] = dict(self.snap(csa, ("stat", "cbar", "trow_ind", "trow_vec")),
                                                       rtol=float(args[1]), n=csa["n"], delta=float(csa["delta"]),
                                                       trow_num=int(csa["trow_num"]), trow_nnz=int(csa["trow_nnz"]))
            else:
                self.cur.update(q=int(csa["q"]), new_dq=float(csa["new_dq"]))
                self.seq.append(self.cur)
                if it in self.capture_at:
                    self.vec["d_chuzc_%d" % it].update(q=int(csa["q"]), new_dq=float(csa["new_dq"]))
                if csa["q"] != 0:
                    self.n_iter += 1
        elif qname == "spx_dual.eval_trow":
            if when == "exit" and it in self.capture_at:
                rho = args[1]
                self.vec["d_trow_%d" % it] = dict(self.snap(csa, ("A_ptr", "A_ind", "A_val", "head", "stat", "trow_vec")),
                                                  rho=arr(rho), m=csa["m"], n=csa["n"])
        elif qname in ("spx_primal.update_gamma", "spx_dual.update_gamma"):
            # n_iter was already advanced by the ratio-test hook of this iteration
            itg = it - 1
            tag = ("p" if qname.startswith("spx_primal") else "d") + "_gamma_%d" % itg
            if itg in self.capture_gamma_at:
                if when == "enter":
                    names = ("type", "head", "stat", "refsp", "gamma", "tcol_ind", "tcol_vec", "trow_ind", "trow_vec",
                             "A_ptr", "A_ind", "A_val")
                    self.vec[tag] = dict(self.snap(csa, names), m=csa["m"], n=csa["n"], p=int(csa["p"]), q=int(csa["q"]),
                                         tcol_nnz=int(csa["tcol_nnz"]), trow_nnz=int(csa["trow_nnz"]))
                else:
                    self.vec[tag]["gamma_out"] = arr(csa["gamma"])
                    self.vec[tag]["refsp_out"] = arr(csa["refsp"], np.int8)
        elif qname == "spx_primal.sort_tcol":
            if when == "exit" and it in self.capture_at:
                self.vec["p_sort_%d" % it] = dict(tcol_ind=arr(csa["tcol_ind"], np.int32), tcol_num=int(csa["tcol_num"]),
                                                  tcol_nnz=int(csa["tcol_nnz"]))


def traced_solve(ref, text, meth, presolve=0, capture_at=(), capture_gamma_at=(), **kw):
    lp = ref.make(text)
    tr = Tracer(ref, capture_at, capture_gamma_at)
    t0 = time.time()
    ret = ref.call("glp_simplex", lp, ref.smcp(meth=meth, presolve=presolve, **kw))
    tr.close()
    res = ref.lp_result(lp)
    res.update(ret=int(ret), seconds=round(time.time() - t0, 2), pivots=tr.seq)
    return lp, res, tr.vec


def flatten(vec, prefix, out):
    for tag, d in vec.items():
        for k, v in d.items():
            out["%s/%s/%s" % (prefix, tag, k)] = np.asarray(v)


def main():
    import oracle_lib as O
    import helpers as H
    ref = Ref()
    runs = {"_about": "generated by oracle/jsref/make_ref_golden.py from the unmodified reference sources "
                      "(lib/*.js executed by oracle/jsref/minijs.py); files loaded: " + " ".join(ref.I.loaded_files)}
    vectors = {}
    # ---- the reference's fixtures: what test/test.js does, both presolve settings ----
    prev = {}
    if os.environ.get("REF_REUSE_FIXTURES") and os.path.exists(os.path.join(GOLD, "ref_runs.json")):
        prev = json.load(open(os.path.join(GOLD, "ref_runs.json")))
        old = np.load(os.path.join(GOLD, "ref_vectors.npz"))
        for k in old.files:
            if k.split("/")[0].split("_")[0] in ("test", "gap", "todd"):
                vectors[k] = old[k]
    for name in ("test", "gap", "todd"):
        if name in prev:
            runs[name] = prev[name]
            continue
        text = open(os.path.join(refjs.REF, "test", name + ".lpt")).read()
        entry = {}
        for presolve in (0, 1):
            lp = ref.problem(text)
            t0 = time.time()
            r1 = ref.call("glp_simplex", lp, ref.smcp(presolve=presolve))
            lpres = ref.lp_result(lp)
            lpres["ret"] = int(r1)
            n_nodes = [0]
            ref.I.hooks["ios_solve_node"] = lambda when, q, a, r: n_nodes.__setitem__(0, n_nodes[0] + (when == "enter"))
            r2 = ref.call("glp_intopt", lp, ref.iocp(presolve=presolve))
            ref.I.hooks.clear()
            mip = ref.mip_result(lp)
            mip.update(ret=int(r2), nodes_solved=n_nodes[0])
            entry["presolve_%d" % presolve] = dict(lp=lpres, mip=mip, seconds=round(time.time() - t0, 1))
            print(name, "presolve", presolve, lpres["obj"], lpres["it_cnt"], mip["mip_obj"], mip["nodes_solved"], flush=True)
        # pivot sequences of the three methods, presolve off
        for meth, mname in ((GLP_PRIMAL, "primal"), (GLP_DUAL, "dual"), (GLP_DUALP, "dualp")):
            cap = (0, 1, 2, 5, 10, 20) if name != "todd" else (0, 1, 2)
            _, res, vec = traced_solve(ref, text, meth, capture_at=cap, capture_gamma_at=cap)
            entry["trace_" + mname] = res
            flatten(vec, name + "_" + mname, vectors)
            print(name, mname, res["ret"], res["obj"], res["it_cnt"], len(res["pivots"]), flush=True)
        runs[name] = entry
    # ---- generated LPs: the random general LPs of tests/helpers.py through the oracle's LP writer ----
    gen = {}
    for seed in (1, 2, 3, 5, 8, 13, 21, 34):
        d = H.random_lp(seed)
        e = {"gen": "helpers.random_lp(%d)" % seed, "checksum": float(np.sum(d["A_val"]) + np.sum(d["c_coef"]))}
        for meth, mname in ((GLP_PRIMAL, "primal"), (GLP_DUAL, "dual")):
            _, res, vec = traced_solve(ref, d, meth, capture_at=(0, 3, 7), capture_gamma_at=(0, 3, 7))
            e["trace_" + mname] = res
            flatten(vec, "rand%d_%s" % (seed, mname), vectors)
            print("random_lp", seed, mname, res["ret"], res["status"], res["obj"], res["it_cnt"], flush=True)
        gen["random_lp_%d" % seed] = e
    for kind, kw in (("packing", dict(m=24, n=48, density=0.3, seed=5)), ("covering", dict(m=40, n=80, kmin=3, kspan=4, seed=7))):
        dn = O.generate(kind, **kw)
        meth, mname = (GLP_PRIMAL, "primal") if kind == "packing" else (GLP_DUAL, "dual")
        _, res, vec = traced_solve(ref, H.to_oracle(dn), meth, capture_at=(0, 5, 15, 30), capture_gamma_at=(0, 5, 15, 30))
        gen[kind] = {"gen": dict(kw, kind=kind), "trace_" + mname: res}
        flatten(vec, "%s_%s" % (kind, mname), vectors)
        print(kind, res["ret"], res["status"], res["obj"], res["it_cnt"], flush=True)
    # ---- small MIPs: knapsack of the C5 family and random integer programs ----
    for tag, dn in (("mkp_5x30", O.generate("mkp", m=5, n=30, seed=20240701)),
                    ("mkp_4x16", O.generate("mkp", m=4, n=16, seed=11))):
        lp = ref.problem_from_arrays(H.to_oracle(dn))
        r1 = ref.call("glp_simplex", lp, ref.smcp())
        lpres = ref.lp_result(lp)
        n_nodes = [0]
        ref.I.hooks["ios_solve_node"] = lambda when, q, a, r: n_nodes.__setitem__(0, n_nodes[0] + (when == "enter"))
        r2 = ref.call("glp_intopt", lp, ref.iocp())
        ref.I.hooks.clear()
        mip = ref.mip_result(lp)
        mip.update(ret=int(r2), nodes_solved=n_nodes[0])
        gen[tag] = {"lp": dict(lpres, ret=int(r1)), "mip": mip}
        print(tag, lpres["obj"], mip["mip_obj"], mip["nodes_solved"], flush=True)
    for seed in (3, 12, 20):
        d = H.random_mip(seed)
        lp = ref.problem_from_arrays(d)
        r1 = ref.call("glp_simplex", lp, ref.smcp())
        lpres = ref.lp_result(lp)
        mip = None
        if r1 == 0 and lpres["status"] == 5:
            n_nodes = [0]
            ref.I.hooks["ios_solve_node"] = lambda when, q, a, r: n_nodes.__setitem__(0, n_nodes[0] + (when == "enter"))
            r2 = ref.call("glp_intopt", lp, ref.iocp())
            ref.I.hooks.clear()
            mip = ref.mip_result(lp)
            mip.update(ret=int(r2), nodes_solved=n_nodes[0])
        gen["random_mip_%d" % seed] = {"gen": "helpers.random_mip(%d)" % seed, "lp": dict(lpres, ret=int(r1)), "mip": mip}
        print("random_mip", seed, lpres["obj"], mip and mip["mip_obj"], mip and mip["nodes_solved"], flush=True)
    runs["generated"] = gen
    with open(os.path.join(GOLD, "ref_runs.json"), "w") as f:
        json.dump(runs, f)
    np.savez_compressed(os.path.join(GOLD, "ref_vectors.npz"), **vectors)
    print("wrote ref_runs.json (%d bytes), ref_vectors.npz (%d arrays)" % (
        os.path.getsize(os.path.join(GOLD, "ref_runs.json")), len(vectors)))


if __name__ == "__main__":
    main()
