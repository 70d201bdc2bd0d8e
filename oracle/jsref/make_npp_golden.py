#!/usr/bin/env python
"""Golden vectors of the LP / MIP presolver FROM THE REFERENCE ITSELF.

Runs the unmodified lib/glpnpp0[1-5].js (through oracle/jsref/minijs.py) on the
reference's own fixtures and on small generated problems built to trigger the
presolver's transformations (free / empty / singleton rows, fixed / empty /
singleton columns, nearly equal bounds, redundant bounds, forcing rows, implied
slack and implied free variables; for MIPs also bound tightening, hidden packing
/ covering inequalities, coefficient reduction and binarization), in the order
lib/glpapi06.js:41-146 and lib/glpapi09.js:116-256 drive it:

    npp_load_prob, npp_simplex | npp_integer, npp_build_prob,
    scale + crash basis + solve of the REDUCED problem (the reference's own simplex / B&B),
    npp_postprocess, npp_unload_sol

and records per case: the problem, the return code, the number of recovery-stack
entries, the reduced problem exactly as npp_build_prob left it (row / column
order, bounds, costs, element order inside every column, reference numbers), the
solution of the reduced problem that went INTO npp_postprocess, what came out of
it (npp.r_stat, r_pi, c_stat, c_value) and the solution npp_unload_sol stored.

Output (committed): tests/golden/ref_npp.json.  Build container only:
    python oracle/jsref/make_npp_golden.py
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
from minijs import UNDEF  # noqa: E402
from make_ref_golden import Ref  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
DBL_MAX = sys.float_info.max
FR, LO, UP, DB, FX = 1, 2, 3, 4, 5
GLP_SOL, GLP_MIP = 1, 3


def null(x):
    return x is None or x is UNDEF


# ---------------------------------------------------------------- problem generators (recorded in the golden)

def npp_lp(seed, m=10, n=14, wild=False):
    """feasible by construction (bounds drawn around a point x0), rich in presolvable structure"""
    rs = np.random.RandomState(seed)
    A = np.zeros((m, n))
    for i in range(m):
        k = rs.choice([0, 1, 1, 2, 3, 4, 5]) if i < m - 1 else 3
        for j in rs.choice(n, size=min(k, n), replace=False):
            A[i, j] = float(rs.choice([-3, -2, -1, 1, 2, 3, 4])) * rs.choice([1.0, 1.0, 0.5, 1.25])
    for j in rs.choice(n, size=2, replace=False):      # column singletons / empty columns
        nz = np.nonzero(A[:, j])[0]
        for i in nz[rs.randint(0, 2):]:
            A[i, j] = 0.0
    c_type, c_lb, c_ub = np.zeros(n, int), np.zeros(n), np.zeros(n)
    x0 = np.zeros(n)
    for j in range(n):
        t = rs.choice([FR, LO, LO, UP, DB, DB, DB, FX])
        lo = float(rs.randint(-4, 3))
        hi = lo + float(rs.randint(1, 6))
        if t == DB and rs.rand() < 0.15:
            hi = lo + 1e-10                             # nearly fixed (npp_make_fixed)
        c_type[j] = t
        if t == FR:
            x0[j] = float(rs.randint(-2, 3))
        elif t == LO:
            c_lb[j] = lo; x0[j] = lo + rs.choice([0.0, 1.0, 2.5])
        elif t == UP:
            c_ub[j] = hi; x0[j] = hi - rs.choice([0.0, 1.0, 2.5])
        elif t == DB:
            c_lb[j], c_ub[j] = lo, hi; x0[j] = lo + (hi - lo) * rs.choice([0.0, 0.5, 1.0])
        else:
            c_lb[j] = c_ub[j] = lo; x0[j] = lo
    r0 = A @ x0
    r_type, r_lb, r_ub = np.zeros(m, int), np.zeros(m), np.zeros(m)
    for i in range(m):
        t = rs.choice([FR, LO, LO, UP, UP, DB, DB, FX])
        slack_lo, slack_hi = rs.choice([0.0, 0.5, 2.0, 30.0]), rs.choice([0.0, 0.5, 2.0, 30.0])
        r_type[i] = t
        if t == LO:
            r_lb[i] = r0[i] - slack_lo
        elif t == UP:
            r_ub[i] = r0[i] + slack_hi
        elif t == DB:
            r_lb[i], r_ub[i] = r0[i] - slack_lo - 0.25, r0[i] + slack_hi + 0.25
            if rs.rand() < 0.2:
                r_ub[i] = r_lb[i] + 1e-10               # nearly an equality (npp_make_equality)
        elif t == FX:
            r_lb[i] = r_ub[i] = r0[i]
    # forcing rows: the bound equals the extreme activity the column bounds allow
    for i in rs.choice(m, size=2, replace=False):
        nz = np.nonzero(A[i])[0]
        if len(nz) < 2 or any(c_type[j] != DB or c_ub[j] - c_lb[j] < 0.5 for j in nz):
            continue
        lo = sum(A[i, j] * (c_lb[j] if A[i, j] > 0 else c_ub[j]) for j in nz)
        r_type[i], r_lb[i], r_ub[i] = UP, 0.0, lo
        x0[nz] = [c_lb[j] if A[i, j] > 0 else c_ub[j] for j in nz]
        break
    coef = np.array([float(rs.choice([-3, -2, -1, 0, 0, 1, 2, 3, 5])) for _ in range(n)])
    # keep the objective bounded: no improving direction along an infinite bound
    for j in range(n):
        if wild:
            break       # any sign: dual infeasible / unbounded problems (GLP_ENODFS, non-optimal reduced LPs)
        if c_type[j] == FR:
            coef[j] = 0.0
        elif c_type[j] == LO:
            coef[j] = abs(coef[j])
        elif c_type[j] == UP:
            coef[j] = -abs(coef[j])
    return pack(m, n, 1 + seed % 2, float(rs.randint(-3, 4)), A, r_type, r_lb, r_ub, c_type, c_lb, c_ub,
                coef * (1 if seed % 2 == 0 else -1), np.ones(n, int))


def npp_mip(seed, m=9, n=12, wide=False):
    """binaries in knapsack-like rows, bounded general integers (binarization), a few continuous columns"""
    rs = np.random.RandomState(1000 + seed)
    kind = np.array([2] * (n - 3) + [1] * 3)
    c_type, c_lb, c_ub = np.full(n, DB), np.zeros(n), np.ones(n)
    for j in range(n - 3):
        if rs.rand() < 0.3:
            c_lb[j] = float(rs.choice([0, 0, 1, -2])); c_ub[j] = c_lb[j] + float(rs.choice([2, 3, 5, 7, 9]))
            if wide and rs.rand() < 0.5:    # ranges the binarization refuses (> 4095) or bounds beyond 1e6
                c_ub[j] = c_lb[j] + float(rs.choice([4095, 4096, 5000, 2000000]))
    for j in range(n - 3, n):
        c_lb[j] = 0.0; c_ub[j] = float(rs.choice([4, 10, 25]))
    A = np.zeros((m, n))
    r_type, r_lb, r_ub = np.zeros(m, int), np.zeros(m), np.zeros(m)
    binary = [j for j in range(n - 3) if c_lb[j] == 0 and c_ub[j] == 1]
    for i in range(m):
        shape = rs.choice(["pack", "cover", "reduce", "mixed", "mixed", "eqsing", "range"])
        if shape in ("pack", "cover", "reduce", "range") and len(binary) >= 3:
            js = rs.choice(binary, size=min(len(binary), rs.randint(3, 6)), replace=False)
            if shape == "pack":
                a = rs.choice([3.0, 4.0, 5.0, 6.0], size=len(js)) * rs.choice([1, 1, 1, -1], size=len(js))
                A[i, js] = a
                r_type[i], r_ub[i] = UP, 6.0 + float(np.sum(a[a < 0]))
            elif shape == "cover":
                a = rs.choice([2.0, 3.0, 5.0], size=len(js)) * rs.choice([1, 1, 1, -1], size=len(js))
                A[i, js] = a
                r_type[i], r_lb[i] = LO, 2.0 + float(np.sum(a[a < 0]))
            elif shape == "reduce":
                a = rs.choice([1.0, 2.0, 9.0, 12.0], size=len(js))
                A[i, js] = a
                r_type[i], r_lb[i] = LO, 5.0
            else:
                a = rs.choice([3.0, 4.0, 5.0], size=len(js))
                A[i, js] = a
                r_type[i], r_lb[i], r_ub[i] = DB, 3.0, 6.0
        elif shape == "eqsing":
            j = rs.randint(0, n)
            A[i, j] = 2.0
            r_type[i] = rs.choice([FX, UP, LO])
            v = 2.0 * float(np.clip(rs.randint(0, 3), c_lb[j], c_ub[j]))
            r_lb[i] = r_ub[i] = v
        else:
            js = rs.choice(n, size=rs.randint(2, 6), replace=False)
            A[i, js] = rs.choice([-2.0, -1.0, 1.0, 2.0, 3.0, 1.5], size=len(js))
            hi = float(np.sum(np.where(A[i] > 0, A[i] * c_ub, A[i] * c_lb)))
            lo = float(np.sum(np.where(A[i] > 0, A[i] * c_lb, A[i] * c_ub)))
            r_type[i], r_ub[i] = UP, lo + 0.6 * (hi - lo)
    coef = rs.choice([-5.0, -4.0, -3.0, -2.0, -1.0, 1.0, 2.0], size=n)
    return pack(m, n, 1, 0.0, A, r_type, r_lb, r_ub, c_type, c_lb, c_ub, coef, kind)


def pack(m, n, dir_, c0, A, r_type, r_lb, r_ub, c_type, c_lb, c_ub, coef, kind):
    ptr, ind, val = [0], [], []
    for j in range(n):
        for i in range(m):
            if A[i, j] != 0.0:
                ind.append(i); val.append(float(A[i, j]))
        ptr.append(len(ind))
    return dict(m=m, n=n, dir=dir_, c0=c0, r_type=[int(x) for x in r_type], r_lb=[float(x) for x in r_lb],
                r_ub=[float(x) for x in r_ub], c_type=[int(x) for x in c_type], c_lb=[float(x) for x in c_lb],
                c_ub=[float(x) for x in c_ub], c_coef=[float(x) for x in coef], c_kind=[int(x) for x in kind],
                A_ptr=ptr, A_ind=ind, A_val=val)


def problem_arrays(ref, lp):
    """what a binding hands to glpb_npp_load_prob: the reference's problem object, columns in LIST order"""
    m, n = int(lp["m"]), int(lp["n"])
    d = dict(m=m, n=n, dir=int(lp["dir"]), c0=float(lp["c0"]), r_type=[], r_lb=[], r_ub=[], c_type=[], c_lb=[],
             c_ub=[], c_coef=[], c_kind=[], A_ptr=[0], A_ind=[], A_val=[])
    for i in range(1, m + 1):
        r = lp["row"][i]
        d["r_type"].append(int(r["type"])); d["r_lb"].append(float(r["lb"])); d["r_ub"].append(float(r["ub"]))
    for j in range(1, n + 1):
        c = lp["col"][j]
        d["c_type"].append(int(c["type"])); d["c_lb"].append(float(c["lb"])); d["c_ub"].append(float(c["ub"]))
        d["c_coef"].append(float(c["coef"])); d["c_kind"].append(int(c["kind"]))
        a = c["ptr"]
        while not null(a):
            d["A_ind"].append(int(a["row"]["i"]) - 1); d["A_val"].append(float(a["val"]))
            a = a["c_next"]
        d["A_ptr"].append(len(d["A_ind"]))
    return d


# ---------------------------------------------------------------- one case through the reference

def run_case(ref, src, sol, binarize=0):
    g = ref.g
    fn = lambda name, *a: g[name].call(g, list(a))
    P = ref.make(src)
    out = {"problem": problem_arrays(ref, P), "sol": sol, "binarize": binarize}
    npp = fn("npp_create_wksp")
    fn("npp_load_prob", npp, P, 0, sol, 0)
    parm = ref.smcp() if sol == GLP_SOL else ref.iocp(binarize=binarize)
    ret = int(fn("npp_simplex", npp, parm) if sol == GLP_SOL else fn("npp_integer", npp, parm))
    out["ret"] = ret
    n_tse, t = 0, npp["top"]
    while not null(t):
        n_tse += 1
        t = t["link"]
    out["n_tse"] = n_tse
    if ret != 0:
        return out
    lp = ref.call("glp_create_prob")
    fn("npp_build_prob", npp, lp)
    red = problem_arrays(ref, lp)
    red["row_ref"] = [int(x) for x in list(npp["row_ref"])[1:]]
    red["col_ref"] = [int(x) for x in list(npp["col_ref"])[1:]]
    out["reduced"] = red
    m, n = int(lp["m"]), int(lp["n"])
    if m == 0 and n == 0:
        if sol == GLP_SOL:
            lp["pbs_stat"] = lp["dbs_stat"] = 2
            lp["obj_val"] = lp["c0"]
        else:
            lp["mip_stat"] = 5
            lp["mip_obj"] = lp["c0"]
    else:
        # the flow of preprocess_and_solve_lp / _mip on the reduced problem
        fn("glp_scale_prob", lp, 0x80 if sol == GLP_SOL else (0x01 | 0x10 | 0x20 | 0x40))
        fn("glp_adv_basis", lp, 0)
        # the inputs of the hot path proper: scale factors and the crash basis of the reduced problem
        out["prep"] = dict(rii=[float(lp["row"][i]["rii"]) for i in range(1, m + 1)],
                           sjj=[float(lp["col"][j]["sjj"]) for j in range(1, n + 1)],
                           row_stat=[int(lp["row"][i]["stat"]) for i in range(1, m + 1)],
                           col_stat=[int(lp["col"][j]["stat"]) for j in range(1, n + 1)])
        r1 = int(ref.call("glp_simplex", lp, ref.smcp()))
        out["reduced_lp_ret"] = r1
        out["reduced_lp"] = ref.lp_result(lp)
        if not (r1 == 0 and int(lp["pbs_stat"]) == 2 and int(lp["dbs_stat"]) == 2):
            return out
        if sol == GLP_MIP:
            r2 = int(ref.call("glp_intopt", lp, ref.iocp()))
            out["reduced_mip_ret"] = r2
            out["reduced_mip"] = ref.mip_result(lp)
            if int(lp["mip_stat"]) not in (5, 2):
                return out
    if sol == GLP_SOL:
        out["in"] = dict(r_stat=[int(lp["row"][i]["stat"]) for i in range(1, m + 1)],
                         r_dual=[float(lp["row"][i]["dual"]) for i in range(1, m + 1)],
                         c_stat=[int(lp["col"][j]["stat"]) for j in range(1, n + 1)],
                         c_value=[float(lp["col"][j]["prim"]) for j in range(1, n + 1)])
    else:
        out["in"] = dict(c_value=[float(lp["col"][j]["mipx"]) for j in range(1, n + 1)])
    fn("npp_postprocess", npp, lp)
    om, on = int(P["m"]), int(P["n"])
    post = dict(c_value=[float(x) for x in list(npp["c_value"])[1:on + 1]])
    if sol == GLP_SOL:
        post.update(r_stat=[int(x) for x in list(npp["r_stat"])[1:om + 1]],
                    r_pi=[float(x) for x in list(npp["r_pi"])[1:om + 1]],
                    c_stat=[int(x) for x in list(npp["c_stat"])[1:on + 1]])
    out["post"] = post
    fn("npp_unload_sol", npp, P)
    out["unloaded"] = ref.lp_result(P) if sol == GLP_SOL else dict(
        ref.mip_result(P), row_val=[float(P["row"][i]["mipx"]) for i in range(1, om + 1)])
    return out


def terminal_output(ref, src, sol, binarize, msg_lev):
    """what the reference prints for the WHOLE call glp_simplex / glp_intopt(presolve: GLP_ON) at msg_lev"""
    from minijs import NativeFunc, js_to_str
    lines = []
    saved = ref.call("glp_get_print_func")
    ref.call("glp_set_print_func", NativeFunc(lambda this, a: lines.append(js_to_str(a[0]) if a else ""), "print"))
    try:
        P = ref.make(src)
        if sol == GLP_SOL:
            ret = ref.call("glp_simplex", P, ref.smcp(presolve=1, msg_lev=msg_lev))
        else:
            ref.call("glp_simplex", P, ref.smcp(msg_lev=0))
            del lines[:]
            ret = ref.call("glp_intopt", P, ref.iocp(presolve=1, binarize=binarize, msg_lev=msg_lev))
    finally:
        ref.call("glp_set_print_func", saved)
    return dict(ret=int(ret), lines=lines)


def add_terminal_output(ref, case, src):
    """only where the whole call stays on the host: rejected by the presolver or solved by it"""
    if case["ret"] != 0 or (case["reduced"]["m"] == 0 and case["reduced"]["n"] == 0):
        case["terminal"] = {str(lev): terminal_output(ref, src, case["sol"], case["binarize"], lev) for lev in (0, 3)}
    else:
        # the solvers are silent at GLP_MSG_OFF; what still reaches the print function are the messages of
        # glp_scale_prob / glp_adv_basis on the REDUCED problem (the reference's term_out switch is inert)
        case["terminal_off"] = terminal_output(ref, src, case["sol"], case["binarize"], 0)


def main():
    ref = Ref()
    cases = {"_about": "generated by oracle/jsref/make_npp_golden.py from the unmodified reference sources "
                       "(lib/glpnpp01-05.js and their callers, executed by oracle/jsref/minijs.py)"}
    t0 = time.time()
    for name in ("test", "gap", "todd"):
        text = open(os.path.join(refjs.REF, "test", name + ".lpt")).read()
        for sol, tag in ((GLP_SOL, "lp"), (GLP_MIP, "mip")):
            if name == "todd" and sol == GLP_MIP:
                # the B&B of todd.lpt takes hours in the interpreter: presolve + reduced problem only
                c = run_case_presolve_only(ref, text, sol)
            else:
                c = run_case(ref, text, sol)
            cases["%s_%s" % (name, tag)] = c
            print(name, tag, c["ret"], c["n_tse"], c.get("reduced", {}).get("m"), c.get("reduced", {}).get("n"),
                  round(time.time() - t0, 1), flush=True)
    for seed in range(1, 41):
        c = run_case(ref, npp_lp(seed, m=8 + seed % 7, n=10 + seed % 9), GLP_SOL)
        add_terminal_output(ref, c, npp_lp(seed, m=8 + seed % 7, n=10 + seed % 9))
        cases["npp_lp_%d" % seed] = c
        print("npp_lp", seed, c["ret"], c["n_tse"], c.get("reduced", {}).get("m"), c.get("reduced", {}).get("n"),
              c.get("unloaded", {}).get("obj"), round(time.time() - t0, 1), flush=True)
    for seed in range(1, 25):
        c = run_case(ref, npp_mip(seed), GLP_MIP, binarize=seed % 2)
        add_terminal_output(ref, c, npp_mip(seed))
        cases["npp_mip_%d" % seed] = c
        print("npp_mip", seed, c["ret"], c["n_tse"], c.get("reduced", {}).get("m"), c.get("reduced", {}).get("n"),
              c.get("unloaded", {}).get("mip_obj"), round(time.time() - t0, 1), flush=True)
    for seed in range(101, 107):        # larger LPs: a real reduced problem is left for the simplex
        c = run_case(ref, npp_lp(seed, m=40 + 5 * (seed % 4), n=60 + 7 * (seed % 3)), GLP_SOL)
        cases["npp_lp_%d" % seed] = c
        print("npp_lp", seed, c["ret"], c["n_tse"], c.get("reduced", {}).get("m"), c.get("reduced", {}).get("n"),
              c.get("unloaded", {}).get("obj"), round(time.time() - t0, 1), flush=True)
    for seed in range(301, 313):        # objective of any sign: dual infeasibility found by the presolver or by the simplex
        c = run_case(ref, npp_lp(seed, m=7 + seed % 5, n=9 + seed % 6, wild=True), GLP_SOL)
        add_terminal_output(ref, c, npp_lp(seed, m=7 + seed % 5, n=9 + seed % 6, wild=True))
        cases["npp_lp_%d" % seed] = c
        print("npp_lp wild", seed, c["ret"], c["n_tse"], c.get("reduced_lp_ret"), c.get("reduced_lp", {}).get("status"),
              c.get("unloaded", {}).get("obj"), round(time.time() - t0, 1), flush=True)
    for seed in range(201, 205):
        c = run_case(ref, npp_mip(seed, m=14, n=20), GLP_MIP, binarize=seed % 2)
        cases["npp_mip_%d" % seed] = c
        print("npp_mip", seed, c["ret"], c["n_tse"], c.get("reduced", {}).get("m"), c.get("reduced", {}).get("n"),
              c.get("unloaded", {}).get("mip_obj"), round(time.time() - t0, 1), flush=True)
    for seed in (285, 405, 655, 1085):   # integer ranges the binarization refuses or turns into 12 binaries: presolve only
        c = run_case_presolve_only(ref, npp_mip(seed, m=6 + seed % 9, n=9 + seed % 8, wide=True), GLP_MIP, binarize=1)
        cases["npp_wide_%d" % seed] = c
        print("npp_wide", seed, c["ret"], c["n_tse"], c.get("reduced", {}).get("m"), c.get("reduced", {}).get("n"), flush=True)
    with open(os.path.join(GOLD, "ref_npp.json"), "w") as f:
        json.dump(cases, f)
    print("wrote ref_npp.json (%d bytes, %d cases)" % (os.path.getsize(os.path.join(GOLD, "ref_npp.json")),
                                                      len(cases) - 1))


def run_case_presolve_only(ref, src, sol, binarize=0):
    g = ref.g
    fn = lambda name, *a: g[name].call(g, list(a))
    P = ref.make(src)
    out = {"problem": problem_arrays(ref, P), "sol": sol, "binarize": binarize}
    npp = fn("npp_create_wksp")
    fn("npp_load_prob", npp, P, 0, sol, 0)
    out["ret"] = int(fn("npp_integer", npp, ref.iocp(binarize=binarize)))
    n_tse, t = 0, npp["top"]
    while not null(t):
        n_tse += 1
        t = t["link"]
    out["n_tse"] = n_tse
    if out["ret"] == 0:
        lp = ref.call("glp_create_prob")
        fn("npp_build_prob", npp, lp)
        red = problem_arrays(ref, lp)
        red["row_ref"] = [int(x) for x in list(npp["row_ref"])[1:]]
        red["col_ref"] = [int(x) for x in list(npp["col_ref"])[1:]]
        out["reduced"] = red
    return out


if __name__ == "__main__":
    main()
