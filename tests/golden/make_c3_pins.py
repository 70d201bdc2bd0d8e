#!/usr/bin/env python
"""Generates the pins of the two synthetic LPs BASELINE.json names (C2 packing
2048x4096, C3 covering 16384x32768) -- run ONCE in the build container, results
committed as tests/golden/lp_pins.json, c3_mid_basis.npz, c3_oracle_run.json.

  python tests/golden/make_c3_pins.py highs          # independent optima (scipy/HiGHS)
  python tests/golden/make_c3_pins.py oracle-c3      # the oracle's own uninterrupted C3 solve (hours, 1 thread)
  python tests/golden/make_c3_pins.py oracle-c2      # the oracle's C2 solve (minutes)
  python tests/golden/make_c3_pins.py highs-mkp      # HiGHS MIP optima of the knapsacks solved to completion

Problems come from the ORACLE's generator (oracle/gen.cpp); nothing of the
product library is loaded here.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

C2 = dict(which="packing", kw=dict(m=2048, n=4096, density=0.20, seed=20240501))
C3 = dict(which="covering", kw=dict(m=16384, n=32768, kmin=8, kspan=17, seed=20240601))


def to_oracle(d):
    m = d["m"]
    return dict(m=m, n=d["n"], dir=d["dir"], c0=d["c0"], r_type=d["type"][:m], r_lb=d["lb"][:m],
                r_ub=d["ub"][:m], c_type=d["type"][m:], c_lb=d["lb"][m:], c_ub=d["ub"][m:],
                c_coef=d["coef"], c_kind=d["kind"], A_ptr=d["A_ptr"], A_ind=d["A_ind"], A_val=d["A_val"])


def highs(d, method="highs"):
    """method "highs" = HiGHS' choice (dual simplex; fine for C2, did not finish C3 within 2.3 h in the
    build container), "highs-ipm" = interior point with crossover (C3: 329 s)"""
    from scipy.optimize import linprog
    from scipy.sparse import csc_matrix, vstack
    m, n = d["m"], d["n"]
    A = csc_matrix((d["A_val"], d["A_ind"], d["A_ptr"]), shape=(m, n)).tocsr()
    sign = 1.0 if d["dir"] == O.GLP_MIN else -1.0
    t, lb, ub = d["type"], d["lb"], d["ub"]
    up = np.isin(t[:m], (O.GLP_UP, O.GLP_DB, O.GLP_FX))
    lo = np.isin(t[:m], (O.GLP_LO, O.GLP_DB, O.GLP_FX))
    A_ub = vstack([A[up], -A[lo]]).tocsr()
    b_ub = np.concatenate([ub[:m][up], -lb[:m][lo]])
    bounds = []
    for j in range(n):
        tj = t[m + j]
        bounds.append((lb[m + j] if tj in (O.GLP_LO, O.GLP_DB, O.GLP_FX) else None,
                       ub[m + j] if tj in (O.GLP_UP, O.GLP_DB, O.GLP_FX) else None))
    t0 = time.time()
    r = linprog(sign * d["coef"], A_ub=A_ub, b_ub=b_ub, bounds=bounds, method=method,
                options=dict(primal_feasibility_tolerance=1e-9, dual_feasibility_tolerance=1e-9))
    return dict(status=int(r.status), obj=float(sign * r.fun), seconds=time.time() - t0,
                nit=int(getattr(r, "nit", -1)), method=method)


def load_pins():
    p = os.path.join(HERE, "lp_pins.json")
    return json.load(open(p)) if os.path.exists(p) else {}


def save_pins(pins):
    merged = load_pins()            # other modes of this script may have written in the meantime
    for k, v in pins.items():
        merged.setdefault(k, {}).update(v)
    with open(os.path.join(HERE, "lp_pins.json"), "w") as f:
        json.dump(merged, f, indent=1, sort_keys=True)


def oracle_run(cfg, name, meth, snap_every, mid_at):
    d = O.generate(cfg["which"], **cfg["kw"])
    P = O.Problem.from_arrays(to_oracle(d))
    m, n = d["m"], d["n"]
    log, t0 = [], time.perf_counter()
    ev_iter = O.EV_D_ITER if meth == O.GLP_DUAL else O.EV_P_ITER
    state = {"n": 0}

    def hook(ev, csa):
        if ev != ev_iter:
            return
        state["n"] += 1
        it = state["n"]
        if it % snap_every and it != mid_at:
            return
        s = O.csa_scalars(csa)
        lu = O.csa_lu_stats(csa)
        head = O.csa_get(csa, "head")
        k_struct = int(np.count_nonzero(head[1:m + 1] > m))
        log.append(dict(it=it, it_cnt=s["it_cnt"], seconds=time.perf_counter() - t0, k=k_struct, **lu))
        print(log[-1], flush=True)
        if it == mid_at:
            stat = np.zeros(m + n, np.int8)
            nst = O.csa_get(csa, "stat")
            stat[head[1:m + 1] - 1] = O.GLP_BS
            stat[head[m + 1:m + n + 1] - 1] = nst[1:n + 1]
            np.savez_compressed(os.path.join(HERE, name + "_mid_basis.npz"), stat=stat, it=it)
        with open(os.path.join(HERE, name + "_oracle_run.json"), "w") as f:
            json.dump(dict(partial=True, log=log), f)

    P.set_hook(hook)
    rc = P.simplex(meth=meth)
    dt = time.perf_counter() - t0
    s = P.solution()
    out = dict(partial=False, rc=int(rc), status=int(s["status"]), obj=float(s["obj"]), it_cnt=int(s["it_cnt"]),
               seconds=dt, cores=1, log=log, bfd=P.bfd_stats(),
               where="build container, %s" % open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t"))
    with open(os.path.join(HERE, name + "_oracle_run.json"), "w") as f:
        json.dump(out, f, indent=1)
    pins = load_pins()
    pins.setdefault(name, {})["oracle"] = {k: out[k] for k in ("rc", "status", "obj", "it_cnt", "seconds")}
    save_pins(pins)


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "highs":
        pins = load_pins()
        for name, cfg in (("c2", C2), ("c3", C3)):
            d = O.generate(cfg["which"], **cfg["kw"])
            pins.setdefault(name, {})["config"] = dict(cfg["kw"], gen=cfg["which"], nnz=int(d["nnz"]))
            pins[name]["highs"] = highs(d, "highs" if name == "c2" else "highs-ipm")
            print(name, pins[name]["highs"], flush=True)
            save_pins(pins)
    elif what == "highs-mkp":
        # independent MIP optima (scipy.optimize.milp = HiGHS) of the C5-family knapsacks the bench and the
        # tests solve to completion; 30x100 does not finish within 600 s in HiGHS and stays unproven
        from scipy.optimize import milp, LinearConstraint, Bounds
        from scipy.sparse import csc_matrix
        pins = load_pins()
        out = pins.setdefault("mkp", {})
        for (m, n, seed) in ((30, 60, 20240701), (30, 100, 20240701), (10, 40, 3)):
            d = O.generate("mkp", m=m, n=n, seed=seed)
            A = csc_matrix((d["A_val"], d["A_ind"], d["A_ptr"]), shape=(m, n))
            t0 = time.time()
            r = milp(-d["coef"], constraints=LinearConstraint(A, -np.inf, d["ub"][:m]), integrality=np.ones(n),
                     bounds=Bounds(0, 1), options=dict(mip_rel_gap=0.0, time_limit=600))
            out["mkp_%dx%d_seed%d" % (m, n, seed)] = dict(m=m, n=n, seed=seed, highs_status=int(r.status),
                                                          highs_obj=float(-r.fun), seconds=round(time.time() - t0, 2))
            save_pins(pins)
    elif what == "oracle-c3":
        oracle_run(C3, "c3", O.GLP_DUAL, 5000, 60000)
    elif what == "oracle-c2":
        oracle_run(C2, "c2", O.GLP_PRIMAL, 500, 3000)
