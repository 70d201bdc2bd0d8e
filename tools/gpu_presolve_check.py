#!/usr/bin/env python
"""presolve: GLP_ON end to end on the device against the reference's own runs (tests/golden/ref_npp.json,
ref_runs.json): prints, per case, return code / iteration count / objective / status vectors vs the reference."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from glpk_js_b200 import glpk as F
import helpers as H
import test_presolve as T

REF = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_runs.json")))
bad = 0
for name in ("test", "gap", "todd"):
    lp = F.glp_create_prob()
    F.glp_read_lp_from_string(lp, None, H.golden_text(name))
    p = F.SMCP({"presolve": F.GLP_ON}); p.msg_lev = 0
    ret = F.glp_simplex(lp, p)
    r = REF[name]["presolve_1"]["lp"]
    cs = [lp.col[j].stat for j in range(1, lp.n + 1)] == r["col_stat"]
    rs = [lp.row[i].stat for i in range(1, lp.m + 1)] == r["row_stat"]
    dx = max(abs(lp.col[j].prim - r["col_prim"][j - 1]) for j in range(1, lp.n + 1))
    print("fixture", name, "ret", ret, r["ret"], "it", lp.it_cnt, r["it_cnt"], "obj", lp.obj_val, r["obj"], "stat_eq", cs, rs, "dx", dx, flush=True)
    io = F.IOCP({"presolve": F.GLP_ON}); io.msg_lev = 0
    ret = F.glp_intopt(lp, io)
    rm = REF[name]["presolve_1"]["mip"]
    print("   mip ret", ret, rm["ret"], "obj", lp.mip_obj, rm["mip_obj"], "x_eq", [lp.col[j].mipx for j in range(1, lp.n + 1)] == rm["col_val"], flush=True)
for name, case in sorted(T.CASES.items()):
    if not name.startswith("npp_"):
        continue
    P = T.facade_problem(case["problem"])
    if case["sol"] == 1:
        p = F.SMCP({"presolve": F.GLP_ON}); p.msg_lev = 0
        ret = F.glp_simplex(P, p)
        un = case.get("unloaded")
        if un is None or "post" not in case:
            print(name, "ret", ret, "ref presolve ret", case["ret"], case.get("reduced_lp_ret"), flush=True)
            continue
        cs = [P.col[j].stat for j in range(1, P.n + 1)] == un["col_stat"]
        rs = [P.row[i].stat for i in range(1, P.m + 1)] == un["row_stat"]
        dx = max([abs(P.col[j].prim - un["col_prim"][j - 1]) for j in range(1, P.n + 1)] + [0])
        dd = max([abs(P.row[i].dual - un["row_dual"][i - 1]) for i in range(1, P.m + 1)] + [0])
        ok = ret == 0 and cs and rs and dx < 1e-9 and dd < 1e-9 and P.it_cnt == case.get("reduced_lp", {}).get("it_cnt", 0)
        bad += not ok
        print(name, "OK " if ok else "DIFF", "ret", ret, "it", P.it_cnt, case.get("reduced_lp", {}).get("it_cnt"), "obj", P.obj_val, un["obj"], cs, rs, dx, dd, flush=True)
    else:
        io = F.IOCP({"presolve": F.GLP_ON, "binarize": F.GLP_ON if case["binarize"] else F.GLP_OFF}); io.msg_lev = 0
        ret = F.glp_intopt(P, io)
        un = case.get("unloaded")
        if un is None or "post" not in case:
            print(name, "ret", ret, "ref presolve ret", case["ret"], case.get("reduced_lp_ret"), case.get("reduced_mip_ret"), "mip_stat", P.mip_stat, flush=True)
            continue
        xe = [P.col[j].mipx for j in range(1, P.n + 1)] == un["col_val"]
        ok = ret == 0 and P.mip_obj == un["mip_obj"]
        bad += not ok
        print(name, "OK " if ok else "DIFF", "ret", ret, "obj", P.mip_obj, un["mip_obj"], "x_eq", xe, flush=True)
print("cases that differ:", bad)
