#!/usr/bin/env python
"""node LPs of the serial device branch-and-bound (glpb_mip_begin / run / end) on the reference's fixtures,
next to the reference's own counts (tests/golden/ref_runs.json, presolve OFF)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import glpk_js_b200 as G
import helpers as H
nat = G.native
REF = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_runs.json")))
for name in ("test", "gap", "todd"):
    P = nat.Problem(H.to_native(H.load_golden(name)))
    rc = P.simplex(meth=nat.GLP_PRIMAL)
    P.mip_begin()
    total, state = 0, 1
    while state == 1:
        state, solved = P.mip_run(100000)
        total += solved
    P.mip_end(state)
    mp = P.mip()
    r = REF[name]["presolve_0"]["mip"]
    print(name, "device nodes", total, mp["nodes"], "obj", mp["mip_obj"], "| reference nodes", r["nodes_solved"], "obj", r["mip_obj"], flush=True)
    P.close()
