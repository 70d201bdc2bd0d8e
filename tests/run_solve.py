"""Helper of test_gpu_engine.py: one solve through the C ABI in a fresh process
(the library reads its GLPB_* tuning variables once per process), result as a
JSON line on stdout.

    run_solve.py packing|covering m n [seed]
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import glpk_js_b200 as G  # noqa: E402
import helpers as H  # noqa: E402

nat = G.native
which, m, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
seed = int(sys.argv[4]) if len(sys.argv) > 4 else None
if which == "packing":
    d = nat.generate("packing", m=m, n=n, density=0.2, seed=seed or 20240501)
    meth = nat.GLP_PRIMAL
else:
    d = nat.generate("covering", m=m, n=n, kmin=8, kspan=17, seed=seed or 20240601)
    meth = nat.GLP_DUAL
P = nat.Problem(d)
rc = P.simplex(meth=meth)
s = P.solution()
r = H.kkt(d, s)
c = P.counters()
print(json.dumps({"rc": int(rc), "status": int(s["status"]), "obj": float(s["obj"]), "it": int(s["it_cnt"]),
                  "kkt": r, "refac": int(c["refactorizations"]), "k": int(c["k"]), "launches": int(c["launches"]),
                  "solve_us": int(c["solve_us"])}))
