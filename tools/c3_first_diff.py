"""C3 at full size: first iteration at which the device's (q, p) differs from the oracle's"""
import os, sys, json, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..")); sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import numpy as np
import glpk_js_b200 as G
import oracle_lib as O, helpers as H
nat = G.native
K = int(sys.argv[1]) if len(sys.argv) > 1 else 11000
d = nat.generate("covering", m=16384, n=32768, kmin=8, kspan=17, seed=20240601)
Q = O.Problem.from_arrays(H.to_oracle(d))
seq, cur = [], {}
def hook(ev, csa):
    if ev == O.EV_D_CHUZR:
        s = O.csa_scalars(csa); cur["p"] = s["p"]; cur["delta"] = s["delta"]
    elif ev == O.EV_D_CHUZC:
        s = O.csa_scalars(csa); seq.append((s["q"], cur["p"], cur["delta"], s["new_dq"]))
Q.set_hook(hook)
Q.simplex(meth=O.GLP_DUAL, it_lim=K)
P = nat.Problem(d); P.set_pivot_log(K + 16)
P.simplex(meth=nat.GLP_DUAL, it_lim=K)
got = P.pivot_log(K + 16); cnt = P.counters(); P.close()
want = [(a, b) for (a, b, _, _) in seq]
k = next((i for i, (a, b) in enumerate(zip(got, want)) if tuple(a) != tuple(b)), None)
print(json.dumps(dict(K=K, oracle_iters=len(want), device_iters=len(got), first_difference=k, counters=cnt,
                      device=[list(x) for x in got[k - 2:k + 3]] if k is not None else None,
                      oracle=[list(x) for x in seq[k - 2:k + 3]] if k is not None else None)))
