"""debug aid: device state after K iterations against the oracle's state at its (K+1)-th pricing"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import glpk_js_b200 as G, oracle_lib as O, helpers as H
nat = G.native
seed, K = int(sys.argv[1]), int(sys.argv[2])
d = H.transport_lp(seed)
Q = O.Problem.from_arrays(d)
cnt = {"n": 0}
snap = {}
def hook(ev, csa):
    if ev == O.EV_P_CHUZC:
        cnt["n"] += 1
        if cnt["n"] == K + 1:
            s = O.csa_scalars(csa)
            for k in ("stat", "cbar", "gamma", "head", "bbar", "coef", "refsp"):
                try: snap[k] = O.csa_get(csa, k).copy()
                except Exception as e: print("no", k, e)
            snap["s"] = s
Q.set_hook(hook); Q.simplex(meth=O.GLP_PRIMAL)
m, n = d["m"], d["n"]
P = nat.Problem(H.to_native(d)); P.simplex(meth=nat.GLP_PRIMAL, it_lim=K)
print("device it", P.solution()["it_cnt"], "oracle q", snap["s"]["q"], "phase", snap["s"]["phase"])
head = P.debug_get("head", m + n).astype(int); cbar = P.debug_get("cbar", n); gamma = P.debug_get("gamma", n); stat = P.debug_get("stat", n).astype(int)
print("head equal", np.array_equal(head, snap["head"][1:]))
for j in range(n):
    a = (stat[j], cbar[j], gamma[j]); b = (int(snap["stat"][j + 1]), snap["cbar"][j + 1], snap["gamma"][j + 1])
    if (a[0], a[2]) != (b[0], b[2]) or abs(b[1]) > 0: print("col", j + 1, "dev", a, "orc", b, "DIFF" if (a[0], a[2]) != (b[0], b[2]) else "")
