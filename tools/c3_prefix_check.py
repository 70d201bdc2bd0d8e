"""C3 at full size: device basis after K iterations against the oracle's basis after K iterations (oracle run here)."""
import os, sys, json, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..")); sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import numpy as np
import glpk_js_b200 as G
import oracle_lib as O, helpers as H
nat = G.native
d = nat.generate("covering", m=16384, n=32768, kmin=8, kspan=17, seed=20240601)
for K in [int(x) for x in sys.argv[1:]]:
    Q = O.Problem.from_arrays(H.to_oracle(d))
    t0 = time.time(); Q.simplex(meth=O.GLP_DUAL, it_lim=K); to = time.time() - t0
    ref = np.asarray(Q.solution()["stat"]).astype(int)
    P = nat.Problem(d)
    t0 = time.time(); P.simplex(meth=nat.GLP_DUAL, it_lim=K); td = time.time() - t0
    s = P.solution(); stat = np.asarray(s["stat"]).astype(int)
    P.close()
    print(json.dumps(dict(K=K, oracle_s=round(to, 1), device_s=round(td, 2), basic_in_both=int(np.sum((stat == 1) & (ref == 1))),
                          basic=int(np.sum(ref == 1)), statuses_equal=int(np.sum(stat == ref)), of=len(ref))), flush=True)
