"""device basis after K iterations of C3 against the committed oracle bases tests/golden/c3_basis_K.npz / c3_mid_basis.npz"""
import os, sys, json, glob
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..")); sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import numpy as np
import glpk_js_b200 as G
nat = G.native
d = nat.generate("covering", m=16384, n=32768, kmin=8, kspan=17, seed=20240601)
for f in sorted(glob.glob(os.path.join(HERE, "..", "tests", "golden", "c3_basis_*.npz"))) + [os.path.join(HERE, "..", "tests", "golden", "c3_mid_basis.npz")]:
    z = np.load(f); K = int(z["it"]); ref = z["stat"].astype(int)
    P = nat.Problem(d); P.simplex(meth=nat.GLP_DUAL, it_lim=K); s = P.solution(); P.close()
    stat = np.asarray(s["stat"]).astype(int)
    print(json.dumps(dict(K=K, it=int(s["it_cnt"]), basic_in_both=int(np.sum((stat == 1) & (ref == 1))), statuses_equal=int(np.sum(stat == ref)), of=len(ref))), flush=True)
