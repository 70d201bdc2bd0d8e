"""C3 at full size: the basis the device holds after K iterations against the basis the ORACLE held after the same
number of iterations of its own uninterrupted run (tests/golden/c3_mid_basis.npz, K = 60000)."""
import os, sys, json, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..")); sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import numpy as np
import glpk_js_b200 as G
nat = G.native
z = np.load(os.path.join(HERE, "..", "tests", "golden", "c3_mid_basis.npz"))
K = int(z["it"]); ref = z["stat"].astype(int)
d = nat.generate("covering", m=16384, n=32768, kmin=8, kspan=17, seed=20240601)
P = nat.Problem(d)
t0 = time.time()
rc = P.simplex(meth=nat.GLP_DUAL, it_lim=K)
s = P.solution()
stat = np.asarray(s["stat"]).astype(int)
BS = 1
same_basis = int(np.sum((stat == BS) == (ref == BS)))
print(json.dumps(dict(rc=int(rc), it=int(s["it_cnt"]), seconds=round(time.time() - t0, 2), m_plus_n=len(ref),
                      basic_ref=int(np.sum(ref == BS)), basic_dev=int(np.sum(stat == BS)),
                      basic_in_both=int(np.sum((stat == BS) & (ref == BS))),
                      statuses_equal=int(np.sum(stat == ref)), same_basis_flags=same_basis)))
P.close()
