"""development aid: solve one synthetic LP through the C ABI and print counters
usage: dbg_solve.py packing|covering m n [it_lim]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glpk_js_b200 as G
nat = G.native
which, m, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
it_lim = int(sys.argv[4]) if len(sys.argv) > 4 else None
if which == "packing":
    d = nat.generate("packing", m=m, n=n, density=0.2, seed=20240501); meth = nat.GLP_PRIMAL
else:
    d = nat.generate("covering", m=m, n=n, kmin=8, kspan=17, seed=20240601); meth = nat.GLP_DUAL
P = nat.Problem(d)
t0 = time.time()
kw = dict(meth=meth)
if it_lim: kw["it_lim"] = it_lim
rc = P.simplex(**kw)
s = P.solution()
print(which, m, n, "rc", rc, "status", s["status"], "obj", s["obj"], "it", s["it_cnt"], "wall %.2fs" % (time.time() - t0), P.counters(), flush=True)
