"""development aid: B&B on the synthetic knapsack, counters per node"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glpk_js_b200 as G
nat = G.native
m, n, lim = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
d = nat.generate("mkp", m=m, n=n, seed=20240701)
P = nat.Problem(d)
assert P.simplex(meth=nat.GLP_PRIMAL) == 0
c0 = P.counters()
t0 = time.time()
rc = P.intopt(node_lim=lim, msg_lev=0)
dt = time.time() - t0
c1 = P.counters(); mp = P.mip()
nodes = max(1, mp["nodes"])
print("rc", rc, "nodes", mp["nodes"], "wall %.3f s" % dt, "ms/node %.3f" % (1000 * dt / nodes),
      {k: round((c1[k] - c0[k]) / nodes, 2) for k in ("iterations", "refactorizations", "launches", "syncs")}, "obj", mp["mip_obj"])
