"""Per-kernel launch counts and device time per branch-and-bound node (one worker handle)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glpk_js_b200 as G
nat = G.native
d = nat.generate("mkp", m=30, n=500, seed=20240701)
P = nat.Problem(d)
assert P.simplex(meth=nat.GLP_PRIMAL) == 0
c0 = P.counters()
if os.environ.get("PROF", "1") != "0":
    P.set_profile(1)
import time
t0 = time.perf_counter()
P.intopt(node_lim=int(sys.argv[1]) if len(sys.argv) > 1 else 400, msg_lev=0)
dt = time.perf_counter() - t0
prof = P.profile()
c1 = P.counters()
nodes = P.mip()["nodes"]
print("nodes", nodes, "wall ms/node %.3f" % (1e3 * dt / nodes), {k: (c1[k] - c0[k]) / nodes for k in ("launches", "syncs", "iterations", "graph_launches")}, "mip_obj", P.mip()["mip_obj"])
tot = 0
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    if v["count"]:
        print("%-28s count/node %7.2f  us/node %8.2f  us/launch %7.2f" % (k, v["count"] / nodes, 1e3 * v["ms"] / nodes, 1e3 * v["ms"] / v["count"]))
        if not k.startswith("eng_") and not k.startswith("ref_"):
            tot += v["ms"]
print("device us/node (sum of kernels) %.1f" % (1e3 * tot / nodes))
