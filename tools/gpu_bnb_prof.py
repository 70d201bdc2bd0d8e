"""A few rounds of the batched branch-and-bound on the knapsack of BASELINE.json configs[4]
(for ncu captures of k_bnb_nodes).  Usage: python tools/gpu_bnb_prof.py [rounds] [batch]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import glpk_js_b200 as G
nat = G.native
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 592
P = nat.Problem(nat.generate("mkp", m=30, n=500, seed=20240701))
assert P.simplex(meth=nat.GLP_PRIMAL) == 0
assert P.bnb_begin(batch=batch) == 0
for r in range(rounds):
    rc, done = P.bnb_round()
    if rc != 1:
        break
print(P.bnb_stats())
P.bnb_end(13)
P.close()
