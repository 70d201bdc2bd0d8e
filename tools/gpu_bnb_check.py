"""Quick device check of the batched branch-and-bound: optima on the fixtures and the
knapsack family, nodes/s at several batch sizes, and where the node time goes
(SM cycles of thread 0 per phase).  Usage (GPU box): python tools/gpu_bnb_check.py"""
import sys, time, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import glpk_js_b200 as G, helpers as H
nat = G.native
def run(dn, name, **kw):
    P = nat.Problem(dn)
    assert P.simplex(meth=nat.GLP_PRIMAL) == 0
    t = time.time(); rc = P.bnb_begin(**kw); assert rc == 0, rc
    while True:
        rc, done = P.bnb_round()
        if rc != 1: break
    st = P.bnb_stats(); dt = time.time() - t
    P.bnb_end(rc); mp = P.mip()
    cyc = {k[4:]: st[k] for k in st if k.startswith("cyc_")}
    tot = sum(cyc.values()) or 1
    print("%-16s rc %d stat %d obj %.10g solved %d tasks %d rounds %d iters %d refacs %d | %.3fs %.0f nodes/s | us/node %.1f | %s" % (
        name, rc, mp["mip_stat"], mp["mip_obj"], st["solved"], st["tasks"], st["rounds"], st["iters"], st["refacs"], dt,
        st["solved"] / dt, tot / 1965.0 / max(1, st["tasks"]), " ".join("%s %.0f%%" % (k, 100.0 * v / tot) for k, v in cyc.items())), flush=True)
    P.close(); return mp
for name in ("todd", "gap"):
    d = H.load_golden(name); run(H.to_native(d), name); print("   expect", d["highs_mip_obj"])
for (m, n, seed) in ((5, 30, 20240701), (10, 40, 3), (30, 60, 20240701)):
    run(nat.generate("mkp", m=m, n=n, seed=seed), "mkp%dx%d" % (m, n))
dn = nat.generate("mkp", m=30, n=500, seed=20240701)
for batch in (148, 592, 1184, 2368):
    run(dn, "mkp30x500 b%d" % batch, node_lim=100000, batch=batch)
