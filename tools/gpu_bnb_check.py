import sys, time, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import glpk_js_b200 as G, oracle_lib as O, helpers as H
nat = G.native
def run(dn, name, **kw):
    P = nat.Problem(dn)
    assert P.simplex(meth=nat.GLP_PRIMAL) == 0
    t = time.time(); rc = P.bnb_begin(**kw); assert rc == 0, rc
    while True:
        rc, done = P.bnb_round()
        if rc != 1: break
    st = P.bnb_stats(); dt = time.time() - t
    ret = P.bnb_end(rc); mp = P.mip()
    print(name, "rc", rc, "stat", mp["mip_stat"], "obj", mp["mip_obj"], st, "%.3fs %.0f nodes/s" % (dt, st["solved"]/dt), flush=True)
    P.close(); return mp
for name in ("todd", "gap"):
    d = H.load_golden(name); dn = H.to_native(d)
    mp = run(dn, name); print("   expect", d["highs_mip_obj"])
for (m,n,seed) in ((5,30,20240701),(10,40,3),(30,60,20240701)):
    dn = nat.generate("mkp", m=m, n=n, seed=seed); run(dn, "mkp%dx%d" % (m,n))
for seed in (8, 21, 32):
    dn = H.to_native(H.random_mip(seed)); run(dn, "rand%d" % seed)
dn = nat.generate("mkp", m=30, n=500, seed=20240701)
for batch in (148, 592, 1184):
    run(dn, "mkp30x500 b%d" % batch, node_lim=20000, batch=batch)
