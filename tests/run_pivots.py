"""Helper of test_gpu_engine.py: solves the LPs whose pivot sequences were recorded from the
reference (tests/golden/ref_runs.json) through the C ABI with the pivot log on, in a fresh
process (GLPB_ENGINE is read once per process); one JSON line with the comparison."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402

import glpk_js_b200 as G  # noqa: E402
import helpers as H  # noqa: E402

nat = G.native
VEC = np.load(os.path.join(H.GOLDEN, "ref_vectors.npz"))


def weights_after(dn, meth, name, mname):
    """PSE weights, reduced costs and basic values after K iterations against what the reference held
    when it entered update_gamma of iteration K (ref_vectors.npz, captured from lib/glpspx0[12].js)"""
    out = []
    tag = "%s_%s/%s_gamma_" % (name.replace("random_lp_", "rand"), mname, mname[0])
    ks = sorted({int(k[len(tag):].split("/")[0]) for k in VEC.files if k.startswith(tag)})
    for K in ks:
        if K == 0:
            continue
        ref_gamma = VEC["%s%d/gamma" % (tag, K)]
        ref_head = VEC["%s%d/head" % (tag, K)]
        P = nat.Problem(dn)
        P.simplex(meth=meth, it_lim=K)
        m, n = dn["m"], dn["n"]
        if P.solution()["it_cnt"] != K:
            P.close()
            continue
        head = P.debug_get("head", m + n).astype(int)
        cnt = n if mname == "primal" else m
        gamma = P.debug_get("gamma", cnt)
        P.close()
        # before iteration K the reference had not yet swapped head for pivot K: same basis as the device after K iterations
        same_head = bool(np.array_equal(head, ref_head[1:]))
        err = float(np.max(np.abs(gamma - ref_gamma[1:cnt + 1]) / np.maximum(1.0, np.abs(ref_gamma[1:cnt + 1]))))
        out.append(dict(K=K, same_head=same_head, gamma_rel_err=err))
    return out


with open(os.path.join(H.GOLDEN, "ref_runs.json")) as f:
    REF = json.load(f)["generated"]
with open(os.path.join(H.GOLDEN, "ref_runs.json")) as f:
    FIX = json.load(f)
out = []
# the reference's own fixtures (LP relaxations of gap.lpt / todd.lpt are combinatorial: exact ties in nearly
# every ratio test, settled only by the order sort_tcol / sort_trow leave) and the generated LPs
CASES = [(n, FIX[n]) for n in ("test", "gap", "todd")] + list(REF.items())
for name, e in CASES:
    if name in ("test", "gap", "todd"):
        dn = H.to_native(H.load_golden(name))
    elif name.startswith("random_lp_"):
        dn = H.to_native(H.random_lp(int(name.rsplit("_", 1)[1])))
    elif name in ("packing", "covering"):
        g = dict(e["gen"])
        dn = nat.generate(g.pop("kind"), **g)
    else:
        continue
    for mname, meth in (("primal", nat.GLP_PRIMAL), ("dual", nat.GLP_DUAL), ("dualp", nat.GLP_DUALP)):
        ref = e.get("trace_" + mname)
        if ref is None:
            continue
        P = nat.Problem(dn)
        P.set_pivot_log(4096)
        rc = P.simplex(meth=meth)
        s = P.solution()
        got = P.pivot_log(4096)
        P_ties = P.counters().get("ties", 0)
        P.close()
        want = [[r["q"], r["p"]] for r in ref["pivots"] if r.get("q", 0) != 0 and r.get("p", 0) != 0]
        got = [list(x) for x in got]
        k = next((i for i, (a, b) in enumerate(zip(got, want)) if a != b), min(len(got), len(want)))
        out.append(dict(name=name, meth=mname, rc_ok=bool(rc == ref["ret"] and s["status"] == ref["status"]),
                        obj_ok=bool(ref["status"] != 5 or abs(s["obj"] - ref["obj"]) <= 1e-9 * max(1.0, abs(ref["obj"]))),
                        iterations=int(s["it_cnt"]), ref_iterations=int(ref["it_cnt"]),
                        same_sequence=bool(got == want), first_difference=int(k), ties=int(P_ties),
                        weights=weights_after(dn, meth, name, mname) if mname != "dualp" and name != "gap" else [],
                        around=[got[max(0, k - 1):k + 2], want[max(0, k - 1):k + 2]] if got != want else None))


def oracle_pivots(d, meth):
    """(q, p) per iteration from the oracle (which reproduces the reference's sequences, test_ref_golden.py)"""
    import oracle_lib as O
    Q = O.Problem.from_arrays(d)
    seq, cur = [], {}

    def hook(ev, csa):
        if ev == O.EV_P_CHUZC:
            cur["q"] = O.csa_scalars(csa)["q"]
        elif ev == O.EV_P_CHUZR:
            seq.append([cur["q"], O.csa_scalars(csa)["p"]])
        elif ev == O.EV_D_CHUZR:
            cur["p"] = O.csa_scalars(csa)["p"]
        elif ev == O.EV_D_CHUZC:
            seq.append([O.csa_scalars(csa)["q"], cur["p"]])
    Q.set_hook(hook)
    rc = Q.simplex(meth=meth)
    s = Q.solution()
    return rc, s, [x for x in seq if x[0] != 0 and x[1] != 0]


# exact ties: transportation LPs (unit coefficients, integer data: every quantity is an integer on every
# implementation), the sequence against the oracle's
for seed in range(1, 7):
    d = H.transport_lp(seed)
    dn = H.to_native(d)
    for mname, meth in (("primal", nat.GLP_PRIMAL), ("dual", nat.GLP_DUAL)):
        orc, osol, want = oracle_pivots(d, meth)
        P = nat.Problem(dn)
        P.set_pivot_log(4096)
        rc = P.simplex(meth=meth)
        s = P.solution()
        got = [list(x) for x in P.pivot_log(4096)]
        ties = P.counters().get("ties", 0)
        P.close()
        k = next((i for i, (a, b) in enumerate(zip(got, want)) if a != b), min(len(got), len(want)))
        out.append(dict(name="transport_%d" % seed, meth=mname, rc_ok=bool(rc == orc and s["status"] == osol["status"]),
                        obj_ok=bool(abs(s["obj"] - osol["obj"]) <= 1e-9 * max(1.0, abs(osol["obj"]))),
                        iterations=int(s["it_cnt"]), ref_iterations=int(osol["it_cnt"]), same_sequence=bool(got == want),
                        first_difference=int(k), ties=int(ties), weights=[],
                        around=[got[max(0, k - 1):k + 2], want[max(0, k - 1):k + 2]] if got != want else None))
print(json.dumps(out))
