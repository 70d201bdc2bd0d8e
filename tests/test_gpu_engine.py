"""GPU tests of the persistent iteration engine and the cooperative
refactorisation: every execution mode (replicated vs grid-wide ratio test,
direct vs deferred basis changes, shared-memory vs distributed panel, grid
sizes, per-kernel path) must reach the same optimum as the oracle, with KKT
residuals <= 1e-9, and take (nearly) the same number of iterations; plus the
size-independent properties at BASELINE.json's full sizes."""
import json
import os
import subprocess
import sys

import pytest

import glpk_js_b200 as G
import oracle_lib as O
import helpers as H

nat = G.native
pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

# objective of C2 (packing LP 2048 x 4096, seed 20240501) from a full solve of the
# oracle (228 s on the build container); the device must reproduce it to 1e-9
C2_OBJ = 89464.4101453462
# independent optima of both benchmark LPs from HiGHS (tests/golden/make_c3_pins.py)
with open(os.path.join(HERE, "golden", "lp_pins.json")) as _f:
    LP_PINS = json.load(_f)


def solve_in_subprocess(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(HERE, "run_solve.py")] + [str(a) for a in args],
                         env=e, capture_output=True, text=True, timeout=timeout)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


MODES = [
    ("default", {}),
    ("grid-wide ratio test + deferred basis changes", {"GLPB_LOCAL_MAX": "0"}),
    ("grid-wide ratio test, direct basis changes", {"GLPB_LOCAL_MAX": "0", "GLPB_DEFER": "0"}),
    ("eight CTAs, TMA-staged T*v stream once k >= 256", {"GLPB_GRID": "8", "GLPB_TMA": "1"}),
    ("eight CTAs, register-staged stream", {"GLPB_GRID": "8"}),
    ("three CTAs", {"GLPB_GRID": "3"}),
    ("one CTA", {"GLPB_GRID": "1"}),
    ("basis header read from global memory (no per-CTA copies)", {"GLPB_HDR": "0"}),
    ("distributed panel", {"GLPB_REF_SINGLE": "0"}),
    ("distributed panel held by two CTAs", {"GLPB_REF_SINGLE": "0", "GLPB_REF_PG": "2"}),
    ("refactorise every nfs_max updates", {"GLPB_REFAC_AUTO": "0"}),
    ("per-kernel path", {"GLPB_ENGINE": "0"}),
]


@pytest.mark.parametrize("which,m,n", [("covering", 1024, 2048), ("packing", 256, 512)])
def test_every_mode_reaches_the_oracle_optimum(which, m, n):
    if which == "packing":
        d = nat.generate("packing", m=m, n=n, density=0.2, seed=20240501)
        meth = O.GLP_PRIMAL
    else:
        d = nat.generate("covering", m=m, n=n, kmin=8, kspan=17, seed=20240601)
        meth = O.GLP_DUAL
    Q = O.Problem.from_arrays(H.to_oracle(d))
    assert Q.simplex(meth=meth) == 0
    o = Q.solution()
    its = {}
    for name, env in MODES:
        r = solve_in_subprocess([which, m, n], env)
        assert (r["rc"], r["status"]) == (0, o["status"]), (name, r)
        assert abs(r["obj"] - o["obj"]) <= 1e-9 * max(1.0, abs(o["obj"])), (name, r["obj"], o["obj"])
        assert max(r["kkt"].values()) <= 1e-9, (name, r["kkt"])
        its[name] = r["it"]
    # the modes differ from each other only in summation order (same pivots up to the odd
    # tie); against the oracle the inverse is built in another pivot order, so on these
    # degenerate LPs ties fall differently now and then: a few percent either way
    ref = its["default"]
    for name, it in its.items():
        assert abs(it - ref) <= 0.05 * ref + 5, (name, its)
        assert abs(it - o["it_cnt"]) <= 0.10 * o["it_cnt"] + 5, (name, its, o["it_cnt"])


def test_c2_full_size_objective_and_kkt():
    """BASELINE.json configs[1] at full size: optimum pinned by the oracle's full solve"""
    r = solve_in_subprocess(["packing", 2048, 4096])
    assert (r["rc"], r["status"]) == (0, O.GLP_OPT)
    assert abs(r["obj"] - C2_OBJ) <= 1e-9 * C2_OBJ, r["obj"]
    hp = LP_PINS["c2"]["highs"]
    assert hp["status"] == 0 and abs(r["obj"] - hp["obj"]) <= 1e-9 * abs(hp["obj"]), (r["obj"], hp["obj"])   # independent solver
    assert max(r["kkt"].values()) <= 1e-9, r["kkt"]
    assert r["launches"] < r["it"], "the iterations must run inside the persistent engine"


def test_c3_full_size_properties():
    """BASELINE.json configs[2] at full size (the oracle would need the better part
    of an hour): optimality through the size-independent properties -- status,
    primal/dual feasibility and the reduced-cost equation (glp_check_kkt
    residuals), with the basis changes deferred and the distributed panel in use"""
    r = solve_in_subprocess(["covering", 16384, 32768], timeout=900)
    assert (r["rc"], r["status"]) == (0, O.GLP_OPT)
    assert max(r["kkt"].values()) <= 1e-9, r["kkt"]
    hp = LP_PINS["c3"]["highs"]           # HiGHS interior point + crossover, 329 s in the build container
    assert hp["status"] == 0 and abs(r["obj"] - hp["obj"]) <= 1e-9 * abs(hp["obj"]), (r["obj"], hp["obj"])
    assert r["k"] > 2560, "kernel larger than the shared-memory panel: distributed mode was exercised"
    # same optimum whichever way the basis changes are applied
    r2 = solve_in_subprocess(["covering", 16384, 32768], {"GLPB_DEFER": "0", "GLPB_REFAC_DIV": "4"}, timeout=900)
    assert (r2["rc"], r2["status"]) == (0, O.GLP_OPT)
    assert abs(r["obj"] - r2["obj"]) <= 1e-9 * abs(r["obj"]), (r["obj"], r2["obj"])


# ---- engine wiring: the (q, p) sequence of whole solves against THE REFERENCE'S own pivot sequence ----
@pytest.mark.parametrize("engine", ["1", "0"])
def test_pivot_sequence_equals_the_references(engine):
    """glpb_simplex with the pivot log on (tests/run_pivots.py, fresh process): the entering / leaving pair
    of every iteration equals the one lib/glpspx01.js / glpspx02.js chose on the same LP
    (tests/golden/ref_runs.json, generated by running the reference), for the persistent engine
    (GLPB_ENGINE=1) and the per-kernel path (=0).  The generated problems have continuous random data (no
    exact ties); the LP relaxations of the reference's fixtures gap.lpt / todd.lpt are combinatorial and tie in
    nearly every ratio test: there the sequence is the reference's only because ties are settled in the order
    sort_tcol / sort_trow leave (k_sort_list; the engine hands tied iterations to the list-ordered path)."""
    e = dict(os.environ)
    e["GLPB_ENGINE"] = engine
    out = subprocess.run([sys.executable, os.path.join(HERE, "run_pivots.py")], capture_output=True, text=True, env=e, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert len(res) >= 12
    bad = [r for r in res if not (r["rc_ok"] and r["obj_ok"])]
    assert not bad, bad
    # gap.lpt apart (its ties are exact in the reference only because the reference's LU
    # arithmetic repeats bit for bit on two structurally identical rows; same optimum over another path)
    differ = [r for r in res if not r["same_sequence"] and r["name"] != "gap"]
    assert len(differ) <= 1, differ                 # identical pivots, iteration for iteration
    exact = [r for r in res if r["name"] in ("test", "todd") or r["name"].startswith("transport_")]
    assert len(exact) == 18 and not [r for r in exact if not r["same_sequence"]], [r for r in exact if not r["same_sequence"]]
    assert all(r["iterations"] == r["ref_iterations"] for r in exact)
    if engine == "1":
        # the engine did meet exact ties and handed those iterations to the list-ordered path
        assert sum(r["ties"] for r in exact) >= 20
    assert sum(r["iterations"] for r in res) >= 400
    # update_gamma (lib/glpspx01.js:1178-1255, lib/glpspx02.js:1075-1188): the projected steepest edge weights the
    # device holds after K iterations equal the ones the reference held at the same point, basis header included
    w = [x for r in res for x in r["weights"]]
    assert len(w) >= 20 and all(x["same_head"] for x in w), [x for x in w if not x["same_head"]][:3]
    assert max(x["gamma_rel_err"] for x in w) <= 1e-9, sorted(w, key=lambda x: -x["gamma_rel_err"])[:3]


def test_c3_full_size_pivot_sequence_equals_the_oracles():
    """BASELINE.json configs[2] at FULL size (16384 x 32768): the first 5000 iterations take the oracle's (q, p) one for
    one and leave the oracle's basis -- all 49152 statuses (the oracle needs 7 s for them, the device 0.4 s).  Measured
    once beyond that (tools/c3_first_diff.py, tools/c3_prefix_check.py): identical through iteration 10705, basis
    identical after 10000 iterations; at iteration 10706 dual pricing picks another row and the paths separate (76 % of
    the basis still shared after 60000 iterations, same optimum, same 120552 iterations in total on the device)."""
    import numpy as np
    K = 5000
    d = nat.generate("covering", m=16384, n=32768, kmin=8, kspan=17, seed=20240601)
    Q = O.Problem.from_arrays(H.to_oracle(d))
    seq, cur = [], {}

    def hook(ev, csa):
        if ev == O.EV_D_CHUZR:
            cur["p"] = O.csa_scalars(csa)["p"]
        elif ev == O.EV_D_CHUZC:
            seq.append((O.csa_scalars(csa)["q"], cur["p"]))
    Q.set_hook(hook)
    Q.simplex(meth=O.GLP_DUAL, it_lim=K)
    Q.set_hook(None)
    ref = np.asarray(Q.solution()["stat"]).astype(int)
    P = nat.Problem(d)
    P.set_pivot_log(K + 8)
    P.simplex(meth=nat.GLP_DUAL, it_lim=K)
    got = [tuple(x) for x in P.pivot_log(K + 8)]
    stat = np.asarray(P.solution()["stat"]).astype(int)
    cnt = P.counters()
    P.close()
    assert len(seq) == K and len(got) == K
    first = next((i for i, (a, b) in enumerate(zip(got, seq)) if a != b), None)
    assert first is None, (first, got[first - 1:first + 2], seq[first - 1:first + 2])
    assert np.array_equal(stat, ref)
    assert cnt["launches"] < K // 10, "the iterations ran inside the persistent engine"
