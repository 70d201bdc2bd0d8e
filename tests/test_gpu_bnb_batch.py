"""GPU parity tests of the batched, device-resident branch-and-bound
(k_bnb_nodes: one CTA per node; glpb_bnb_* and glpb_intopt through the C ABI):
identical MIP optimum and status as the oracle, the HiGHS pins and the
reference's own runs (tests/golden/ref_runs.json)."""
import json
import os

import numpy as np
import pytest

import glpk_js_b200 as G
import helpers as H
import oracle_lib as O

nat, bnb = G.native, G.bnb
pytestmark = pytest.mark.gpu


def oracle_mip(dn, **kw):
    Q = O.Problem.from_arrays(H.to_oracle(dn))
    rc = Q.simplex(meth=O.GLP_PRIMAL)
    if rc != 0 or Q.solution()["status"] != O.GLP_OPT:
        return None, None
    return Q.intopt(**kw), Q.mip()


def run_batched(dn, batch=0, **kw):
    P = nat.Problem(dn)
    assert P.simplex(meth=nat.GLP_PRIMAL) == 0
    assert P.bnb_begin(batch=batch, **kw) == 0
    while True:
        rc, _ = P.bnb_round()
        if rc != 1:
            break
    st = P.bnb_stats()
    ret = P.bnb_end(rc)
    mp = P.mip()
    lp_obj = P.solution()["obj"]
    P.close()
    return ret, mp, st, lp_obj


def check_solution(dn, mp):
    m = dn["m"]
    x = mp["mipx"][m:]
    ints = dn["kind"] == nat.GLP_IV
    assert np.all(x[ints] == np.round(x[ints]))
    assert abs(float(dn["coef"] @ x) + dn["c0"] - mp["mip_obj"]) <= 1e-9 * max(1.0, abs(mp["mip_obj"]))
    ax = H.spmv(dn, x)
    np.testing.assert_allclose(ax, mp["mipx"][:m], atol=1e-7)
    t, lb, ub = dn["type"], dn["lb"], dn["ub"]
    full = np.concatenate([ax, x])
    has_lb, has_ub = np.isin(t, (2, 4, 5)), np.isin(t, (3, 4, 5))
    ubx = np.where(t == 5, lb, ub)
    assert np.all(full[has_lb] >= lb[has_lb] - 1e-6) and np.all(full[has_ub] <= ubx[has_ub] + 1e-6)


@pytest.mark.parametrize("name", ["gap", "todd"])
@pytest.mark.parametrize("batch", [1, 0])
def test_fixtures_optimum_matches_oracle_highs_and_reference(name, batch):
    d = H.load_golden(name)
    dn = H.to_native(d)
    ret, mp, st, lp_obj = run_batched(dn, batch=batch)
    oret, omp = oracle_mip(dn)
    assert ret == oret == 0 and mp["mip_stat"] == omp["mip_stat"] == nat.GLP_OPT
    assert mp["mip_obj"] == omp["mip_obj"] == d["highs_mip_obj"]
    ref = os.path.join(H.GOLDEN, "ref_runs.json")
    if os.path.exists(ref):
        with open(ref) as f:
            rr = json.load(f)[name]["presolve_0"]["mip"]
        assert rr["mip_stat"] == 5 and mp["mip_obj"] == rr["mip_obj"]
    check_solution(dn, mp)
    assert st["solved"] > 0 and st["open"] == 0 and st["a_in_smem"] == 1
    assert abs(lp_obj - d["highs_lp_obj"]) <= 1e-9 * abs(d["highs_lp_obj"])      # the LP relaxation is left as it was


@pytest.mark.parametrize("seed", range(24))
def test_random_mips(seed):
    dn = H.to_native(H.random_mip(seed))
    oret, omp = oracle_mip(dn)
    if oret is None:
        pytest.skip("root LP not optimal")
    ret, mp, st, _ = run_batched(dn)
    assert ret == oret == 0
    assert mp["mip_stat"] == omp["mip_stat"]
    if omp["mip_stat"] == nat.GLP_OPT:
        assert abs(mp["mip_obj"] - omp["mip_obj"]) <= 1e-9 * max(1.0, abs(omp["mip_obj"]))
        check_solution(dn, mp)


@pytest.mark.parametrize("br,bt", [(nat.GLP_BR_DTH, nat.GLP_BT_BLB), (nat.GLP_BR_MFV, nat.GLP_BT_DFS),
                                   (nat.GLP_BR_FFV, nat.GLP_BT_BFS), (nat.GLP_BR_LFV, nat.GLP_BT_BPH)])
def test_knapsack_all_rules_through_glpb_intopt(br, bt):
    """glpb_intopt takes the batched path by itself for problems the node engine takes"""
    dn = nat.generate("mkp", m=5, n=30, seed=20240701)
    P = nat.Problem(dn)
    assert P.simplex(meth=nat.GLP_PRIMAL) == 0
    launches0 = P.counters()["launches"]
    assert P.intopt(br_tech=br, bt_tech=bt) == 0
    mp = P.mip()
    assert mp["mip_stat"] == nat.GLP_OPT and mp["mip_obj"] == 12084.0 and mp["nodes"] > 0
    assert P.counters()["launches"] - launches0 < mp["nodes"]           # many nodes per launch
    P.close()
    oret, omp = oracle_mip(dn, br_tech=br, bt_tech=bt)
    assert omp["mip_obj"] == 12084.0


def test_knapsack_30x60_pinned_by_highs_and_scaled_copy():
    with open(os.path.join(H.GOLDEN, "lp_pins.json")) as f:
        pin = json.load(f)["mkp"]["mkp_30x60_seed20240701"]
    assert pin["highs_status"] == 0
    dn = nat.generate("mkp", m=30, n=60, seed=20240701)
    ret, mp, st, _ = run_batched(dn)
    assert ret == 0 and mp["mip_stat"] == nat.GLP_OPT and abs(mp["mip_obj"] - pin["highs_obj"]) < 1e-6
    check_solution(dn, mp)
    # the same knapsack 10x40 with power-of-two scale factors: same optimum through the scaled path
    d2 = nat.generate("mkp", m=10, n=40, seed=3)
    rng = np.random.default_rng(1)
    rii, sjj = 2.0 ** rng.integers(-3, 4, d2["m"]), 2.0 ** rng.integers(-3, 4, d2["n"])
    P = nat.Problem(d2, rii=rii, sjj=sjj)
    assert P.simplex(meth=nat.GLP_PRIMAL) == 0 and P.intopt() == 0
    assert P.mip()["mip_obj"] == 15965.0
    P.close()


def test_limits_gap_and_serial_switch(monkeypatch):
    dn = nat.generate("mkp", m=10, n=40, seed=3)
    P = nat.Problem(dn)
    assert P.simplex(meth=nat.GLP_PRIMAL) == 0
    assert P.intopt(node_lim=50) == nat.GLP_ESTOP
    assert P.intopt(mip_gap=0.05) == nat.GLP_EMIPGAP           # lib/glpios03.js:615-625
    mp = P.mip()
    assert mp["mip_stat"] == nat.GLP_FEAS and 0.95 * 15965.0 - 1 <= mp["mip_obj"] <= 15965.0
    monkeypatch.setenv("GLPB_BNB", "serial")                    # the one-LP-at-a-time driver of mip.cu
    assert P.intopt() == 0 and P.mip()["mip_obj"] == 15965.0
    P.close()


def test_migration_between_two_handles_through_device_buffers():
    """two pools on one GPU share a search; node records travel device-to-device through a torch tensor"""
    import torch
    dn = nat.generate("mkp", m=10, n=40, seed=3)
    A, B = nat.Problem(dn), nat.Problem(dn)
    for P in (A, B):
        assert P.simplex(meth=nat.GLP_PRIMAL) == 0 and P.bnb_begin(batch=8) == 0
    B.bnb_clear()
    rb = A.bnb_record_bytes()
    assert rb == B.bnb_record_bytes() and rb >= 18 * (dn["m"] + dn["n"]) + 32
    buf = torch.zeros(64 * rb, dtype=torch.uint8, device="cuda")
    moved = 0
    for it in range(100000):
        ra, _ = A.bnb_round()
        rbb, _ = B.bnb_round()
        assert ra in (0, 1) and rbb in (0, 1)
        for src, dst in ((A, B), (B, A)):
            if src.bnb_open_count() > 2 * dst.bnb_open_count() + 4:
                cnt = src.bnb_export(min(64, (src.bnb_open_count() - dst.bnb_open_count()) // 2), buf.data_ptr())
                dst.bnb_import(buf.data_ptr(), cnt)
                moved += cnt
        objs = [P.bnb_incumbent()[1] for P in (A, B) if abs(P.bnb_incumbent()[1]) < 1e300]
        if objs:
            A.bnb_set_cutoff(max(objs))
            B.bnb_set_cutoff(max(objs))
        if A.bnb_open_count() == 0 and B.bnb_open_count() == 0:
            break
    assert moved > 0
    A.bnb_end(0)
    B.bnb_end(0)
    best = [P.mip() for P in (A, B) if P.mip()["mip_stat"] == nat.GLP_OPT]
    assert best and max(b["mip_obj"] for b in best) == 15965.0
    A.close()
    B.close()


def test_sharded_driver_single_rank_on_device():
    dn = nat.generate("mkp", m=10, n=40, seed=3)
    P = nat.Problem(dn)
    assert P.simplex(meth=nat.GLP_PRIMAL) == 0
    res = bnb.sharded_bnb_batched(bnb.BatchWorker(P), bnb.TensorComm(), minimize=False)
    assert res["ret"] == 0 and res["obj"] == 15965.0 and res["holder"] == 0 and res["open_left"] == 0
    assert P.mip()["mip_stat"] == nat.GLP_OPT and P.mip()["mip_obj"] == 15965.0
    Q = nat.Problem(nat.generate("mkp", m=30, n=500, seed=20240701))
    assert Q.simplex(meth=nat.GLP_PRIMAL) == 0
    r2 = bnb.sharded_bnb_batched(bnb.BatchWorker(Q), bnb.TensorComm(), minimize=False, node_lim=5000)
    assert r2["ret"] == 0 and 5000 <= r2["total_nodes"] < 5000 + 3 * 2368 * 2 and r2["open_left"] > 0
    P.close()
    Q.close()
