"""CPU tests of the batched branch-and-bound (csrc/nodeengine.cuh +
csrc/bnbpool.cuh) through its host emulation build (tests/emul, one thread per
CTA): the node logic -- preprocessing, warm-started dual simplex with both
phases, rounding of bounds, integrality, reduced-cost fixing, Driebeck-Tomlin
branching -- and the tree / migration code are the SAME SOURCE that runs on the
device as k_bnb_nodes; the GPU tests (test_gpu_bnb_batch.py) repeat the
comparisons on the device.  Bar: identical MIP optimum and status as the oracle
(and the HiGHS pins) -- the tree differs from the serial one, the optimum not."""
import json
import os

import numpy as np
import pytest

import helpers as H
import ne_emul as NE
import oracle_lib as O


def oracle_root_and_mip(dn, **kw):
    Q = O.Problem.from_arrays(H.to_oracle(dn))
    rc = Q.simplex(meth=O.GLP_PRIMAL)
    root = Q.solution()
    if rc != 0 or root["status"] != O.GLP_OPT:
        return None, None, root
    ret = Q.intopt(**kw)
    return ret, Q.mip(), root


def std_stat(dn):
    """all rows basic, columns at their default non-basic status"""
    st = np.full(dn["m"] + dn["n"], O.GLP_NL, np.int32)
    st[:dn["m"]] = O.GLP_BS
    t = dn["type"][dn["m"]:]
    st[dn["m"]:] = np.where(t == O.GLP_FR, O.GLP_NF, np.where(t == O.GLP_UP, O.GLP_NU, np.where(t == O.GLP_FX, O.GLP_NS, O.GLP_NL)))
    return st


@pytest.mark.parametrize("name", ["gap", "todd"])
@pytest.mark.parametrize("batch", [1, 8])
@pytest.mark.parametrize("start", ["optimal", "standard"])
def test_fixtures(name, batch, start):
    d = H.load_golden(name)
    dn = H.to_native(d)
    oret, omp, root = oracle_root_and_mip(dn)
    stat = root["stat"] if start == "optimal" else std_stat(dn)   # the standard basis exercises dual phase 1
    P = NE.Pool(dn, stat, batch=batch)
    assert P.run() == 0 == oret
    inc = P.incumbent()
    assert inc["have_sol"] and inc["obj"] == omp["mip_obj"] == d["highs_mip_obj"]
    x = inc["x"][dn["m"]:]
    assert np.all(x == np.round(x))
    assert abs(float(dn["coef"] @ x) + dn["c0"] - inc["obj"]) < 1e-9
    np.testing.assert_allclose(H.spmv(dn, x), inc["x"][:dn["m"]], atol=1e-9)
    assert P.stats()["solved"] > 0 and P.open_count() == 0


@pytest.mark.parametrize("seed", range(24))
def test_random_mips_match_oracle_and_highs(seed):
    dn = H.to_native(H.random_mip(seed))
    oret, omp, root = oracle_root_and_mip(dn)
    if oret is None:
        pytest.skip("root LP not optimal")
    with open(os.path.join(H.GOLDEN, "random_pins.json")) as f:
        pins = json.load(f)
    pin = {r["seed"]: r for r in pins["mip"]}.get(seed)
    for batch in (1, 16):
        P = NE.Pool(dn, root["stat"], batch=batch, cap=16384)
        assert P.run() == oret == 0
        inc = P.incumbent()
        if omp["mip_stat"] == O.GLP_NOFEAS:
            assert not inc["have_sol"]
        else:
            assert inc["have_sol"] and abs(inc["obj"] - omp["mip_obj"]) <= 1e-9 * max(1.0, abs(omp["mip_obj"]))
            if pin and pin["highs"] == "optimal":
                assert abs(inc["obj"] - pin["obj"]) <= 1e-7 * max(1.0, abs(pin["obj"]))


@pytest.mark.parametrize("br,bt", [(4, 3), (3, 1), (1, 2), (2, 4)])
def test_small_knapsack_all_rules(br, bt):
    dn = O.generate("mkp", m=5, n=30, seed=20240701)
    oret, omp, root = oracle_root_and_mip(dn, br_tech=br, bt_tech=bt)
    P = NE.Pool(dn, root["stat"], br_tech=br, bt_tech=bt, batch=32)
    assert P.run() == 0 == oret
    assert P.incumbent()["obj"] == omp["mip_obj"]


def test_knapsack_10x40_and_scaled_problem():
    dn = O.generate("mkp", m=10, n=40, seed=3)
    oret, omp, root = oracle_root_and_mip(dn)
    P = NE.Pool(dn, root["stat"], batch=64, cap=65536)
    assert P.run() == 0
    assert P.incumbent()["obj"] == omp["mip_obj"] == 15965.0
    # the same problem with scale factors (powers of two, as GLP_SF_2N leaves them): same optimum
    rng = np.random.default_rng(1)
    rii = 2.0 ** rng.integers(-3, 4, dn["m"])
    sjj = 2.0 ** rng.integers(-3, 4, dn["n"])
    Q = NE.Pool(dn, root["stat"], batch=64, cap=65536, rii=rii, sjj=sjj)
    assert Q.run() == 0
    assert Q.incumbent()["obj"] == 15965.0


def test_node_limit_gap_and_cutoff():
    dn = O.generate("mkp", m=10, n=40, seed=3)
    _, omp, root = oracle_root_and_mip(dn)
    P = NE.Pool(dn, root["stat"], batch=8, node_lim=40)
    rc = P.run()
    assert rc == 13 and 40 <= P.stats()["solved"] <= 40 + 8 * 3        # GLP_ESTOP
    # relative mip gap reached -> GLP_EMIPGAP (lib/glpios03.js:615-625)
    Q = NE.Pool(dn, root["stat"], batch=8, mip_gap=0.05)
    assert Q.run() == 14
    inc = Q.incumbent()
    assert inc["have_sol"] and inc["obj"] <= omp["mip_obj"] and inc["obj"] >= 0.95 * omp["mip_obj"] - 1
    # an incumbent objective learnt from another rank prunes: with the optimum as cut-off nothing better exists
    R = NE.Pool(dn, root["stat"], batch=8)
    R.set_cutoff(omp["mip_obj"])
    assert R.run() == 0
    inc = R.incumbent()
    assert not inc["have_sol"] and inc["have_cut"] and inc["obj"] == omp["mip_obj"]
    assert R.stats()["solved"] < P.stats()["solved"] + 5000


def test_migration_between_two_pools_keeps_the_optimum():
    """two pools share one search: every few rounds the richer one ships half of its
    nodes (every second of the bound order) to the other and both take the better
    incumbent -- the optimum is the serial one"""
    dn = O.generate("mkp", m=10, n=40, seed=3)
    _, omp, root = oracle_root_and_mip(dn)
    A = NE.Pool(dn, root["stat"], batch=4, cap=32768)
    B = NE.Pool(dn, root["stat"], batch=4, cap=32768)
    B.clear()                                  # rank 1 starts empty and is fed by migration
    assert B.open_count() == 0
    moved = 0
    for it in range(100000):
        ra, _ = A.round()
        rb, _ = B.round()
        assert ra in (0, 1) and rb in (0, 1)
        for src, dst in ((A, B), (B, A)):
            if src.open_count() > 2 * dst.open_count() + 4:
                buf, cnt = src.export_nodes((src.open_count() - dst.open_count()) // 2)
                dst.import_nodes(buf, cnt)
                moved += cnt
        ia, ib = A.incumbent(), B.incumbent()
        best = max([x["obj"] for x in (ia, ib) if x["have_cut"]], default=None)
        if best is not None:
            A.set_cutoff(best)
            B.set_cutoff(best)
        if A.open_count() == 0 and B.open_count() == 0:
            break
    assert moved > 0
    objs = [x["obj"] for x in (A.incumbent(), B.incumbent()) if x["have_sol"]]
    assert objs and max(objs) == omp["mip_obj"]


# ---- the sharding protocol of glpk_js_b200.bnb.sharded_bnb_batched over the emulated engine ----
def _mkp_case():
    dn = O.generate("mkp", m=10, n=40, seed=3)
    _, omp, root = oracle_root_and_mip(dn)
    return dn, root["stat"], omp["mip_obj"]


def test_sharded_batched_single_rank_and_node_limit():
    import glpk_js_b200 as G
    dn, stat, want = _mkp_case()
    w = NE.EmulWorker(dn, stat)
    res = G.bnb.sharded_bnb_batched(w, G.bnb.TensorComm(), minimize=False, batch=16, slab_nodes=32768)
    assert res["ret"] == 0 and res["obj"] == want and res["holder"] == 0 and res["open_left"] == 0
    assert w.final["have_sol"] and w.final["obj"] == want
    w2 = NE.EmulWorker(dn, stat)
    res2 = G.bnb.sharded_bnb_batched(w2, G.bnb.TensorComm(), minimize=False, batch=16, slab_nodes=32768, node_lim=200)
    assert res2["ret"] == 0 and 200 <= res2["total_nodes"] <= 200 + 16 * 3 and res2["open_left"] > 0


def _gloo_batched_rank(rank, world, port, q):
    import torch.distributed as dist
    import glpk_js_b200 as G
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dn, stat, want = _mkp_case()
    w = NE.EmulWorker(dn, stat)
    res = G.bnb.sharded_bnb_batched(w, G.bnb.TensorComm(), minimize=False, batch=8, slab_nodes=32768, max_ship=16)
    q.put((rank, res["obj"], res["ret"], res["holder"], res["nodes"], res["moved_in"], res["moved_out"], res["open_left"],
           w.final["have_sol"], w.final["obj"], want))
    dist.destroy_process_group()


def test_sharded_batched_over_gloo_world_size_2():
    """two processes, torch.distributed gloo: rank 1 starts empty and is fed by migration through the
    fused all-gather; both end with the serial optimum and agree on who holds the solution"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 23500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_batched_rank, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    out = sorted(q.get(timeout=300) for _ in range(2))
    [p.join(30) for p in procs]
    want = out[0][10]
    assert all(o[1] == want and o[2] == 0 and o[7] == 0 for o in out), out
    assert out[0][3] == out[1][3] and out[0][3] in (0, 1)
    holder = out[0][3]
    assert out[holder][8] and out[holder][9] == want            # the holder really has the solution vector
    assert out[1][5] > 0 and out[0][6] > 0                     # nodes moved from rank 0 to rank 1
    assert out[0][4] > 0 and out[1][4] > 0                     # both ranks solved node LPs
