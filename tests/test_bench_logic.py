"""CPU tests of bench.py's host logic: the reference arm's contract (and that it
loads nothing of the product), identical `config` in both arms, the roofline
object, and the branch-and-bound block with the device replaced by a scripted
stand-in (an exception there would cost the headline line on the GPU box)."""
import argparse
import contextlib
import importlib.util
import io
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import glpk_js_b200 as G
import helpers as H

ROOT = os.path.abspath(os.path.join(H.HERE, ".."))
_spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(bench)


def test_reference_arm_contract_and_no_product_library():
    """--impl reference in a fresh interpreter: one JSON line, impl/cpu_baseline/e2e keys, zero copies,
    and neither the product package nor its .so loaded (the problem comes from the oracle's generator)"""
    code = ("import runpy, sys, json\n"
            "sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'c2s', '--steps', '1', '--warmup', '0']\n"
            "runpy.run_path(%r, run_name='__main__')\n"
            "maps = open('/proc/self/maps').read()\n"
            "print(json.dumps({'product_module': any(m.startswith('glpk_js_b200') or m.startswith('glpk.js_b200') for m in sys.modules),"
            " 'product_so': 'libglpb200' in maps, 'oracle_so': 'libglpo' in maps}))\n" % os.path.join(ROOT, "bench.py"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 2
    d, probe = json.loads(lines[0]), json.loads(lines[1])
    assert probe == {"product_module": False, "product_so": False, "oracle_so": True}
    assert d["impl"] == "reference" and d["unit"] == "iter/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["e2e"] == {"value": d["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["config"] == bench.config_of(bench.WORKLOADS["c2s"])


def test_reference_arm_other_ranks_do_nothing():
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        bench.run_reference(argparse.Namespace(gpus=2, steps=1, warmup=0), bench.WORKLOADS["c2s"], 1)
    assert out.getvalue() == ""


def test_config_is_identical_in_both_arms_and_names_the_workload():
    for name, w in bench.WORKLOADS.items():
        c = bench.config_of(w)
        assert c["workload"] == w["name"] and set(c) == {"workload"} | set(w["kw"])
        json.dumps(c)
    assert bench.WORKLOADS["c3"]["kw"] == dict(m=16384, n=32768, kmin=8, kspan=17, seed=20240601)


def test_oracle_generator_equals_product_generator():
    """bench's CPU legs build their problems with oracle/gen.cpp; the GPU legs with csrc/gen.cpp"""
    import oracle_lib as O
    for which, kw in (("packing", dict(m=64, n=96, density=0.2, seed=5)), ("covering", dict(m=128, n=256, kmin=8, kspan=17, seed=7)),
                      ("mkp", dict(m=30, n=500, seed=20240701))):
        a, b = O.generate(which, **kw), G.native.generate(which, **kw)
        for k in a:
            assert (np.array_equal(a[k], b[k]) if isinstance(a[k], np.ndarray) else a[k] == b[k]), (which, k)


def test_roofline_object_from_a_profile():
    prof = {"k_engine_dual": dict(count=2, ms=0.3, bytes=0.0), "eng_D2_trow": dict(count=8, ms=0.05, bytes=4.0e5),
            "eng_D4_tcol": dict(count=8, ms=0.10, bytes=3.2e6), "k_beta_rhs": dict(count=2, ms=0.04, bytes=3.0e5),
            "ref_build": dict(count=1, ms=0.01, bytes=0.0)}
    roof, detail = bench.roofline_of(prof, "k_engine_dual", 6453.1, "measured", {"dram_bytes_per_iteration": 1.5e5, "source": "x"})
    assert roof["bound"] == "hbm" and roof["kernel"] == "k_engine_dual" and roof["unit"] == "GB/s"
    assert abs(roof["achieved"] - 3.6e6 / 0.15e-3 / 1e9) < 1e-9 and abs(roof["frac"] - roof["achieved"] / 6453.1) < 1e-12
    assert roof["traffic"] == 1.5e5 and roof["top_phase"]["name"] == "D4_tcol"
    assert abs(roof["bytes_per_launch"] - 3.6e6 / 8) < 1e-9 and abs(roof["us_per_launch"] - 150.0 / 8) < 1e-9
    assert set(detail["phase_table"]) == {"D2_trow", "D4_tcol"} and "build" in detail["refactor_split_ms"]
    assert bench.roofline_of({"k_x": dict(count=1, ms=1.0, bytes=1.0)}, "k_engine_dual", 1.0, "", {}) == (None, None)
    json.dumps(roof)


class _FakeProblem:
    def __init__(self, d, device=0, rii=None, sjj=None):
        self.m, self.n, self.launches = d["m"], d["n"], 0

    def simplex(self, **kw):
        return 0

    def counters(self):
        return dict(launches=self.launches)

    def mip(self):
        x = np.zeros(self.m + self.n)
        return dict(mip_stat=5, mip_obj=0.0, mipx=x, nodes=1)

    def close(self):
        pass


def test_bnb_block_with_a_scripted_device(monkeypatch):
    import torch
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(G.native, "Problem", _FakeProblem)

    class Comm:
        rank, world = 0, 1

    monkeypatch.setattr(G.bnb, "TensorComm", lambda: Comm())
    calls = []

    def fake_search(worker, comm, minimize, batch=0, node_lim=None, **kw):
        calls.append(node_lim)
        worker.P.launches += 10
        if node_lim is None:       # the completion run
            return dict(ret=0, obj=25038.0, holder=0, nodes=1000, total_nodes=1000, rounds=5, exchanges=5, moved_in=0,
                        moved_out=0, open_left=0)
        return dict(ret=0, obj=None, holder=None, nodes=node_lim, total_nodes=node_lim, rounds=7, exchanges=7, moved_in=0,
                    moved_out=0, open_left=99)

    monkeypatch.setattr(G.bnb, "sharded_bnb_batched", fake_search)
    args = argparse.Namespace(bnb_nodes=1234, bnb_batch=0)
    blk = bench.bnb_block(args, 0, 0, 1, 3, 2)
    assert blk["metric"] == "bnb_nodes_per_sec" and blk["unit"] == "nodes/s" and blk["value"] > 0 and blk["n_gpus"] == 1
    assert blk["steps"] == 3 and blk["warmup"] == 2 and blk["nodes_per_step"] == 1234 and blk["gpu_launches"] == 30
    assert calls == [1234] * 5 + [None]
    oc = blk["optimum_check"]
    assert oc["optimum"] == 25038.0 and oc["expected"] == pytest.approx(25038.0) and blk["optimum_ok"] is True
    json.dumps(blk)
