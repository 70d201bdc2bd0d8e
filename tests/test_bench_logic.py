"""CPU test of bench.py's branch-and-bound leg (control flow, JSON contract) with the
device replaced by a scripted stand-in: the default line embeds this leg as its `bnb`
block, so an exception there would cost the headline line on the GPU box."""
import argparse
import importlib.util
import io
import json
import os
import contextlib

import numpy as np
import pytest

import glpk_js_b200 as G
import helpers as H

_spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(H.HERE, "..", "bench.py"))
bench = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(bench)


class _FakeProblem:
    created = 0

    def __init__(self, d, device=0, rii=None, sjj=None):
        type(self).created += 1
        self.m, self.n = d["m"], d["n"]
        self.c = dict(iterations=0, refactorizations=0, launches=0, syncs=0, updates=0, k=0, solve_us=0, graph_launches=0)
        self.prof = False

    def simplex(self, **kw):
        self.c["launches"] += 40
        return 0

    def counters(self):
        return dict(self.c)

    def set_profile(self, on):
        self.prof = bool(on)

    def profile(self):
        return {"k_engine_dual": dict(count=2, ms=0.3, bytes=0.0), "eng_D2_trow": dict(count=8, ms=0.05, bytes=4.0e5),
                "eng_D3_gemvN_tcol_head": dict(count=8, ms=0.06, bytes=1.0e4), "k_beta_rhs": dict(count=2, ms=0.04, bytes=3.0e5),
                "ref_build": dict(count=1, ms=0.01, bytes=0.0)}

    def close(self):
        pass


@pytest.fixture
def stubbed(monkeypatch):
    import torch
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(G.native, "Problem", _FakeProblem)

    def fake_search(worker, comm, minimize, node_lim=None, **kw):
        worker.P.c["launches"] += 50 * node_lim
        worker.P.c["graph_launches"] += 2 * node_lim
        worker.P.c["syncs"] += 5 * node_lim
        worker.P.c["iterations"] += 8 * node_lim
        return dict(total_nodes=node_lim * comm.world, obj=123.0)      # the driver reports the global node count

    monkeypatch.setattr(G.bnb, "sharded_intopt", fake_search)
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    _FakeProblem.created = 0
    return bench


def _args(**kw):
    a = argparse.Namespace(gpus=1, steps=3, warmup=3, bnb_workers=2, no_cpu_baseline=True)
    a.__dict__.update(kw)
    return a


def test_bnb_leg_embedded_returns_the_block(stubbed):
    line = stubbed.run_bnb(_args(steps=5, warmup=4), stubbed.WORKLOADS["mkp"], 0, 0, 1, embedded=True)
    assert line["metric"] == "bnb_nodes_per_sec" and line["unit"] == "nodes/s" and line["n_gpus"] == 1
    assert line["steps"] == 3 and line["warmup"] == 3                  # capped inside the default line
    assert line["value"] > 0 and line["config"]["workers_per_gpu"] == 2
    assert line["gpu_launches"] == 3 * 2 * 50 * 400                     # timed steps x handles x launches
    assert line["per_node"]["syncs_per_node"] == 5.0 and line["per_node"]["graph_launches_per_node"] == 2.0
    assert line["roofline"]["kernel"] == "k_engine_dual" and line["roofline"]["bound"] == "hbm"
    assert line["cpu_baseline"] is None and line["incumbent"] == 123.0
    json.dumps(line)                                                    # serialisable
    assert _FakeProblem.created == 6 * 2 + 1                            # (warm-up + steps) x handles + profiling pass


def test_bnb_leg_standalone_prints_one_json_line(stubbed):
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        ret = stubbed.run_bnb(_args(steps=2, warmup=1, no_cpu_baseline=False), stubbed.WORKLOADS["mkp"], 0, 0, 1)
    assert ret is None
    lines = [l for l in out.getvalue().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["steps"] == 2 and d["warmup"] == 1 and d["metric"] == "bnb_nodes_per_sec"
    # the CPU leg is the real oracle: 400 nodes of the same search
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] > 0


def test_reference_arm_contract():
    """--impl reference: the oracle timed on a bounded sample; line carries impl, cpu_baseline, e2e with zero copies"""
    out = io.StringIO()
    a = argparse.Namespace(gpus=1, steps=1, warmup=0)
    w = dict(bench.WORKLOADS["c2s"])
    w["cpu_it_lim"], w["cpu_mid"], w["cpu_mid_lim"] = 60, 80, 20
    with contextlib.redirect_stdout(out):
        bench.run_reference(a, w, 0, 1)
    d = json.loads(out.getvalue().strip())
    assert d["impl"] == "reference" and d["unit"] == "iter/s" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    with contextlib.redirect_stdout(io.StringIO()) as other:
        bench.run_reference(a, w, 1, 2)                                 # other ranks: no work, no output
    assert other.getvalue() == ""


def test_default_line_carries_the_bnb_block(stubbed, monkeypatch):
    """main() end to end with the device stubbed out: one JSON line with every key of the
    contract, the `bnb` block embedded, and a failure inside that block contained."""
    import sys
    import torch

    class FakeLP(_FakeProblem):
        def __init__(self, d, device=0, rii=None, sjj=None):
            super().__init__(d, device, rii, sjj)
            self.it = 0

        def std_basis(self):
            return 0

        def simplex(self, meth=None, it_lim=None, **kw):
            self.it += 100
            self.c.update(launches=self.c["launches"] + 30, solve_us=50000, refactorizations=3, k=17)
            return 0

        def solution(self):
            return dict(it_cnt=self.it, status=5, obj=1.5, stat=np.zeros(self.m + self.n, np.int32))

        def profile(self):
            return {"k_engine_primal": dict(count=1, ms=40.0, bytes=0.0), "eng_PA_tcol_head": dict(count=100, ms=10.0, bytes=5.0e7),
                    "eng_PE_trow_svec": dict(count=100, ms=20.0, bytes=2.0e9), "k_refactor": dict(count=3, ms=5.0, bytes=1.0e8),
                    "ref_build": dict(count=3, ms=0.5, bytes=0.0)}

    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(16, dtype=k.get("dtype", torch.uint8)))
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(G.native, "Problem", FakeLP)

    real_lib = G.native.load()

    class FakeLib:
        """the real library (generators, host code) that claims one device"""
        def __getattr__(self, name):
            return getattr(real_lib, name)

        @staticmethod
        def glpb_device_count():
            return 1
    fake_lib = FakeLib()
    monkeypatch.setattr(G.native, "load", lambda: fake_lib)
    monkeypatch.setattr(bench.ClockSampler, "stop", lambda self: {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "samples": 3, "reasons": []})
    for env in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(env, raising=False)

    def run(extra=()):
        monkeypatch.setattr(sys, "argv", ["bench.py", "--no-c3", "--no-cpu-baseline", "--steps", "2", "--warmup", "3",
                                          "--bnb-workers", "2", *extra])
        out = io.StringIO()
        with contextlib.redirect_stdout(out):
            bench.main()
        lines = [l for l in out.getvalue().splitlines() if l.strip()]
        assert len(lines) == 1
        return json.loads(lines[0])

    d = run()
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["metric"] == "simplex_iterations_per_sec" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["steps"] == 2 and d["warmup"] == 3 and d["gpu_launches"] > 0 and "workload" in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] > 0
    assert d["roofline"]["kernel"] == "k_engine_primal" and d["roofline"]["traffic_evidence"]["file"].startswith("profiles/")
    assert 0 < d["roofline"]["frac"] < 1 and d["roofline"]["unit"] == "GB/s"
    assert d["bnb"]["metric"] == "bnb_nodes_per_sec" and d["bnb"]["steps"] == 2 and d["bnb"]["value"] > 0
    assert run(["--no-bnb"])["bnb"] is None
    # a failure inside the block is reported there; the headline line survives
    monkeypatch.setattr(G.bnb, "sharded_intopt", lambda *a, **k: (_ for _ in ()).throw(RuntimeError("boom")))
    d = run(["--bnb-workers", "1"])
    assert "boom" in d["bnb"]["error"] and d["value"] > 0


def test_bnb_leg_worker_failure_is_raised_not_hung(stubbed, monkeypatch):
    """a worker thread that dies aborts the group's barrier and the leg raises in the main thread"""
    def search(worker, comm, minimize, node_lim=None, **kw):
        if comm.lr == 1:
            raise RuntimeError("worker 1 died")
        comm.g.barrier.wait(timeout=20)          # what the real driver does between slices
        return dict(total_nodes=node_lim, obj=0.0)

    monkeypatch.setattr(G.bnb, "sharded_intopt", search)
    with pytest.raises(RuntimeError, match="worker failed"):
        stubbed.run_bnb(_args(steps=1, warmup=0, bnb_workers=3), stubbed.WORKLOADS["mkp"], 0, 0, 1, embedded=True)
