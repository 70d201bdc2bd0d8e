#!/usr/bin/env python
"""bench.py -- headline measurement of the simplex hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c3s|c2s|mkp]
    python bench.py --impl reference ...      # the reference's CPU path (oracle port)

Headline (BASELINE.json north_star / configs[2]): the covering LP 16384 x 32768,
~16 non-zeros per column, dual simplex with projected steepest edge and the
Harris ratio test.  A *step* is one complete solve from the standard basis to
optimality through the C ABI with HOST buffers: handle creation (pinned host ->
device copy of the problem), solve, solution read-back, handle destruction.

  e2e    iterations / wall time of the whole step (copies inside the timed region)
  value  iterations / device time of the solve alone (problem resident in HBM when
         the clock starts): CUDA events on the solve stream, taken inside the same
         K steps; max over ranks
  roofline  the persistent engine in an extra profiled solve: algorithmic bytes
         (SURVEY 8d formulas, accumulated on the device per phase) over the CUDA-
         event time of its launches; `traffic` = DRAM bytes per iteration from the
         committed ncu capture (profiles/traffic.json)
  cpu_baseline  the oracle (C++ port of the reference, 1 thread) on a bounded
         sample of the same LP (rank 0, N=1): the first iterations from the
         standard basis + a window warm-started from the ORACLE'S OWN basis after
         60000 iterations (tests/golden/c3_mid_basis.npz)
  parity  status / objective against the independent HiGHS and oracle pins
         (tests/golden/lp_pins.json), KKT residuals of the returned solution
  c2     nested: one solve of BASELINE.json configs[1] (packing LP 2048 x 4096)
  bnb    LAST: branch-and-bound nodes/s (configs[4], knapsack 30 x 500) with the
         nodes sharded over ALL ranks of this run, and a knapsack of the same
         family solved to completion whose optimum is checked against its pins

Single LPs do not shard ("replicas only"): with --gpus N every rank
solves its own replica of the LP and `value` is the aggregate; the part of the
metric that scales across GPUs is the `bnb` block.  Per-phase tables go to
gpurun_out/bench_detail.json (the JSON line stays short).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")

WORKLOADS = {
    "c3": dict(gen="covering", kw=dict(m=16384, n=32768, kmin=8, kspan=17, seed=20240601), meth="dual", pin="c3",
               cpu_it_lim=30, cpu_mid_lim=30, mid_basis="c3_mid_basis.npz",
               name="covering LP m=16384 n=32768 ~16 nnz/col, dual simplex, PSE pricing, Harris ratio test"),
    "c2": dict(gen="packing", kw=dict(m=2048, n=4096, density=0.20, seed=20240501), meth="primal", pin="c2",
               cpu_it_lim=600, cpu_mid=3000, cpu_mid_lim=150,
               name="packing LP m=2048 n=4096 20% dense, primal simplex, PSE pricing, Harris ratio test"),
    "c2s": dict(gen="packing", kw=dict(m=512, n=1024, density=0.20, seed=20240501), meth="primal",
                cpu_it_lim=600, cpu_mid=400, cpu_mid_lim=300, name="packing LP m=512 n=1024 20% dense (quarter-size check)"),
    "c3s": dict(gen="covering", kw=dict(m=4096, n=8192, kmin=8, kspan=17, seed=20240601), meth="dual",
                cpu_it_lim=1500, name="covering LP m=4096 n=8192 (quarter-size check)"),
    "mkp": dict(gen="mkp", kw=dict(m=30, n=500, seed=20240701), meth="bnb", node_lim=150000,
                name="multi-dimensional knapsack MIP m=30 n=500, branch-and-bound (DTH branching, best-local-bound "
                     "backtracking), batched node LPs, nodes sharded across the ranks"),
}
MKP_CHECK = dict(m=30, n=60, seed=20240701, pin="mkp_30x60_seed20240701")


def config_of(w):
    """the `config` object: identical in both arms"""
    return {"workload": w["name"], **w["kw"]}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_pins():
    try:
        with open(os.path.join(GOLDEN, "lp_pins.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi SM clocks and throttle reasons while the timed region runs"""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "500"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": reasons}


def oracle_run_record(w):
    """what the oracle's own uninterrupted solve of this workload in the build container has logged
    (tests/golden/<pin>_oracle_run.json, written by tests/golden/make_c3_pins.py while it runs)"""
    try:
        with open(os.path.join(GOLDEN, w.get("pin", "-") + "_oracle_run.json")) as f:
            r = json.load(f)
        last = r["log"][-1]
        out = {"complete": not r.get("partial", True), "iterations": int(r.get("it_cnt", last["it"])),
               "seconds": round(float(r.get("seconds", last["seconds"])), 1), "cores": 1,
               "where": "build container, tests/golden/make_c3_pins.py"}
        out["iter_per_s"] = out["iterations"] / max(1e-9, out["seconds"])
        if not out["complete"]:
            out["note"] = "still running when committed"
        return out
    except Exception:
        return None


def to_oracle(d):
    m = d["m"]
    return dict(m=m, n=d["n"], dir=d["dir"], c0=d["c0"], r_type=d["type"][:m], r_lb=d["lb"][:m],
                r_ub=d["ub"][:m], c_type=d["type"][m:], c_lb=d["lb"][m:], c_ub=d["ub"][m:],
                c_coef=d["coef"], c_kind=d["kind"], A_ptr=d["A_ptr"], A_ind=d["A_ind"], A_val=d["A_val"])


class CpuSampler:
    """The reference's CPU path (the oracle: C++ port, one thread like the
    reference) on a BOUNDED sample of the workload, nothing of the product
    library involved: window A = the first cpu_it_lim iterations from the
    standard basis; window B = cpu_mid_lim iterations warm-started from a basis
    deep inside the solve -- the oracle's own basis after 60000 of the ~120000
    iterations (committed fixture, tests/golden/make_c3_pins.py) or, for the
    small workloads, the basis the oracle reaches in an untimed set-up run.
    The time per iteration grows by three orders of magnitude as the basis
    fills with structural columns (C3: 1200 it/s at the start, ~3 it/s at the
    midpoint; tests/golden/c3_oracle_run.json), so for C3 both windows have the
    SAME number of iterations: iterations / seconds over both is then the
    inverse of the mean time per iteration at the two points, an estimate of
    the full-solve rate; a start-only window would flatter the CPU a
    hundredfold.  The oracle's own uninterrupted solve in the build container
    is reported next to it as `full_solve` when it exists."""

    def __init__(self, w):
        import oracle_lib as O
        self.O, self.w = O, w
        self.d = O.generate(w["gen"], **w["kw"])
        self.od = to_oracle(self.d)
        self.meth = O.GLP_PRIMAL if w["meth"] == "primal" else O.GLP_DUAL
        self.mid_stat, self.mid_from = None, None
        path = os.path.join(GOLDEN, w.get("mid_basis", "-"))
        if os.path.exists(path):
            z = np.load(path)
            self.mid_stat = z["stat"].astype(np.int32)
            self.mid_from = "the oracle's own basis after %d iterations (fixture)" % int(z["it"])
        elif w.get("cpu_mid"):
            P = O.Problem.from_arrays(self.od)
            P.simplex(meth=self.meth, it_lim=w["cpu_mid"])
            self.mid_stat = P.solution()["stat"].copy()
            self.mid_from = "the oracle's own basis after %d iterations (untimed set-up run)" % w["cpu_mid"]

    def step(self):
        """one bounded sample; returns (iterations, seconds)"""
        O, w = self.O, self.w
        P = O.Problem.from_arrays(self.od)
        t0 = time.perf_counter()
        P.simplex(meth=self.meth, it_lim=w["cpu_it_lim"])
        dt = time.perf_counter() - t0
        it = P.solution()["it_cnt"]
        self.windows = {"start": {"iterations": int(it), "seconds": round(dt, 3)}}
        if self.mid_stat is not None:
            # The warm-started call factorises the basis and recomputes bbar / cbar before its first iteration --
            # what the reference's loop does once per refactorisation period (bfcp.nfs_max = 100 basis updates),
            # not once per 30 iterations.  The window is therefore charged its iterations plus the share
            # itm / 100 of that set-up, i.e. the cost per iteration of the steady loop at this point of the solve.
            Q = O.Problem.from_arrays(self.od)
            Q.set_stat(self.mid_stat)
            ev_iter = O.EV_D_ITER if self.meth == O.GLP_DUAL else O.EV_P_ITER
            stamps = []

            def hook(ev, csa):
                if ev == ev_iter:
                    stamps.append(time.perf_counter())
            Q.set_hook(hook)
            t0 = time.perf_counter()
            Q.simplex(meth=self.meth, it_lim=w["cpu_mid_lim"])
            t1 = time.perf_counter()
            Q.set_hook(None)
            itm = Q.solution()["it_cnt"]
            setup = (stamps[0] - t0) if stamps else (t1 - t0)
            iterating = (t1 - stamps[0]) if stamps else 0.0
            dtm = iterating + setup * itm / 100.0
            self.windows["mid"] = {"iterations": int(itm), "seconds": round(dtm, 3), "seconds_iterating": round(iterating, 3),
                                   "seconds_factorise_bbar_cbar": round(setup, 3),
                                   "charged": "iterating + set-up x iterations / 100"}
            dt += dtm
            it += itm
        return it, dt

    def describe(self):
        w = self.w
        s = "first %d iterations from the standard basis" % w["cpu_it_lim"]
        if self.mid_stat is not None:
            s += " + %d iterations (+ amortised share of one refactorisation per 100 updates) from %s" % (w["cpu_mid_lim"], self.mid_from)
        return s + "; C++ port of the reference, 1 thread"


def run_reference(args, w, rank):
    """--impl reference: the reference's own CPU implementation of the path.  The
    reference is JavaScript and no JS engine exists on the box, so this arm times
    the oracle (C++ port, single thread like the reference).  It loads nothing of
    the product: the problem comes from the oracle's generator."""
    if rank != 0:
        return
    smp = CpuSampler(w)
    times, iters = [], []
    for s in range(args.warmup + args.steps):
        it, dt = smp.step()
        if s >= args.warmup:
            times.append(dt)
            iters.append(it)
    val = sum(iters) / sum(times)
    line = {"impl": "reference", "metric": "simplex_iterations_per_sec", "value": val, "unit": "iter/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_of(w),
            "cpu_baseline": {"value": val, "unit": "iter/s", "cores": 1, "kind": "port", "sample": smp.describe(),
                             "windows": smp.windows},
            "e2e": {"value": val, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    rec = oracle_run_record(w)
    if rec:
        line["cpu_baseline"]["oracle_run"] = rec
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# branch-and-bound block (nodes sharded over the ranks of the run)
# ----------------------------------------------------------------------------------------------
def bnb_block(args, rank, local_rank, world, n_steps, n_warm):
    """nodes/s of the batched branch-and-bound on the knapsack of BASELINE.json
    configs[4], per-rank node budget fixed (weak scaling), identical per-GPU
    configuration at every N; then a knapsack of the same family solved to
    completion, sharded over the same ranks, optimum checked against its pins."""
    import torch
    import torch.distributed as dist
    import glpk_js_b200 as G
    from glpk_js_b200 import bnb
    nat = G.native
    w = WORKLOADS["mkp"]
    d = nat.generate(w["gen"], **w["kw"])
    comm = bnb.TensorComm()
    node_lim = args.bnb_nodes if args.bnb_nodes > 0 else w["node_lim"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    times, nodes, launches, rounds = [], [], 0, 0
    for s in range(n_warm + n_steps):
        P = nat.Problem(d, device=local_rank)                 # host buffers -> device every step
        assert P.simplex(meth=nat.GLP_PRIMAL) == 0            # root LP, as solve_mip requires (lib/glpapi09.js:67-72)
        barrier()
        l0 = P.counters()["launches"]
        t0 = time.perf_counter()
        res = bnb.sharded_bnb_batched(bnb.BatchWorker(P), comm, minimize=False, batch=args.bnb_batch, node_lim=node_lim,
                                      max_ship=getattr(args, "bnb_ship", 64), msg_lev=0)
        barrier()
        dt = time.perf_counter() - t0
        if s >= n_warm:
            times.append(dt)
            nodes.append(res["total_nodes"])
            launches += P.counters()["launches"] - l0
            rounds += res["rounds"]
        P.close()
    my = float(sum(times))
    if world > 1:
        t = torch.tensor([my], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        my = float(t.item())
    value = sum(nodes) / my
    # ---- completion run: identical MIP optimum at this N ----
    dc = nat.generate("mkp", m=MKP_CHECK["m"], n=MKP_CHECK["n"], seed=MKP_CHECK["seed"])
    Pc = nat.Problem(dc, device=local_rank)
    assert Pc.simplex(meth=nat.GLP_PRIMAL) == 0
    barrier()
    t0 = time.perf_counter()
    rc = bnb.sharded_bnb_batched(bnb.BatchWorker(Pc), comm, minimize=False, batch=args.bnb_batch, max_ship=getattr(args, "bnb_ship", 64), msg_lev=0)
    barrier()
    tc = time.perf_counter() - t0
    x_ok = None
    if rc["holder"] == rank:
        mp = Pc.mip()
        x = mp["mipx"][dc["m"]:]
        act = np.zeros(dc["m"])
        cols = np.repeat(np.arange(dc["n"]), np.diff(dc["A_ptr"]))
        np.add.at(act, dc["A_ind"], dc["A_val"] * x[cols])
        x_ok = bool(np.all(x == np.round(x)) and np.all(act <= dc["ub"][:dc["m"]] + 1e-9)
                    and abs(float(dc["coef"] @ x) - mp["mip_obj"]) < 1e-9)
    Pc.close()
    if world > 1:
        flag = torch.tensor([1.0 if x_ok else (0.0 if x_ok is not None else -1.0)], device="cuda", dtype=torch.float64)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        x_ok = bool(flag.item() > 0)
    if rank != 0:
        return None
    pin = load_pins().get("mkp", {}).get(MKP_CHECK["pin"], {})
    expected = pin.get("highs_obj")
    blk = {"metric": "bnb_nodes_per_sec", "value": value, "unit": "nodes/s", "n_gpus": world, "steps": n_steps,
           "warmup": n_warm, "ms_per_step": 1000.0 * my / max(1, n_steps), "scaling": "weak",
           "node_lim_per_gpu": node_lim, "batch_per_gpu": args.bnb_batch or "16 x SM count", "host_threads_per_gpu": 1,
           "nodes_per_step": sum(nodes) / max(1, n_steps), "gpu_launches": int(launches) * world,
           "rounds_per_step": rounds / max(1, n_steps),
           "workload": "knapsack MIP m=30 n=500 (BASELINE.json configs[4]), batched node LPs sharded over the ranks",
           "optimum_check": {"instance": "knapsack m=%d n=%d seed=%d, solved to completion on %d GPU(s)" % (
               MKP_CHECK["m"], MKP_CHECK["n"], MKP_CHECK["seed"], world), "optimum": rc["obj"], "expected": expected,
               "expected_from": "HiGHS + oracle, tests/golden/lp_pins.json",
               "ret": rc["ret"], "open_left": rc["open_left"], "nodes": rc["total_nodes"], "seconds": round(tc, 3),
               "solution_feasible_integral": x_ok, "migrated_nodes": rc["moved_out"]},
           "optimum_ok": bool(expected is not None and rc["obj"] is not None and rc["ret"] == 0 and rc["open_left"] == 0
                              and abs(rc["obj"] - expected) <= 1e-6 and x_ok)}
    return blk


def bnb_cpu_baseline(node_lim=1500):
    """the oracle's serial branch-and-bound (ios_driver restated) on the same knapsack, 1 thread"""
    import oracle_lib as O
    w = WORKLOADS["mkp"]
    d = O.generate(w["gen"], **w["kw"])
    Q = O.Problem.from_arrays(to_oracle(d))
    Q.simplex(meth=O.GLP_PRIMAL)
    t0 = time.perf_counter()
    Q.intopt(node_lim=node_lim)
    dt = time.perf_counter() - t0
    return {"value": Q.mip()["nodes"] / dt, "unit": "nodes/s", "cores": 1, "kind": "port",
            "sample": "%d nodes of the same search, C++ port, 1 thread" % Q.mip()["nodes"]}


def run_bnb_line(args, rank, local_rank, world):
    """--workload mkp: the branch-and-bound block as its own JSON line"""
    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    sampler = ClockSampler(local_rank)
    sampler.start()
    blk = bnb_block(args, rank, local_rank, world, args.steps, args.warmup)
    clocks = sampler.stop()
    if rank == 0:
        w = WORKLOADS["mkp"]
        line = {"metric": "bnb_nodes_per_sec", "value": blk["value"], "unit": "nodes/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": blk["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_of(w),
                "clocks": clocks, "gpu_launches": blk["gpu_launches"],
                "e2e": {"value": blk["value"], "unit": "nodes/s",
                        "h2d_bytes_per_step": int(8 * 2 * 530 + 12 * 15000 + 8 * 500), "d2h_bytes_per_step": 64 * 600},
                "cpu_baseline": None if (world > 1 or args.no_cpu_baseline) else bnb_cpu_baseline(), "bnb": blk}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
def solve_steps(nat, torch, d, meth, local_rank, n_warm, n_steps, flush_l2, barrier, sampler=None):
    """K timed steps: host buffers -> handle -> solve -> read-back.  Returns per-step
    (wall seconds, device solve ms, iterations) lists, the last solution and counters."""
    pinned = {}
    for k_, a in d.items():
        pinned[k_] = torch.from_numpy(a.copy()).pin_memory().numpy() if isinstance(a, np.ndarray) else a
    wall, dev_ms, iters, launches = [], [], [], 0
    sol, cnt = None, None
    for s in range(n_warm + n_steps):
        if s == n_warm:
            barrier()
            if sampler:
                sampler.start()
        flush_l2()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        Q = nat.Problem(pinned, device=local_rank)
        rc = Q.simplex(meth=meth)
        so = Q.solution()
        dt = time.perf_counter() - t0
        c = Q.counters()
        Q.close()
        if s >= n_warm:
            wall.append(dt)
            dev_ms.append(c["solve_us"] / 1000.0)
            iters.append(so["it_cnt"])
            launches += c["launches"]
        sol, cnt = so, c
        sol["rc"] = rc
    barrier()
    h2d = int(sum(a.nbytes for a in pinned.values() if isinstance(a, np.ndarray)))
    return wall, dev_ms, iters, launches, sol, cnt, h2d


def parity_of(d, sol, pins):
    """status / objective vs the independent pins, KKT residuals (glp_check_kkt definitions)"""
    from helpers import kkt
    out = {"status_ok": bool(sol["status"] == 5 and sol["rc"] == 0)}
    ref = None
    for src in ("highs", "oracle"):
        if pins.get(src) and pins[src].get("obj") is not None:
            rel = abs(sol["obj"] - pins[src]["obj"]) / max(1.0, abs(pins[src]["obj"]))
            out["obj_rel_vs_" + src] = rel
            ref = rel if ref is None else max(ref, rel)
    out["obj_rel"] = ref
    r = kkt(d, sol)
    out["kkt_max"] = float(max(r.values()))
    out["ok"] = bool(out["status_ok"] and (ref is not None and ref <= 1e-9) and out["kkt_max"] <= 1e-9)
    return out


def roofline_of(prof, eng_name, peak, peak_src, traffic):
    """roofline object of the dominant unit of a profiled solve: the persistent engine
    (the sum of its phases) or a stand-alone kernel; per-phase table goes to the detail file"""
    units = {k: v for k, v in prof.items() if not k.startswith("k_engine_") and not k.startswith("ref_")}
    tot_ms = sum(v["ms"] for v in units.values()) or 1.0
    phases = {k: v for k, v in units.items() if k.startswith("eng_") and v["count"] > 0}
    eng = {"ms": sum(v["ms"] for v in phases.values()), "bytes": sum(v["bytes"] for v in phases.values()),
           "count": max([v["count"] for v in phases.values()] or [0])}
    if eng["count"] == 0 or eng["ms"] <= 0:
        return None, None
    top_phase = max(phases, key=lambda k: phases[k]["ms"])
    tp = phases[top_phase]
    ach = eng["bytes"] / (eng["ms"] * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": eng_name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": traffic.get("dram_bytes_per_iteration"), "traffic_source": "profiles/traffic.json (ncu --set full, per iteration)",
            "peak_source": peak_src, "launch_unit": "one simplex iteration of the persistent engine (all phases)",
            "bytes_per_launch": eng["bytes"] / eng["count"], "us_per_launch": 1000.0 * eng["ms"] / eng["count"],
            "share_of_device_time": eng["ms"] / tot_ms,
            "top_phase": {"name": top_phase[4:], "GBps": tp["bytes"] / max(1e-9, tp["ms"]) / 1e6,
                          "frac": tp["bytes"] / max(1e-9, tp["ms"]) / 1e6 / peak,
                          "us_per_iteration": 1000.0 * tp["ms"] / tp["count"],
                          "share_of_engine": tp["ms"] / eng["ms"]}}
    detail = {"phase_table": {k[4:]: {"us": round(1000.0 * x["ms"] / max(1, x["count"]), 3),
                                      "GBps": round(x["bytes"] / max(1e-9, x["ms"]) / 1e6, 1)} for k, x in sorted(phases.items())},
              "kernel_shares": {k: round(x["ms"] / tot_ms, 4) for k, x in sorted(
                  list({k: v for k, v in units.items() if not k.startswith("eng_")}.items()) + [(eng_name, eng)],
                  key=lambda kv: -kv[1]["ms"])[:10]},
              "refactor_split_ms": {k[4:]: round(x["ms"], 2) for k, x in prof.items() if k.startswith("ref_")}}
    return roof, detail


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="skip the extra profiled solve (roofline)")
    ap.add_argument("--no-c2", action="store_true", help="skip the nested solve of the 2048x4096 packing LP")
    ap.add_argument("--no-bnb", action="store_true", help="skip the branch-and-bound block")
    ap.add_argument("--bnb-nodes", type=int, default=0, help="node LPs per GPU and step of the bnb block (0 = 150000)")
    ap.add_argument("--bnb-batch", type=int, default=0, help="open nodes per launch and GPU (0 = 16 x SM count)")
    ap.add_argument("--bnb-ship", type=int, default=64, help="node records a donor rank ships per exchange")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if w["meth"] == "bnb":
            if rank == 0:
                cb = bnb_cpu_baseline(400 * max(1, args.steps))
                print(json.dumps({"impl": "reference", "metric": "bnb_nodes_per_sec", "value": cb["value"], "unit": "nodes/s",
                                  "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
                                  "config": config_of(w), "cpu_baseline": cb,
                                  "e2e": {"value": cb["value"], "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
            return
        run_reference(args, w, rank)
        return
    if w["meth"] == "bnb":
        run_bnb_line(args, rank, local_rank, world)
        return

    import torch
    import torch.distributed as dist
    import glpk_js_b200 as G
    nat = G.native
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    if nat.load().glpb_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")

    d = nat.generate(w["gen"], **w["kw"])
    m, n, nnz = d["m"], d["n"], len(d["A_val"])
    meth = nat.GLP_PRIMAL if w["meth"] == "primal" else nat.GLP_DUAL

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush_l2():
        flush_buf.fill_(1)
        torch.cuda.synchronize()

    # ---- the K timed steps: e2e (wall, copies inside) and value (device time of the solve) ----
    sampler = ClockSampler(local_rank)
    t_region = time.perf_counter()
    wall, dev_ms, iters, launches, sol, cnt, h2d = solve_steps(nat, torch, d, meth, local_rank, args.warmup, args.steps,
                                                                flush_l2, barrier, sampler)
    clocks = sampler.stop()
    t_region = time.perf_counter() - t_region
    d2h = (m + n) * (4 + 8 + 8) + m * 4 + 40

    def over_ranks(my_time, my_units):
        if world == 1:
            return my_time, my_units
        t = torch.tensor([my_time], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        u = torch.tensor([my_units], device="cuda", dtype=torch.float64)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        return float(t.item()), float(u.item())

    max_ms, all_it = over_ranks(float(sum(dev_ms)), float(sum(iters)))
    value = all_it / (max_ms / 1000.0)
    max_wall, _ = over_ranks(float(sum(wall)), float(sum(iters)))
    e2e_val = all_it / max_wall

    pins = load_pins().get(w.get("pin", "-"), {})
    parity = parity_of(d, sol, pins) if rank == 0 else None

    # ---- profiled extra solve: per-kernel device time (CUDA events on the solve stream) and, inside the
    #      persistent engine, per-phase time and algorithmic bytes ----
    roofline, detail = None, {}
    if rank == 0 and not args.no_profile:
        P = nat.Problem(d, device=local_rank)
        P.set_profile(1)
        P.simplex(meth=meth)
        prof = P.profile()
        P.close()
        peak, peak_src = peaks()
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(args.workload, {})
        except Exception:
            traffic = {}
        roofline, det = roofline_of(prof, "k_engine_primal" if w["meth"] == "primal" else "k_engine_dual", peak, peak_src, traffic)
        detail[args.workload] = det

    # ---- CPU baseline (rank 0, N=1 only): the oracle on a bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        smp = CpuSampler(w)
        itc, dt = smp.step()
        cpu = {"value": itc / dt, "unit": "iter/s", "cores": 1, "kind": "port", "sample": smp.describe(), "windows": smp.windows}
        rec = oracle_run_record(w)
        if rec:
            cpu["oracle_run"] = rec

    # ---- nested: BASELINE.json configs[1], one step (rank 0, N=1) ----
    c2 = None
    if rank == 0 and world == 1 and args.workload == "c3" and not args.no_c2:
        w2 = WORKLOADS["c2"]
        d2 = nat.generate(w2["gen"], **w2["kw"])
        wl2, dm2, it2, _, sol2, cnt2, _ = solve_steps(nat, torch, d2, nat.GLP_PRIMAL, local_rank, 1, 1, flush_l2, lambda: None)
        par2 = parity_of(d2, sol2, load_pins().get("c2", {}))
        c2 = {"workload": w2["name"], "value": it2[0] / (dm2[0] / 1000.0), "e2e": it2[0] / wl2[0], "unit": "iter/s",
              "iterations": int(it2[0]), "ms": dm2[0], "parity_ok": par2["ok"], "obj_rel": par2["obj_rel"], "kkt_max": par2["kkt_max"]}

    # ---- LAST: branch-and-bound, nodes sharded over all ranks of this run ----
    bnb = None
    if args.workload == "c3" and not args.no_bnb:
        try:
            bnb = bnb_block(args, rank, local_rank, world, min(args.steps, 3), min(args.warmup, 2))
            if rank == 0 and world == 1 and not args.no_cpu_baseline:
                bnb["cpu_baseline"] = bnb_cpu_baseline()
        except Exception as e:      # the headline line must not depend on the extra block
            bnb = {"error": repr(e)[:300], "optimum_ok": False}

    if rank == 0:
        line = {"metric": "simplex_iterations_per_sec", "value": value, "unit": "iter/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": max_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_of(w), "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "iter/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": 1000.0 * max_wall / args.steps},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
                "iterations_per_step": float(sum(iters)) / args.steps, "time_to_optimal_ms": max_ms / args.steps,
                "status": int(sol["status"]), "objective": sol["obj"], "nnz": nnz,
                "refactorizations": cnt["refactorizations"], "kernel_size_k": cnt["k"],
                "parallelism": "replicas only (a single LP does not shard)" if world > 1 else "1 GPU",
                "l2": "flushed between steps (256 MiB write)", "wall_s_timed_region": t_region,
                "c2": c2, "bnb": bnb}
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "bench_detail.json"), "w") as f:
                json.dump({"line": line, "detail": detail}, f, indent=1)
        except Exception:
            pass
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
