#!/usr/bin/env python
"""bench.py -- headline measurement of the simplex hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c2s|c3s]
    python bench.py --impl reference ...      # the reference's CPU path (oracle port)

A *step* is one complete solve of the workload LP from the standard (all-slack)
basis to optimality through the C ABI.  Metric = simplex iterations per second
(BASELINE.json); time-to-optimal is ms_per_step.

  value  problem resident in HBM when the timed region starts (only the basis
         is reset between steps); timed on the device (CUDA events on the solve
         stream, max over ranks).
  e2e    the same metric through the public API with HOST buffers: handle
         creation (pinned host -> device copy of the problem) + solve + solution
         read-back, every step.
  roofline  the kernel with the largest share of device time in an extra
         profiled step (CUDA events around every launch on the solve stream).
  cpu_baseline  the oracle (C++ port of the reference, 1 thread) on a bounded
         sample (iteration limit) of the same LP, rank 0, N=1 only.

Single LPs do not shard ("replicas only"): with --gpus N every rank
solves its own replica and value is the aggregate.  Workloads: c2 = packing LP
2048x4096, 20 % dense, primal simplex + projected steepest edge + Harris
(BASELINE.json configs[1]); c3 = covering LP 16384x32768, ~16 nnz/col, dual
simplex (configs[2]); c2s/c3s are 1/4-size versions for quick checks.
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

WORKLOADS = {
    "c2": dict(gen="packing", kw=dict(m=2048, n=4096, density=0.20, seed=20240501), meth="primal",
               cpu_it_lim=600, cpu_mid=3000, cpu_mid_lim=150, name="packing LP m=2048 n=4096 20% dense, primal simplex, PSE pricing, Harris ratio test"),
    "c3": dict(gen="covering", kw=dict(m=16384, n=32768, kmin=8, kspan=17, seed=20240601), meth="dual",
               cpu_it_lim=1500, name="covering LP m=16384 n=32768 ~16 nnz/col, dual simplex, PSE pricing, Harris ratio test"),
    "c2s": dict(gen="packing", kw=dict(m=512, n=1024, density=0.20, seed=20240501), meth="primal",
                cpu_it_lim=600, cpu_mid=400, cpu_mid_lim=300, name="packing LP m=512 n=1024 20% dense (quarter-size check)"),
    "mkp": dict(gen="mkp", kw=dict(m=30, n=500, seed=20240701), meth="bnb", node_lim=400,
                name="multi-dimensional knapsack MIP m=30 n=500 (BASELINE.json configs[4]), branch-and-bound "
                     "(DTH branching, best-local-bound backtracking), nodes sharded across the ranks"),
    "c3s": dict(gen="covering", kw=dict(m=4096, n=8192, kmin=8, kspan=17, seed=20240601), meth="dual",
                cpu_it_lim=1500, name="covering LP m=4096 n=8192 (quarter-size check)"),
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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
                 "--format=csv,noheader,nounits", "-lms", "200"],
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


def oracle_problem(d):
    import oracle_lib as O
    from helpers import to_oracle
    return O, O.Problem.from_arrays(to_oracle(d))


def meth_code(nat, w):
    return nat.GLP_PRIMAL if w["meth"] == "primal" else nat.GLP_DUAL


class CpuSampler:
    """The reference's CPU path (the oracle: C++ port, one thread like the
    reference) on a BOUNDED sample of the workload: window A = the first
    cpu_it_lim iterations from the standard basis; window B (when the workload
    defines cpu_mid) = cpu_mid_lim iterations warm-started from the basis the
    oracle itself reaches after cpu_mid iterations (set-up run, not timed).
    The iteration rate falls as the basis fills with structural columns, so a
    start-only window would flatter the CPU."""

    def __init__(self, d, w):
        import oracle_lib as O
        from helpers import to_oracle
        self.O, self.d, self.w, self.od = O, d, w, to_oracle(d)
        self.meth = O.GLP_PRIMAL if w["meth"] == "primal" else O.GLP_DUAL
        self.mid_stat = None
        self.setup_s = 0.0
        if w.get("cpu_mid"):
            t0 = time.perf_counter()
            P = O.Problem.from_arrays(self.od)
            P.simplex(meth=self.meth, it_lim=w["cpu_mid"])
            self.mid_stat = P.solution()["stat"].copy()
            self.setup_s = time.perf_counter() - t0

    def step(self):
        """one bounded sample; returns (iterations, seconds)"""
        O, w = self.O, self.w
        P = O.Problem.from_arrays(self.od)
        t0 = time.perf_counter()
        P.simplex(meth=self.meth, it_lim=w["cpu_it_lim"])
        dt = time.perf_counter() - t0
        it = P.solution()["it_cnt"]
        if self.mid_stat is not None:
            Q = O.Problem.from_arrays(self.od)
            Q.set_stat(self.mid_stat)
            t0 = time.perf_counter()
            Q.simplex(meth=self.meth, it_lim=w["cpu_mid_lim"])
            dt += time.perf_counter() - t0
            it += Q.solution()["it_cnt"]
        return it, dt

    def describe(self):
        w = self.w
        s = "first %d iterations of the same LP from the standard basis" % w["cpu_it_lim"]
        if self.mid_stat is not None:
            s += " + %d iterations warm-started from the oracle's own basis after %d iterations" % (
                w["cpu_mid_lim"], w["cpu_mid"])
        return s + " (C++ port of the reference, single thread; the JS reference cannot run here: no JS engine)"


# DRAM traffic of the dominant kernels from the committed `ncu --set full` captures (never
# measured inside a bench run): whole-launch sums, one launch = several hundred iterations
TRAFFIC_EVIDENCE = {
    "primal": {"file": "profiles/r01j_ncu_full_engine_primal.csv", "kernel": "k_engine_primal",
               "dram_read_bytes_per_launch": 41.924864e6, "dram_write_bytes_per_launch": 2.069248e6,
               "launch_ms_under_ncu": 16.451,
               "reading": "C2 is L2-resident: ~0.1 MB of DRAM traffic per iteration against 45.8 MB of "
                          "algorithmic bytes -- the iteration is latency-bound, not bandwidth-bound"},
    "dual": {"file": "profiles/r01h_ncu_full_prof44_engine_dual.csv", "kernel": "k_engine_dual",
             "dram_read_bytes_per_launch": 285.441696e9, "dram_write_bytes_per_launch": 13.03564e9,
             "launch_ms_under_ncu": 158.701,
             "reading": "C3, k ~ 5000, ~1000 iterations in the launch: DRAM read = 8 k^2 bytes x iterations, "
                        "i.e. traffic = algorithmic bytes of the dense T*v stream, no re-reads"},
}


def roofline_from_profile(prof, eng_name, peak, peak_src, evidence=None):
    """roofline object from a glpb_profile_report: per-kernel CUDA-event time and algorithmic
    bytes; the phases of the persistent engine (eng_*) are summed into one unit `eng_name`."""
    ref_split = {k: v for k, v in prof.items() if k.startswith("ref_")}          # inside k_refactor: informational
    units = {k: v for k, v in prof.items() if not k.startswith("k_engine_") and not k.startswith("ref_")}
    tot_prof_ms = sum(v["ms"] for v in units.values()) or 1.0
    # the dominant kernel is the persistent engine: its roofline entry is the sum of the
    # algorithmic bytes of all its phases over the CUDA-event time of its launches
    phases = {k: v for k, v in units.items() if k.startswith("eng_")}
    agg = {"ms": sum(v["ms"] for v in phases.values()), "bytes": sum(v["bytes"] for v in phases.values()),
           "count": max([v["count"] for v in phases.values()] or [0])}
    cands = {k: v for k, v in units.items() if not k.startswith("eng_") and v["bytes"] > 0 and v["count"] > 0}
    if agg["ms"] > 0 and agg["count"] > 0:
        cands[eng_name] = agg
    top = max(cands, key=lambda k: cands[k]["ms"]) if cands else None
    if not top:
        return None
    v = cands[top]
    ach = (v["bytes"] / v["count"]) / (v["ms"] / v["count"] * 1e-3) / 1e9
    shares = {k: x["ms"] for k, x in units.items() if not k.startswith("eng_")}
    shares[eng_name] = agg["ms"]
    return {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak, "unit": "GB/s",
            "frac": ach / peak, "traffic": None, "traffic_evidence": evidence, "peak_source": peak_src,
            "bytes_per_launch": v["bytes"] / v["count"], "us_per_launch": 1000.0 * v["ms"] / v["count"],
            "share_of_device_time": v["ms"] / tot_prof_ms,
            "note": "for the persistent engine one 'launch' = one simplex iteration (all phases); "
                    "phase_table splits it (SM-cycle stamps of CTA 0 between grid barriers)",
            "kernel_shares": {k: round(x / tot_prof_ms, 4) for k, x in sorted(shares.items(), key=lambda kv: -kv[1])[:8]},
            "phase_table": {k[4:]: {"us": round(1000.0 * x["ms"] / max(1, x["count"]), 3),
                                    "GBps": round(x["bytes"] / max(1e-9, x["ms"]) / 1e6, 1)}
                            for k, x in sorted(phases.items())},
            "refactor_split_ms": {k[4:]: round(x["ms"], 2) for k, x in ref_split.items()}}


def run_reference(args, w, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.
    The reference is JavaScript and no JS engine exists in this image, so this
    arm times the oracle (C++ port, single thread like the reference)."""
    if rank != 0:
        return
    import glpk_js_b200 as G
    nat = G.native
    d = nat.generate(w["gen"], **w["kw"])
    smp = CpuSampler(d, w)
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
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], **{k: v for k, v in w["kw"].items()}},
            "cpu_baseline": {"value": val, "unit": "iter/s", "cores": 1, "kind": "port", "sample": smp.describe()},
            "e2e": {"value": val, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_bnb(args, w, rank, local_rank, world, embedded=False):
    """--workload mkp: branch-and-bound throughput (nodes/s).  A step = node_lim
    node LPs per rank (ios_solve_node: warm-started dual simplex + Driebeck-Tomlin
    branching); nodes are sharded across the ranks (glpk.js_b200/bnb.py: incumbent
    all-reduce + node migration over NCCL).  Weak scaling: the per-rank node budget
    is fixed.  embedded: called from the default run (process group and device already
    set up); returns the line (rank 0) instead of printing it."""
    import torch
    import torch.distributed as dist
    import glpk_js_b200 as G
    from glpk_js_b200 import bnb
    nat = G.native
    if not embedded:
        if world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        torch.cuda.set_device(local_rank)
    n_steps = min(args.steps, 3) if embedded else args.steps
    n_warm = min(args.warmup, 3) if embedded else args.warmup
    d = nat.generate(w["gen"], **w["kw"])
    node_lim = w["node_lim"]
    # worker threads (device handles) per GPU; every node costs host work (set-up, tree), so the
    # useful number is bounded by the host cores this rank can count on
    W = args.bnb_workers if args.bnb_workers > 0 else max(1, min(8, ((os.cpu_count() or 8) - 2) // world))
    outer = bnb.TorchComm() if world > 1 else None
    import threading

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    times, nodes, objs = [], [], []
    acct = {"launches": 0, "graph_launches": 0, "syncs": 0, "iterations": 0, "refactorizations": 0}
    for s in range(n_warm + n_steps):
        probs = []
        for _ in range(W):                              # host buffers -> device every step (e2e == value here)
            P = nat.Problem(d, device=local_rank)
            assert P.simplex(meth=nat.GLP_PRIMAL) == 0  # root LP, as solve_mip requires (lib/glpapi09.js:67-72)
            probs.append(P)
        if s == n_warm:
            sampler.start()
        group = bnb.LocalGroup(W)
        results = [None] * W

        errors = []

        def work(r):
            try:
                torch.cuda.set_device(local_rank)
                comm = bnb.HybridComm(group, r, outer)
                results[r] = bnb.sharded_intopt(bnb.Worker(probs[r]), comm, minimize=(d["dir"] == nat.GLP_MIN),
                                                node_lim=node_lim, msg_lev=0)
            except BaseException as e:       # a dead worker must not leave the others waiting at the barrier
                errors.append(e)
                try:
                    group.barrier.abort()
                except Exception:
                    pass

        barrier()
        c0 = [P.counters() for P in probs]
        t0 = time.perf_counter()
        threads = [threading.Thread(target=work, args=(r,)) for r in range(1, W)]
        for t in threads:
            t.start()
        work(0)
        for t in threads:
            t.join()
        if errors:
            for P in probs:
                P.close()
            raise RuntimeError("branch-and-bound worker failed: %r" % (errors[0],))
        barrier()
        dt = time.perf_counter() - t0
        c1 = [P.counters() for P in probs]
        for P in probs:
            P.close()
        if s >= n_warm:
            for key in acct:
                acct[key] += sum(b[key] - a[key] for a, b in zip(c0, c1))
            times.append(dt)
            nodes.append(results[0]["total_nodes"])
            objs.append(results[0]["obj"])
    clocks = sampler.stop()
    my = float(sum(times))
    if world > 1:
        t = torch.tensor([my], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        my = float(t.item())
    value = sum(nodes) / my
    # per-kernel device time of one worker's search (CUDA events on its stream; separate,
    # untimed pass): which kernel the node time goes to and its algorithmic bytes
    roofline, node_acct = None, None
    if rank == 0:
        Pp = nat.Problem(d, device=local_rank)
        assert Pp.simplex(meth=nat.GLP_PRIMAL) == 0
        Pp.set_profile(1)
        rp = bnb.sharded_intopt(bnb.Worker(Pp), bnb.HybridComm(bnb.LocalGroup(1), 0, None),
                                minimize=(d["dir"] == nat.GLP_MIN), node_lim=node_lim, msg_lev=0)
        prof = Pp.profile()
        Pp.close()
        peak, peak_src = peaks()
        roofline = roofline_from_profile(prof, "k_engine_dual", peak, peak_src)
        if roofline:
            roofline["note"] = ("one worker handle, %d nodes, profiling pass outside the timed region; node LPs "
                                "(m=30) run in a ONE-CTA engine out of shared memory/L2, so the HBM fraction is "
                                "low by construction: node time is launch + synchronisation latency"
                                % rp["total_nodes"])
        local_nodes = max(1, sum(nodes) // world)      # acct covers this rank's handles only
        node_acct = {k + "_per_node": round(v / local_nodes, 2) for k, v in acct.items()}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle_lib as O
        from helpers import to_oracle
        Q = O.Problem.from_arrays(to_oracle(d))
        Q.simplex(meth=O.GLP_PRIMAL)
        t0 = time.perf_counter()
        Q.intopt(node_lim=node_lim)
        dtc = time.perf_counter() - t0
        cpu = {"value": Q.mip()["nodes"] / dtc, "unit": "nodes/s", "cores": 1, "kind": "port",
               "sample": "%d nodes of the same search (node limit), C++ port of the reference, single thread" % Q.mip()["nodes"]}
    if rank == 0:
        line = {"metric": "bnb_nodes_per_sec", "value": value, "unit": "nodes/s", "n_gpus": world, "steps": n_steps,
                "warmup": n_warm, "ms_per_step": 1000.0 * my / n_steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": w["name"], **w["kw"], "node_lim_per_worker": node_lim,
                           "workers_per_gpu": W,
                           "parallelism": "nodes sharded over %d GPU(s) x %d worker handles" % (world, W)},
                "clocks": clocks, "e2e": {"value": value, "unit": "nodes/s", "h2d_bytes_per_step": int(sum(
                    a.nbytes for a in d.values() if isinstance(a, np.ndarray))), "d2h_bytes_per_step": 8 * (d["m"] + d["n"])},
                "gpu_launches": int(acct["launches"]) * world, "roofline": roofline, "cpu_baseline": cpu,
                "per_node": node_acct, "incumbent": objs[-1] if objs else None}
        if embedded:
            return line
        print(json.dumps(line), flush=True)
    if world > 1 and not embedded:
        dist.destroy_process_group()
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-full", action="store_true", help="profile the whole solve, not its first 1500 iterations")
    ap.add_argument("--no-c3", action="store_true", help="skip the extra full solve of the 16384x32768 LP")
    ap.add_argument("--no-bnb", action="store_true", help="skip the branch-and-bound block (nodes/s, C5) of the default line")
    ap.add_argument("--c3-mid", type=int, default=60000, help="iteration at which the C3 solve is split for the CPU sample")
    ap.add_argument("--c3-cpu-mid-lim", type=int, default=100)
    ap.add_argument("--bnb-workers", type=int, default=0,
                    help="mkp workload: B&B worker handles (threads) per GPU; 0 = min(8, (host cores - 2) / ranks)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return
    if w["meth"] == "bnb":
        run_bnb(args, w, rank, local_rank, world)
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
    meth = meth_code(nat, w)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush_l2():
        flush_buf.fill_(1)
        torch.cuda.synchronize()

    # ---- resident leg: value ----
    P = nat.Problem(d, device=local_rank)
    sampler = ClockSampler(local_rank)
    dev_ms, iters, launches0 = [], [], 0
    for s in range(args.warmup + args.steps):
        if s == args.warmup:
            barrier()
            sampler.start()
            launches0 = P.counters()["launches"]
            t_wall0 = time.perf_counter()
        flush_l2()
        P.std_basis()
        it_before = P.solution()["it_cnt"]      # it_cnt accumulates per handle, like glp_prob.it_cnt
        rc = P.simplex(meth=meth)
        c = P.counters()
        if s >= args.warmup:
            dev_ms.append(c["solve_us"] / 1000.0)
            iters.append(P.solution()["it_cnt"] - it_before)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    cnt = P.counters()
    launches = cnt["launches"] - launches0
    sol = P.solution()
    status, obj = sol["status"], sol["obj"]

    my_ms = float(sum(dev_ms))
    tot_it = float(sum(iters))
    if world > 1:
        t = torch.tensor([my_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        it = torch.tensor([tot_it], device="cuda", dtype=torch.float64)
        dist.all_reduce(it, op=dist.ReduceOp.SUM)
        max_ms, all_it = float(t.item()), float(it.item())
    else:
        max_ms, all_it = my_ms, tot_it
    value = all_it / (max_ms / 1000.0)

    # ---- profiled extra step: per-kernel device time (CUDA events on the solve stream) and,
    #      inside the persistent engine, per-phase time (SM cycle stamps between grid barriers
    #      scaled to the CUDA-event time of the engine launches) ----
    P.std_basis()
    P.set_profile(1)
    P.simplex(meth=meth, it_lim=int(iters[-1]) if (args.profile_full or iters[-1] <= 20000) else 1500)
    prof = P.profile()
    P.set_profile(0)
    peak, peak_src = peaks()
    roofline = roofline_from_profile(prof, "k_engine_primal" if w["meth"] == "primal" else "k_engine_dual",
                                     peak, peak_src, TRAFFIC_EVIDENCE.get(w["meth"]))
    P.close()

    # ---- e2e leg: host buffers every step ----
    pinned = {}
    for k_, a in d.items():
        if isinstance(a, np.ndarray):
            t_ = torch.from_numpy(a.copy()).pin_memory()
            pinned[k_] = t_.numpy()
        else:
            pinned[k_] = a
    h2d = int(sum(a.nbytes for a in pinned.values() if isinstance(a, np.ndarray)))
    d2h = (m + n) * (4 + 8 + 8) + m * 4 + 40
    e2e_t, e2e_it = [], []
    for s in range(args.warmup + args.steps):
        if s == args.warmup:
            barrier()
        flush_l2()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        Q = nat.Problem(pinned, device=local_rank)
        Q.simplex(meth=meth)
        so = Q.solution()
        dt = time.perf_counter() - t0
        Q.close()
        if s >= args.warmup:
            e2e_t.append(dt)
            e2e_it.append(so["it_cnt"])
    barrier()
    my_e2e = float(sum(e2e_t))
    if world > 1:
        t = torch.tensor([my_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        it = torch.tensor([float(sum(e2e_it))], device="cuda", dtype=torch.float64)
        dist.all_reduce(it, op=dist.ReduceOp.SUM)
        e2e_val = float(it.item()) / float(t.item())
    else:
        e2e_val = sum(e2e_it) / my_e2e

    # ---- CPU baseline (rank 0, N=1 only): the oracle on a bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        smp = CpuSampler(d, w)
        itc, dt = smp.step()
        cpu = {"value": itc / dt, "unit": "iter/s", "cores": 1, "kind": "port",
               "sample": smp.describe() + "; %d iterations in %.1f s" % (itc, dt)}

    # ---- the north-star shape (BASELINE.json configs[2]): one full dual solve of the
    #      16384 x 32768 covering LP, device-resident, next to a bounded CPU sample ----
    c3 = None
    if rank == 0 and world == 1 and args.workload == "c2" and not args.no_c3:
        w3 = WORKLOADS["c3"]
        d3 = nat.generate(w3["gen"], **w3["kw"])
        P3 = nat.Problem(d3, device=local_rank)
        P3.set_profile(1)
        flush_l2()
        # two calls: the basis at the midpoint seeds the CPU sample below
        P3.simplex(meth=nat.GLP_DUAL, it_lim=args.c3_mid)
        us_a = P3.counters()["solve_us"]
        mid_stat3 = P3.solution()["stat"].copy()
        rc3 = P3.simplex(meth=nat.GLP_DUAL)
        c3cnt, s3 = P3.counters(), P3.solution()
        c3cnt["solve_us"] += us_a
        prof3 = {k: v for k, v in P3.profile().items() if k.startswith("eng_") and v["count"] > 0 and v["bytes"] > 0}
        eng3_ms = sum(v["ms"] for v in prof3.values())
        eng3_bytes = sum(v["bytes"] for v in prof3.values())
        P3.close()
        top3 = max(prof3, key=lambda k: prof3[k]["ms"]) if prof3 else None
        c3 = {"workload": w3["name"], "value": s3["it_cnt"] / (c3cnt["solve_us"] * 1e-6), "unit": "iter/s",
              "time_to_optimal_ms": c3cnt["solve_us"] / 1000.0, "iterations": int(s3["it_cnt"]), "status": int(s3["status"]),
              "rc": int(rc3), "objective": s3["obj"], "refactorizations": c3cnt["refactorizations"],
              "kernel_size_k": c3cnt["k"], "note": "single solve, profiling marks on (a few percent slower)"}
        if top3:
            v3 = prof3[top3]
            ach3 = v3["bytes"] / (v3["ms"] * 1e-3) / 1e9
            c3["roofline"] = {"bound": "hbm", "kernel": "k_engine_dual:" + top3[4:], "achieved": ach3, "peak": peak,
                              "unit": "GB/s", "frac": ach3 / peak, "traffic": None,
                              "traffic_evidence": TRAFFIC_EVIDENCE["dual"],
                              "us_per_iteration": 1000.0 * v3["ms"] / v3["count"],
                              "bytes_per_iteration": v3["bytes"] / v3["count"],
                              "whole_engine": {"achieved": eng3_bytes / max(1e-9, eng3_ms) / 1e6, "unit": "GB/s",
                                               "frac": eng3_bytes / max(1e-9, eng3_ms) / 1e6 / peak,
                                               "us_per_iteration": 1000.0 * eng3_ms / v3["count"]}}
        if not args.no_cpu_baseline:
            smp3 = CpuSampler(d3, w3)
            it3, dt3 = smp3.step()
            # second window: the same port warm-started from the DEVICE's basis after c3_mid
            # iterations (reaching it on the CPU would take the better part of an hour)
            O3 = smp3.O
            Q3 = O3.Problem.from_arrays(smp3.od)
            Q3.set_stat(mid_stat3)
            t0 = time.perf_counter()
            Q3.simplex(meth=O3.GLP_DUAL, it_lim=args.c3_cpu_mid_lim)
            dtm = time.perf_counter() - t0
            itm = Q3.solution()["it_cnt"]
            c3["cpu_baseline"] = {"value": (it3 + itm) / (dt3 + dtm), "unit": "iter/s", "cores": 1, "kind": "port",
                                  "start_window": {"iterations": int(it3), "seconds": dt3, "iter_per_s": it3 / dt3},
                                  "mid_window": {"iterations": int(itm), "seconds": dtm, "iter_per_s": itm / max(dtm, 1e-9),
                                                 "from": "the device's basis after %d iterations" % args.c3_mid},
                                  "sample": smp3.describe() + " + %d iterations warm-started from the device's basis "
                                  "after %d iterations (one refactorisation period, factorisation included)"
                                  % (itm, args.c3_mid)}

    # ---- the multi-GPU part of BASELINE.json's metric: branch-and-bound nodes/s on C5 (configs[4]),
    #      nodes sharded across all ranks of this run; same code as --workload mkp, fewer steps ----
    bnb_block = None
    if args.workload == "c2" and not args.no_bnb:
        try:
            bnb_block = run_bnb(args, WORKLOADS["mkp"], rank, local_rank, world, embedded=True)
        except Exception as e:      # the headline line must not depend on the extra block
            bnb_block = {"error": repr(e)}

    if rank == 0:
        line = {"metric": "simplex_iterations_per_sec", "value": value, "unit": "iter/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": max_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": w["name"], **w["kw"], "nnz": nnz, "step": "one full solve from the standard basis to optimality",
                           "parallelism": "replicas only (a single LP does not shard)" if world > 1 else "1 GPU",
                           "l2": "flushed between steps (256 MiB write); within a solve the working set is L2-resident by design"},
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "iter/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": 1000.0 * my_e2e / args.steps},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "iterations_per_step": tot_it / args.steps, "time_to_optimal_ms": max_ms / args.steps,
                "status": int(status), "objective": obj, "wall_s_timed_region": t_wall,
                "refactorizations": cnt["refactorizations"], "kernel_size_k": cnt["k"],
                "north_star_c3": c3, "bnb": bnb_block}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
