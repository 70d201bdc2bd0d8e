"""CPU tests of the multi-GPU branch-and-bound protocol (glpk.js_b200/bnb.py):
the N>1 host logic with fake workers, in-process and over torch.distributed
gloo with world_size 2.  No CUDA device is involved."""
import itertools
import os
import threading

import numpy as np
import pytest

import glpk_js_b200 as G

bnb = G.bnb


class KnapsackWorker:
    """Tiny 0-1 knapsack branch-and-bound with the Worker interface: a node is
    a vector of n bytes (0/1 fixed, 2 free)."""

    def __init__(self, w, p, cap):
        self.w, self.p, self.cap, self.n = list(w), list(p), cap, len(w)
        self.pool, self.best, self.has, self.solved = [], -float("inf"), False, 0

    def begin(self, **kw):
        self.pool = [np.full(self.n, 2, np.uint8)]
        return 0

    def _bound(self, node):
        wt = sum(self.w[i] for i in range(self.n) if node[i] == 1)
        if wt > self.cap:
            return None, None
        val = sum(self.p[i] for i in range(self.n) if node[i] == 1)
        room, frac = self.cap - wt, None
        for i in sorted((i for i in range(self.n) if node[i] == 2), key=lambda i: -self.p[i] / self.w[i]):
            if self.w[i] <= room:
                room -= self.w[i]
                val += self.p[i]
            else:
                val += self.p[i] * room / self.w[i]
                frac = i
                break
        return val, frac

    def run(self, max_nodes):
        done = 0
        while self.pool and (max_nodes < 0 or done < max_nodes):
            node = self.pool.pop()
            done += 1
            self.solved += 1
            ub, frac = self._bound(node)
            if ub is None or ub <= self.best + 1e-9:
                continue
            if frac is None:
                self.best, self.has = ub, True
                continue
            for v in (0, 1):
                c = node.copy()
                c[frac] = v
                self.pool.append(c)
        return (0 if not self.pool else 1), done

    def incumbent(self):
        return self.has, (self.best if self.best > -float("inf") else -1.7976931348623157e308)

    def set_cutoff(self, obj):
        if obj > self.best:
            self.best, self.has = obj, False

    def open_count(self):
        return len(self.pool)

    def record_bytes(self):
        return self.n

    def export(self, n):
        keep = 0 if n < 0 else 1
        n = abs(n)
        out = []
        while len(out) < n and len(self.pool) > keep:
            out.append(self.pool.pop())
        return (np.concatenate(out) if out else np.zeros(0, np.uint8)), len(out)

    def import_(self, buf, n):
        for i in range(n):
            self.pool.append(np.array(buf[i * self.n:(i + 1) * self.n], np.uint8))

    def end(self, ret):
        return ret


def brute(w, p, cap):
    best = 0
    for bits in itertools.product((0, 1), repeat=len(w)):
        if sum(a * b for a, b in zip(bits, w)) <= cap:
            best = max(best, sum(a * b for a, b in zip(bits, p)))
    return best


def instance(seed, n=14):
    rng = np.random.default_rng(seed)
    w = rng.integers(5, 40, n)
    p = w + rng.integers(0, 15, n)
    return w, p, int(w.sum() * 0.45)


def test_transfer_plan_is_deterministic_and_conservative():
    assert bnb.transfer_plan([0, 0, 0, 0]) == {}
    assert bnb.transfer_plan([10, 0]) == {0: (1, 5)}
    plan = bnb.transfer_plan([40, 0, 3, 0, 25, 1, 0, 9])
    assert plan == bnb.transfer_plan([40, 0, 3, 0, 25, 1, 0, 9])
    donors = set(plan)
    receivers = {r for r, _ in plan.values()}
    assert not donors & receivers
    for d, (r, k) in plan.items():
        assert 0 < k <= bnb.MAX_SHIP and k <= ([40, 0, 3, 0, 25, 1, 0, 9][d] - [40, 0, 3, 0, 25, 1, 0, 9][r]) // 2


@pytest.mark.parametrize("world", [1, 2, 3])
def test_sharded_search_in_process_matches_brute_force(world):
    w, p, cap = instance(3)
    want = brute(w, p, cap)
    group = bnb.LocalGroup(world)
    results = [None] * world

    def body(rank):
        results[rank] = bnb.sharded_intopt(KnapsackWorker(w, p, cap), group.comm(rank), minimize=False,
                                           slice_nodes=7, ramp_nodes=6)
    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join(60) for t in threads]
    assert all(r is not None for r in results)
    assert all(abs(r["obj"] - want) < 1e-9 for r in results), (want, results)
    assert all(r["ret"] == 0 for r in results)
    assert len({r["holder"] for r in results}) == 1 and results[0]["holder"] is not None
    if world > 1:
        assert sum(r["nodes"] > r["ramp_nodes"] for r in results) >= 2      # work really was shared


def _gloo_rank(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, p, cap = instance(5)
    res = bnb.sharded_intopt(KnapsackWorker(w, p, cap), bnb.TorchComm(), minimize=False, slice_nodes=5, ramp_nodes=6)
    q.put((rank, res["obj"], res["ret"], res["holder"], res["nodes"]))
    dist.destroy_process_group()


def test_sharded_search_over_gloo_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_rank, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    out = [q.get(timeout=120) for _ in range(2)]
    [p.join(30) for p in procs]
    w, p, cap = instance(5)
    want = brute(w, p, cap)
    assert all(abs(o[1] - want) < 1e-9 and o[2] == 0 for o in out), (want, out)
    assert out[0][3] == out[1][3]


@pytest.mark.parametrize("procs,workers", [(1, 3), (2, 2)])
def test_hybrid_comm_two_level_search_matches_brute_force(procs, workers):
    """W worker threads per process x P processes (the processes emulated by a
    second LocalGroup): same optimum, every rank agrees, collectives consistent"""
    w, p, cap = instance(5)
    want = brute(w, p, cap)
    outer_group = bnb.LocalGroup(procs) if procs > 1 else None
    groups = [bnb.LocalGroup(workers) for _ in range(procs)]
    results = {}

    def body(pr, lr):
        outer = outer_group.comm(pr) if outer_group is not None else None
        comm = bnb.HybridComm(groups[pr], lr, outer)
        assert comm.world == procs * workers and comm.rank == pr * workers + lr
        got = comm.allgather_ints([comm.rank, 7])
        assert [g[0] for g in got] == list(range(comm.world)) and all(g[1] == 7 for g in got)
        assert comm.allreduce(float(comm.rank), "max") == comm.world - 1
        bufs = comm.allgather_bytes(np.full(3, comm.rank, np.uint8), 5)
        assert [int(b[0]) for b in bufs] == list(range(comm.world)) and all(len(b) == 5 for b in bufs)
        results[(pr, lr)] = bnb.sharded_intopt(KnapsackWorker(w, p, cap), comm, minimize=False,
                                               slice_nodes=5, ramp_nodes=8)

    threads = [threading.Thread(target=body, args=(pr, lr)) for pr in range(procs) for lr in range(workers)]
    [t.start() for t in threads]
    [t.join(60) for t in threads]
    assert len(results) == procs * workers
    assert all(abs(r["obj"] - want) < 1e-9 and r["ret"] == 0 for r in results.values()), (want, results)
    assert len({r["holder"] for r in results.values()}) == 1
