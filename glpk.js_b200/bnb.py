"""Node-sharded branch-and-bound across the GPUs of one box.

A single LP never spans GPUs; the only part of the path that shards is the
``ios_driver`` search (lib/glpios03.js:507-951): open nodes are independent
units.  Every rank owns a device-resident copy of the problem and a local pool
of open nodes (``glpb_mip_*`` in the C ABI).  The search proceeds in rounds:

1. every rank solves up to ``slice`` node LPs from its own pool;
2. incumbent exchange: all-reduce (min or max) of the best objective, applied
   as a cut-off on the ranks that did not find it;
3. load balancing: pool sizes are all-gathered, every rank derives the same
   transfer plan, donors export self-contained node records and one padded
   all-gather moves them (records are ~18(m+n) bytes, so this is latency- not
   bandwidth-bound even over NCCL/NVLink);
4. termination when every pool is empty.

Ramp-up is replicated: all ranks run the identical deterministic search until
there are enough open nodes, then keep every ``world``-th one.  The explored
tree differs from the serial one, the optimum (an objective *value*) does not.

The communicator is abstract so that the logic runs under ``torch.distributed``
(NCCL on GPUs, gloo in CPU tests) or in-process (``LocalGroup``) for tests that
emulate several ranks on one device.
"""
import numpy as np

from . import native

MAX_SHIP = 16        # node records a donor ships per round


class TorchComm:
    """torch.distributed adapter (backend nccl on GPUs, gloo on CPU)."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = device if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))

    def allreduce(self, value, op):
        t = self.torch.tensor([value], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op={"min": self.dist.ReduceOp.MIN, "max": self.dist.ReduceOp.MAX,
                                    "sum": self.dist.ReduceOp.SUM}[op])
        return float(t.item())

    def allgather_ints(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.int64, device=self.device)
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [[int(v) for v in o.tolist()] for o in out]

    def allgather_bytes(self, buf, nbytes):
        t = self.torch.zeros(nbytes, dtype=self.torch.uint8, device=self.device)
        if len(buf):
            t[:len(buf)] = self.torch.from_numpy(np.ascontiguousarray(buf)).to(self.device)
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [o.cpu().numpy() for o in out]


def transfer_plan(counts, low=1, max_ship=MAX_SHIP):
    """Deterministic donor -> receiver plan from the gathered pool sizes.
    Returns {donor: (receiver, n)}; a rank takes part in one transfer a round."""
    order = sorted(range(len(counts)), key=lambda r: (counts[r], r))
    plan, lo, hi = {}, 0, len(order) - 1
    counts = list(counts)
    while lo < hi:
        poor, rich = order[lo], order[hi]
        if counts[poor] > low or counts[rich] - counts[poor] < 2:
            break
        n = min(max_ship, (counts[rich] - counts[poor]) // 2)
        if n > 0:
            plan[rich] = (poor, n)
        lo += 1
        hi -= 1
    return plan


class Worker:
    """What the driver needs from a rank; the GPU implementation wraps a
    ``native.Problem``, tests substitute a fake."""

    def __init__(self, prob):
        self.P = prob

    def begin(self, **iocp):
        return self.P.mip_begin(**iocp)

    def run(self, max_nodes):
        return self.P.mip_run(max_nodes)

    def incumbent(self):
        return self.P.mip_incumbent()

    def set_cutoff(self, obj):
        self.P.mip_set_cutoff(obj)

    def open_count(self):
        return self.P.mip_open_count()

    def record_bytes(self):
        return self.P.L.glpb_mip_record_bytes(self.P.h)

    def export(self, n):
        return self.P.mip_export(n)

    def import_(self, buf, n):
        self.P.mip_import(buf, n)

    def end(self, ret):
        return self.P.mip_end(ret)

    def solution(self):
        return self.P.mip()


def sharded_intopt(worker, comm, minimize, slice_nodes=64, ramp_nodes=None, node_lim=None, **iocp):
    """Run the sharded search on this rank.  Returns a dict with the global
    optimum, who holds the solution, node counts and the per-round log."""
    rank, world = comm.rank, comm.world
    rc = worker.begin(**iocp)
    if rc != 0:
        return dict(ret=rc, obj=None, nodes=0, total_nodes=0, rounds=0)
    rb = worker.record_bytes()
    inf = float("inf")
    my_nodes, rounds, ret = 0, 0, 0
    # ---- replicated ramp-up ----
    if world > 1:
        want = ramp_nodes if ramp_nodes is not None else 4 * world
        state = 1
        while state == 1 and worker.open_count() < want:
            state, solved = worker.run(1)
            my_nodes += solved
        if state > 1:
            ret = state
        if worker.open_count() > 0 and ret == 0:
            buf, cnt = worker.export(-(1 << 20))          # pop every open node (same list on all ranks)
            keep = [i for i in range(cnt) if i % world == rank]
            if keep:
                rec = np.concatenate([buf[i * rb:(i + 1) * rb] for i in keep])
                worker.import_(rec, len(keep))
        ramp = my_nodes
    else:
        ramp = 0
    # ---- sharded rounds ----
    while ret == 0:
        rounds += 1
        budget = slice_nodes
        if node_lim is not None:
            budget = max(0, min(budget, node_lim - my_nodes))
        state, solved = worker.run(budget) if budget > 0 else (1, 0)
        my_nodes += solved
        if state > 1:
            ret = state
        has, obj = worker.incumbent()
        key = obj if minimize else -obj
        best = comm.allreduce(key if obj not in (inf, -inf) and abs(obj) < 1e300 else inf, "min")
        if best < inf:
            worker.set_cutoff(best if minimize else -best)
        stop = comm.allreduce(float(ret != 0 or (node_lim is not None and my_nodes >= node_lim)), "max")
        counts = [c[0] for c in comm.allgather_ints([worker.open_count()])]
        if stop > 0 or sum(counts) == 0:
            break
        plan = transfer_plan(counts)
        ship, dst, n = np.zeros(0, np.uint8), -1, 0
        if rank in plan:
            dst, want_n = plan[rank]
            ship, n = worker.export(want_n)
        meta = comm.allgather_ints([dst, n])
        if any(mm[1] > 0 for mm in meta):
            bufs = comm.allgather_bytes(ship, MAX_SHIP * rb)
            for src, (d, cnt) in enumerate(meta):
                if d == rank and cnt > 0:
                    worker.import_(bufs[src][:cnt * rb], cnt)
    # ---- wrap up: global optimum and its holder ----
    has, obj = worker.incumbent()
    key = (obj if minimize else -obj) if abs(obj) < 1e300 else inf
    best = comm.allreduce(key, "min")
    holder = comm.allreduce(float(rank) if (has and key == best) else float(world), "min")
    total = comm.allreduce(float(my_nodes), "sum")
    any_err = comm.allreduce(float(ret), "max")
    worker.end(int(any_err))
    return dict(ret=int(any_err), obj=(None if best == inf else (best if minimize else -best)),
                holder=(int(holder) if holder < world else None), nodes=my_nodes,
                total_nodes=int(total) - ramp * (world - 1), rounds=rounds, ramp_nodes=ramp)


# ---- in-process emulation of several ranks (tests; one GPU or none) ----
class LocalGroup:
    """Runs ``world`` drivers as cooperative threads with barrier-style
    collectives, so the N>1 protocol can be exercised without N devices."""

    def __init__(self, world):
        import threading
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world
        self.lock = threading.Lock()

    def comm(self, rank):
        return _LocalComm(self, rank)


class _LocalComm:
    def __init__(self, group, rank):
        self.g, self.rank, self.world = group, rank, group.world

    def _exchange(self, value):
        self.g.slots[self.rank] = value
        self.g.barrier.wait()
        out = list(self.g.slots)
        self.g.barrier.wait()
        return out

    def allreduce(self, value, op):
        vals = self._exchange(value)
        return {"min": min, "max": max, "sum": sum}[op](vals)

    def allgather_ints(self, values):
        return self._exchange(list(values))

    def allgather_bytes(self, buf, nbytes):
        b = np.zeros(nbytes, np.uint8)
        b[:len(buf)] = buf
        return self._exchange(b)


class HybridComm:
    """W worker threads per process (each with its own device handle on the SAME
    GPU: node LPs of the C4/C5 sizes occupy one SM, so several run concurrently)
    times P processes (one per GPU, ``torch.distributed``).  Global rank =
    process rank * W + worker.  Collectives run in two levels: the workers of a
    process meet in a ``LocalGroup``; worker 0 alone talks to the other
    processes."""

    def __init__(self, group, local_rank, outer=None):
        self.g, self.lr, self.outer = group, local_rank, outer
        self.local = group.comm(local_rank)
        self.W = group.world
        self.P = outer.world if outer is not None else 1
        self.pr = outer.rank if outer is not None else 0
        self.rank, self.world = self.pr * self.W + local_rank, self.P * self.W

    def allreduce(self, value, op):
        red = {"min": min, "max": max, "sum": sum}[op]
        part = red(self.local._exchange(value))
        if self.outer is None:
            return part
        res = self.outer.allreduce(part, op) if self.lr == 0 else None
        return self.local._exchange(res)[0]

    def allgather_ints(self, values):
        loc = self.local._exchange(list(values))               # [W][len]
        if self.outer is None:
            return loc
        flat = None
        if self.lr == 0:
            got = self.outer.allgather_ints([v for row in loc for v in row])   # [P][W*len]
            n = len(values)
            flat = [row[i * n:(i + 1) * n] for row in got for i in range(self.W)]
        return self.local._exchange(flat)[0]

    def allgather_bytes(self, buf, nbytes):
        b = np.zeros(nbytes, np.uint8)
        b[:len(buf)] = buf
        loc = self.local._exchange(b)                          # [W] arrays
        if self.outer is None:
            return loc
        out = None
        if self.lr == 0:
            got = self.outer.allgather_bytes(np.concatenate(loc), nbytes * self.W)   # [P] arrays
            out = [g[i * nbytes:(i + 1) * nbytes] for g in got for i in range(self.W)]
        return self.local._exchange(out)[0]


# ======================================================================
# Batched engine (glpb_bnb_*: one CTA per node, node states device-resident)
# ======================================================================
HDR_DOUBLES = 8          # [key, open, stop, solved, n_ship, dst, has_sol, spare]


class BatchWorker:
    """Adapter over ``native.Problem.bnb_*``; tests substitute the host
    emulation of the same engine (tests/ne_emul.py)."""

    def __init__(self, prob):
        self.P = prob

    def begin(self, batch=0, slab_nodes=0, **iocp):
        return self.P.bnb_begin(batch=batch, slab_nodes=slab_nodes, **iocp)

    def clear(self):
        self.P.bnb_clear()

    def round(self, max_tasks=-1):
        return self.P.bnb_round(max_tasks)

    def incumbent(self):
        return self.P.bnb_incumbent()

    def set_cutoff(self, obj):
        self.P.bnb_set_cutoff(obj)

    def open_count(self):
        return self.P.bnb_open_count()

    def record_bytes(self):
        return self.P.bnb_record_bytes()

    def export_to(self, ptr, n):
        return self.P.bnb_export(n, ptr)

    def import_from(self, ptr, n):
        self.P.bnb_import(ptr, n)

    def solved(self):
        return self.P.bnb_stats()["solved"]

    def end(self, ret):
        return self.P.bnb_end(ret)


class TensorComm:
    """One fused all-gather per exchange over torch.distributed (NCCL: the node
    payload goes device -> NVLink -> device; gloo in CPU tests).  ``world`` 1
    needs no process group."""

    def __init__(self, device=None):
        import torch
        self.torch = torch
        try:
            import torch.distributed as dist
            self.dist = dist if dist.is_available() and dist.is_initialized() else None
        except Exception:
            self.dist = None
        self.rank = self.dist.get_rank() if self.dist else 0
        self.world = self.dist.get_world_size() if self.dist else 1
        if device is None:
            device = (torch.device("cuda", torch.cuda.current_device())
                      if (self.dist and self.dist.get_backend() == "nccl") or (not self.dist and torch.cuda.is_available())
                      else torch.device("cpu"))
        self.device = device

    def buffers(self, nbytes):
        t = self.torch
        send = t.zeros(nbytes, dtype=t.uint8, device=self.device)
        recv = t.zeros(nbytes * self.world, dtype=t.uint8, device=self.device)
        return send, recv

    def write_header(self, send, values):
        t = self.torch
        h = t.tensor(values, dtype=t.float64).view(t.uint8)
        send[:h.numel()].copy_(h)

    def allgather(self, send, recv):
        if self.dist:
            self.dist.all_gather_into_tensor(recv, send)
        else:
            recv.copy_(send)

    def read_headers(self, recv, nbytes):
        t = self.torch
        rows = recv.view(self.world, nbytes)[:, :8 * HDR_DOUBLES].contiguous().cpu()      # the one host sync of an exchange
        return rows.view(t.float64).view(self.world, HDR_DOUBLES).tolist()

    def sync(self):
        if self.device.type == "cuda":
            self.torch.cuda.synchronize(self.device)


def sharded_bnb_batched(worker, comm, minimize, batch=0, slab_nodes=0, node_lim=None, max_ship=64,
                        exchange_every=1, max_rounds=None, **iocp):
    """The sharded search on this rank with the batched engine.

    Rank 0 starts from the root, the others start empty and are fed by
    migration.  A *round* is one launch of the node kernel over up to ``batch``
    open nodes of the local pool.  Every ``exchange_every`` rounds the ranks meet
    in ONE all-gather of [header | up to max_ship node records]: the header
    carries the incumbent objective, the pool size, stop flags and the number of
    records shipped to which rank.  The transfer plan is derived by every rank
    from the same gathered pool sizes and executed with the next exchange.
    ``node_lim`` is a per-rank budget of node LPs (weak scaling).  Returns the
    global optimum value, who holds the solution vector, and node counts."""
    rank, world = comm.rank, comm.world
    rc = worker.begin(batch=batch, slab_nodes=slab_nodes, **iocp)
    if rc != 0:
        return dict(ret=rc, obj=None, nodes=0, total_nodes=0, rounds=0, exchanges=0)
    if rank != 0:
        worker.clear()
    rb = worker.record_bytes()
    hdr_bytes = 8 * HDR_DOUBLES
    nbytes = hdr_bytes + max_ship * rb
    send, recv = comm.buffers(nbytes)
    inf = float("inf")
    rounds = exchanges = 0
    ret = 0
    my_plan = None                      # (dst, n) decided at the previous exchange
    moved_in = moved_out = 0
    while True:
        # ---- local work ----
        for _ in range(exchange_every):
            if ret != 0:
                break
            left = -1
            if node_lim is not None:
                left = node_lim - worker.solved()
                if left <= 0:
                    break
            state, done = worker.round(left)
            if state == 1:
                rounds += 1
            elif state == 0:
                break                   # local pool empty
            else:
                ret = state
        if world == 1 and node_lim is None and ret == 0 and worker.open_count() > 0 and (max_rounds is None or rounds < max_rounds):
            continue                    # nothing to exchange with
        # ---- one fused exchange ----
        has, obj = worker.incumbent()
        key = (obj if minimize else -obj) if abs(obj) < 1e300 else inf
        n_ship, dst = 0, -1
        if my_plan is not None and ret == 0:
            dst, want = my_plan
            want = min(want, max_ship, max(0, worker.open_count() - 1))
            if want > 0:
                n_ship = worker.export_to(send.data_ptr() + hdr_bytes, want)
            if n_ship == 0:
                dst = -1
        solved = worker.solved()
        at_limit = node_lim is not None and solved >= node_lim
        stop = 1.0 if (ret != 0 or (max_rounds is not None and rounds >= max_rounds)) else 0.0
        comm.write_header(send, [key, float(worker.open_count()), stop, float(solved), float(n_ship), float(dst),
                                 1.0 if has else 0.0, 1.0 if at_limit else 0.0])
        comm.allgather(send, recv)
        hdrs = comm.read_headers(recv, nbytes)
        exchanges += 1
        moved_out += n_ship
        for src, h in enumerate(hdrs):
            if int(h[5]) == rank and int(h[4]) > 0 and src != rank:
                worker.import_from(recv.data_ptr() + src * nbytes + hdr_bytes, int(h[4]))
                moved_in += int(h[4])
        best = min(h[0] for h in hdrs)
        if best < inf:
            worker.set_cutoff(best if minimize else -best)
        counts = [int(h[1]) for h in hdrs]
        counts[rank] = worker.open_count()
        in_flight = sum(int(h[4]) for h in hdrs)
        any_stop = any(h[2] > 0 for h in hdrs)
        all_limit = all(h[7] > 0 for h in hdrs)
        if any_stop or all_limit:
            break
        if sum(int(h[1]) for h in hdrs) == 0 and in_flight == 0:
            break
        # every rank derives the same plan from the same numbers (its own count as gathered)
        plan = transfer_plan([int(h[1]) for h in hdrs], low=max(1, (batch or 1)), max_ship=max_ship)
        my_plan = plan.get(rank)
    # ---- wrap up: optimum and its holder ----
    has, obj = worker.incumbent()
    key = (obj if minimize else -obj) if abs(obj) < 1e300 else inf
    solved = worker.solved()
    comm.write_header(send, [key, float(worker.open_count()), float(ret), float(solved), 0.0, -1.0, 1.0 if has else 0.0, 0.0])
    comm.allgather(send, recv)
    hdrs = comm.read_headers(recv, nbytes)
    best = min(h[0] for h in hdrs)
    holder = next((r for r, h in enumerate(hdrs) if h[6] > 0 and h[0] == best), None)
    total = int(sum(h[3] for h in hdrs))
    err = int(max(h[2] for h in hdrs))
    open_left = int(sum(h[1] for h in hdrs))
    worker.end(err)
    return dict(ret=err, obj=(None if best == inf else (best if minimize else -best)), holder=holder, nodes=solved,
                total_nodes=total, rounds=rounds, exchanges=exchanges, moved_in=moved_in, moved_out=moved_out,
                open_left=open_left)
