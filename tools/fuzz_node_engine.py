#!/usr/bin/env python
"""Fuzz of the product's batched branch-and-bound (csrc/nodeengine.cuh + bnbpool.cuh, compiled for the host by
tests/emul) against the CPU oracle (itself bit-identical to the reference, oracle/jsref/fuzz_oracle.py):
generated MIPs, several batch sizes; the optimum must agree, the returned point must be feasible and integral.

    python tools/fuzz_node_engine.py [first_seed] [count]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "jsref"))

import helpers as H  # noqa: E402
import ne_emul as NE  # noqa: E402
import oracle_lib as O  # noqa: E402
import make_npp_golden as G  # noqa: E402  (problem generators only; the reference is not loaded)


def as_np(d):
    out = dict(d)
    for k in ("r_type", "c_type", "c_kind", "A_ptr", "A_ind"):
        out[k] = np.array(d[k], dtype=np.int32)
    for k in ("r_lb", "r_ub", "c_lb", "c_ub", "c_coef", "A_val"):
        out[k] = np.array(d[k], dtype=np.float64)
    return out


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    bad, skipped, refused, t0 = 0, 0, 0, time.time()
    for seed in range(first, first + count):
        k = seed % 3
        if k == 0:
            d = as_np(G.npp_mip(seed, m=5 + seed % 7, n=9 + seed % 9))
        elif k == 1:
            d = H.random_mip(seed)
        else:
            d = H.to_oracle(O.generate("mkp", m=2 + seed % 4, n=10 + seed % 14, seed=seed))
        dn = H.to_native(d)
        Q = O.Problem.from_arrays(d)
        rc = Q.simplex(meth=O.GLP_PRIMAL)
        root = Q.solution()
        if rc != 0 or root["status"] != O.GLP_OPT:
            skipped += 1
            continue
        t1 = time.time()
        oret = Q.intopt()
        omp = Q.mip()
        t_oracle = time.time() - t1
        for batch in (1, 7, 32):
            try:
                P = NE.Pool(dn, root["stat"], batch=batch, cap=1 << 16)
            except Exception as e:
                refused += 1
                break
            ret = P.run()
            inc = P.incumbent()
            ok = ret == oret
            if omp["mip_stat"] == O.GLP_NOFEAS:
                ok = ok and not inc["have_sol"]
            else:
                ok = ok and inc["have_sol"] and abs(inc["obj"] - omp["mip_obj"]) <= 1e-9 * max(1.0, abs(omp["mip_obj"]))
                if ok:
                    x = inc["x"][dn["m"]:]
                    ints = np.asarray(dn["kind"]) == 2
                    ok = bool(np.all(x[ints] == np.round(x[ints])))
            if time.time() - t1 > 20:
                print("slow: seed", seed, "batch", batch, "oracle %.1f s, so far %.1f s" % (t_oracle, time.time() - t1), flush=True)
            if not ok:
                bad += 1
                print("MISMATCH seed", seed, "batch", batch, "ret", ret, oret, "oracle", omp["mip_stat"], omp["mip_obj"],
                      "engine", inc["have_sol"], inc.get("obj"), flush=True)
                break
    print("%d problems in %.0f s (%d without an optimal root, %d not taken by the engine), mismatches: %d"
          % (count, time.time() - t0, skipped, refused, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
