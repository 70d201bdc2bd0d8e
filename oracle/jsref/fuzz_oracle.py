#!/usr/bin/env python
"""Differential fuzzing of the CPU ORACLE (oracle/*.cpp, the C++ restatement of the simplex path that the
GPU tests check the device against) against the unmodified reference (lib/glpspx01.js, glpspx02.js,
glpbfd.js, glpfhv.js, glpluf.js, glpios*.js under minijs): random LPs / MIPs, all three methods, presolve
OFF.  Return code, status, ITERATION COUNT, objective, statuses and values must be identical -- the values
bit for bit, the oracle follows the reference's arithmetic statement by statement.  Build container only.

    python oracle/jsref/fuzz_oracle.py [first_seed] [count]
    python oracle/jsref/fuzz_oracle.py --families [first_seed] [count]    packing / covering / knapsack LPs of the
                                                   bench's own generators (oracle/gen.cpp), 20..60 rows"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import make_npp_golden as G  # noqa: E402
import oracle_lib as O  # noqa: E402


def as_np(d):
    out = dict(d)
    for k in ("r_type", "c_type", "c_kind", "A_ptr", "A_ind"):
        out[k] = np.array(d[k], dtype=np.int32)
    for k in ("r_lb", "r_ub", "c_lb", "c_ub", "c_coef", "A_val"):
        out[k] = np.array(d[k], dtype=np.float64)
    return out


def family(seed):
    import helpers as H
    k = seed % 3
    if k == 0:
        return H.to_oracle(O.generate("packing", m=20 + seed % 25, n=40 + seed % 50, density=0.2, seed=seed)), 1, False
    if k == 1:
        return H.to_oracle(O.generate("covering", m=30 + seed % 30, n=60 + seed % 60, kmin=3, kspan=4, seed=seed)), 3, False
    return H.to_oracle(O.generate("mkp", m=3 + seed % 3, n=12 + seed % 8, seed=seed)), 1, True


def main():
    fam = "--families" in sys.argv
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    first = int(argv[0]) if len(argv) > 0 else 1
    count = int(argv[1]) if len(argv) > 1 else 100
    ref = G.Ref()
    bad, inexact, t0, its = 0, 0, time.time(), 0
    for seed in range(first, first + count):
        if fam:
            d, meth, mip = family(seed)
        else:
            mip = seed % 4 == 0
            d = G.npp_mip(seed, m=5 + seed % 6, n=8 + seed % 5) if mip else \
                G.npp_lp(seed, m=5 + seed % 14, n=7 + seed % 19, wild=seed % 3 == 0)
            meth = [1, 3, 2][seed % 3]
        R = ref.make(d)
        pa = G.problem_arrays(ref, R)
        rret = int(ref.call("glp_simplex", R, ref.smcp(meth=meth)))
        rres = ref.lp_result(R)
        Q = O.Problem.from_arrays(as_np(pa))
        oret = Q.simplex(meth=meth)
        o = Q.solution()
        m, n = pa["m"], pa["n"]
        its += rres["it_cnt"]
        same = (oret, o["status"], o["it_cnt"]) == (rret, rres["status"], rres["it_cnt"]) and \
            list(o["stat"][:m]) == rres["row_stat"] and list(o["stat"][m:]) == rres["col_stat"]
        exact = same and o["obj"] == rres["obj"] and list(o["prim"][m:]) == rres["col_prim"] and \
            list(o["dual"][:m]) == rres["row_dual"]
        if same and mip and rret == 0 and rres["status"] == 5:
            r2 = int(ref.call("glp_intopt", R, ref.iocp()))
            rm = ref.mip_result(R)
            o2 = Q.intopt()
            om = Q.mip()
            same = (o2, om["mip_stat"]) == (r2, rm["mip_stat"]) and (om["mip_stat"] not in (2, 5) or om["mip_obj"] == rm["mip_obj"])
        if not same:
            bad += 1
            print("MISMATCH seed", seed, "meth", meth, "ret", rret, oret, "status", rres["status"], o["status"],
                  "it", rres["it_cnt"], o["it_cnt"], "obj", rres["obj"], o["obj"], flush=True)
        elif not exact:
            inexact += 1
    print("%d problems (%d reference iterations) in %.0f s: mismatches %d, equal but not bit-identical values %d"
          % (count, its, time.time() - t0, bad, inexact))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
