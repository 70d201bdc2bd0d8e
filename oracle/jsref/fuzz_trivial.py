#!/usr/bin/env python
"""Differential fuzzing of the host-only paths of glp_simplex / glp_intopt in the facade against the
unmodified reference (lib/glpapi06.js:148-339, lib/glpapi09.js:258-390 under minijs): problems WITHOUT
constraint coefficients (trivial_lp), inconsistent double bounds (GLP_EBOUND), integer columns with
fractional bounds -- return code, every printed line, statuses and values of all rows and columns.
Build container only; nothing is written.

    python oracle/jsref/fuzz_trivial.py [first_seed] [count]"""
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
from minijs import NativeFunc, js_to_str  # noqa: E402
from glpk_js_b200 import glpk as F  # noqa: E402


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    ref = G.Ref()
    bad, skipped, t0 = 0, 0, time.time()
    for seed in range(first, first + count):
        rs = np.random.RandomState(seed)
        m, n = int(rs.randint(0, 5)), int(rs.randint(1, 7))
        R, Q = ref.call("glp_create_prob"), F.glp_create_prob()

        def both(name, *a):
            ref.call(name, R, *a)
            getattr(F, name)(Q, *a)
        if m:
            both("glp_add_rows", m)
        both("glp_add_cols", n)
        both("glp_set_obj_dir", int(rs.choice([1, 2])))
        both("glp_set_obj_coef", 0, float(rs.randint(-3, 4)))
        for i in range(1, m + 1):
            lo = float(rs.randint(-4, 4))
            both("glp_set_row_bnds", i, int(rs.randint(1, 6)), lo, lo + float(rs.choice([-1, 0, 1, 2, 0.5])))
        for j in range(1, n + 1):
            lo = float(rs.choice([-3, -1, 0, 0, 1, 2.5]))
            both("glp_set_col_bnds", j, int(rs.randint(1, 6)), lo, lo + float(rs.choice([-1, 0, 1, 2, 0.5, 3])))
            both("glp_set_obj_coef", j, float(rs.choice([-2, -1, 0, 0, 1, 3, 1e-9])))
            if rs.rand() < 0.3:
                both("glp_set_col_kind", j, 2)
        use_mip = rs.rand() < 0.4
        lev = int(rs.choice([0, 1, 2, 3]))
        presolve = int(rs.rand() < 0.5)
        rl, fl = [], []
        saved = ref.call("glp_get_print_func")
        ref.call("glp_set_print_func", NativeFunc(lambda this, a: rl.append(js_to_str(a[0])), "print"))
        try:
            rret = [int(ref.call("glp_simplex", R, ref.smcp(msg_lev=lev, presolve=presolve)))]
            if use_mip:
                try:
                    rret.append(int(ref.call("glp_intopt", R, ref.iocp(msg_lev=lev, presolve=presolve))))
                except Exception as e:
                    rret.append("throw")
        finally:
            ref.call("glp_set_print_func", saved)
        F.glp_set_print_func(fl.append)
        try:
            p = F.SMCP({"presolve": presolve})
            p.msg_lev = lev
            fret = [F.glp_simplex(Q, p)]
            if use_mip:
                io = F.IOCP({"presolve": presolve})
                io.msg_lev = lev
                try:
                    fret.append(F.glp_intopt(Q, io))
                except F.GlpkError:
                    fret.append("throw")
                except RuntimeError:        # the branch-and-bound proper needs the device: not a host-only path
                    skipped += 1
                    continue
        finally:
            F.glp_set_print_func(None)
        rs_ = ([(int(R["row"][i]["stat"]), float(R["row"][i]["prim"]), float(R["row"][i]["dual"])) for i in range(1, m + 1)],
               [(int(R["col"][j]["stat"]), float(R["col"][j]["prim"]), float(R["col"][j]["dual"])) for j in range(1, n + 1)],
               int(R["pbs_stat"]), int(R["dbs_stat"]), float(R["obj_val"]), int(R["some"]))
        fs_ = ([(r.stat, r.prim, r.dual) for r in Q.row[1:]], [(c.stat, c.prim, c.dual) for c in Q.col[1:]],
               Q.pbs_stat, Q.dbs_stat, Q.obj_val, Q.some)
        if use_mip and rret[-1] != "throw":
            rs_ += (int(R["mip_stat"]), float(R["mip_obj"]), [float(R["col"][j]["mipx"]) for j in range(1, n + 1)])
            fs_ += (Q.mip_stat, Q.mip_obj, [c.mipx for c in Q.col[1:]])
        if rret != fret or rl != fl or rs_ != fs_:
            bad += 1
            print("MISMATCH seed", seed, "ret", rret, fret, "lev", lev, "presolve", presolve, flush=True)
            if rl != fl:
                for a, b in zip(rl + [None] * 9, fl + [None] * 9):
                    if a != b:
                        print("    line", repr(a), "|", repr(b))
                        break
            elif rs_ != fs_:
                for k, (a, b) in enumerate(zip(rs_, fs_)):
                    if a != b:
                        print("    field", k, a, "|", b)
                        break
    print("%d problems in %.0f s (%d needed the device and were skipped), mismatches: %d" % (count, time.time() - t0, skipped, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
