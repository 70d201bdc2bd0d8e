#!/usr/bin/env python
"""Differential fuzzing of the native LP writer and reader (glpb_write_lp / glpb_read_lp, csrc/lpformat.cpp)
against the unmodified reference (lib/glpcpx.js under minijs).  Build container only; nothing is written.

    python oracle/jsref/fuzz_lpformat.py [first_seed] [count]

Per generated problem: (1) the native writer's lines == the reference writer's lines; (2) that text read back
by the native reader == the problem the reference's own reader builds from it (types, bounds, costs, kinds,
matrix by columns, names)."""
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
import make_writer_golden as W  # noqa: E402
import refjs  # noqa: E402
from minijs import JSThrow  # noqa: E402
import test_presolve as T  # noqa: E402
from glpk_js_b200 import native  # noqa: E402

F = T.F


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    ref = G.Ref()
    bad, rejected, t0 = 0, 0, time.time()
    for seed in range(first, first + count):
        kind = seed % 3
        d = (G.npp_lp(seed, m=3 + seed % 9, n=4 + seed % 11, wild=True) if kind == 0 else
             G.npp_mip(seed, m=4 + seed % 5, n=9 + seed % 4) if kind == 1 else W.odd_numbers(seed))
        P = ref.make(d)
        want = W.written(ref, P)["lines"]
        Q = T.facade_problem(G.problem_arrays(ref, P))
        got = []
        F.glp_set_print_func(None)
        F.glp_write_lp(Q, None, got.append)
        if got != want:
            bad += 1
            print("WRITER MISMATCH seed", seed, flush=True)
            continue
        text = "\n".join(want) + "\n"
        R = ref.call("glp_create_prob")
        try:
            rc, err = refjs.read_lp_text(ref.I, R, text), None
        except JSThrow as e:
            rc, err = 1, str(e)
        if rc != 0:
            # the reference's reader rejects what its writer wrote (e.g. a variable named like a number's
            # exponent tail): the native reader must reject it with the same message
            try:
                native.read_lp(text)
                nerr = None
            except ValueError as e:
                nerr = str(e)
            rejected += 1
            if nerr is None or (err is not None and err.split(": ", 1)[-1] not in nerr):
                bad += 1
                print("READER ERROR MISMATCH seed", seed, repr(err), repr(nerr), flush=True)
            continue
        ra = G.problem_arrays(ref, R)
        na, names = native.read_lp(text)
        m = ra["m"]
        same = (na["m"], na["n"], na["dir"]) == (ra["m"], ra["n"], ra["dir"]) and \
            list(na["type"][:m]) == ra["r_type"] and list(na["type"][m:]) == ra["c_type"] and \
            list(na["lb"][:m]) == ra["r_lb"] and list(na["ub"][:m]) == ra["r_ub"] and \
            list(na["lb"][m:]) == ra["c_lb"] and list(na["ub"][m:]) == ra["c_ub"] and \
            list(na["coef"]) == ra["c_coef"] and list(na["kind"]) == ra["c_kind"] and \
            list(na["A_ptr"]) == ra["A_ptr"] and list(na["A_ind"]) == ra["A_ind"] and list(na["A_val"]) == ra["A_val"]
        rnames = [R["row"][i]["name"] for i in range(1, m + 1)] + [R["col"][j]["name"] for j in range(1, ra["n"] + 1)]
        same = same and [js for js in rnames] == names["rows"] + names["cols"]
        if not same:
            bad += 1
            print("READER MISMATCH seed", seed, flush=True)
    print("%d problems in %.0f s (%d texts rejected by both readers), mismatches: %d" % (count, time.time() - t0, rejected, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
