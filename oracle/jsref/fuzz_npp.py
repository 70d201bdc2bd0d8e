#!/usr/bin/env python
"""Differential fuzzing of the native presolver (glpb_npp_*, csrc/presolve.cpp) against the unmodified
reference presolver (lib/glpnpp01-05.js under minijs).  Build container only; nothing is written.

    python oracle/jsref/fuzz_npp.py [first_seed] [count]

Every generated case goes through make_npp_golden.run_case (the reference) and through the comparison of
tests/test_presolve.py (the product): return code, recovery-stack depth, reduced problem incl. element
order, recovery, unloaded solution -- all bit for bit."""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import make_npp_golden as G  # noqa: E402
import test_presolve as T  # noqa: E402


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    ref = G.Ref()
    t0, bad, kinds = time.time(), 0, {}
    for seed in range(first, first + count):
        for tag in ("lp", "mip"):
            try:
                if tag == "lp":
                    case = G.run_case(ref, G.npp_lp(seed, m=6 + seed % 23, n=8 + seed % 31, wild=seed % 3 == 0), G.GLP_SOL)
                else:
                    case = G.run_case(ref, G.npp_mip(seed, m=6 + seed % 9, n=9 + seed % 8, wide=seed % 5 == 0), G.GLP_MIP, binarize=seed % 2)
            except Exception as e:      # the reference itself threw (e.g. xassert): not a case
                print("seed", seed, tag, "reference raised", type(e).__name__, str(e)[:80], flush=True)
                continue
            name = "fuzz_%s_%d" % (tag, seed)
            T.CASES[name] = case
            try:
                T.test_presolver_matches_reference(name)
            except AssertionError as e:
                bad += 1
                print("MISMATCH", name, str(e)[:300], flush=True)
            kinds[case["ret"]] = kinds.get(case["ret"], 0) + 1
            del T.CASES[name]
    print("%d cases in %.0f s, return codes %s, mismatches: %d" % (2 * count, time.time() - t0, kinds, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
