#!/usr/bin/env python
"""Bases of the ORACLE after K iterations of C3 (covering LP 16384 x 32768, dual simplex), K = 20000 and 30000,
from one uninterrupted oracle run in the build container (about half an hour, one thread) -> tests/golden/c3_basis_K.npz.
The GPU test test_c3_full_size_basis_equals_the_oracles compares the device's basis after the same number of
iterations (smaller K are computed by the oracle inside the test)."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402
import helpers as H     # noqa: E402

d = O.generate("covering", m=16384, n=32768, kmin=8, kspan=17, seed=20240601)
m, n = d["m"], d["n"]
P = O.Problem.from_arrays(H.to_oracle(d))
want = [int(x) for x in sys.argv[1:]] or [20000, 30000]
state = {"n": 0}


def hook(ev, csa):
    if ev != O.EV_D_ITER:
        return
    state["n"] += 1
    it = state["n"]
    if it in want:
        head = O.csa_get(csa, "head")
        nst = O.csa_get(csa, "stat")
        stat = np.zeros(m + n, np.int8)
        stat[head[1:m + 1] - 1] = O.GLP_BS
        stat[head[m + 1:m + n + 1] - 1] = nst[1:n + 1]
        np.savez_compressed(os.path.join(HERE, "c3_basis_%d.npz" % it), stat=stat, it=it)
        print("saved", it, flush=True)


P.set_hook(hook)
P.simplex(meth=O.GLP_DUAL, it_lim=max(want) + 1)
