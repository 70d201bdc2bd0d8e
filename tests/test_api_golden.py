"""The facade's problem-building calls against THE REFERENCE'S OWN problem objects (tests/golden/
ref_api_cases.json: 40 random sequences of glp_add_rows/cols, glp_set_*_bnds, glp_set_obj_coef,
glp_set_mat_row/col, glp_load_matrix, glp_sort_matrix, glp_set_col_kind, glp_set_*_stat, glp_std_basis,
glp_del_rows/cols applied to lib/glpapi01/05/09.js under minijs, oracle/jsref/fuzz_api.py --golden).
Everything the input contract of the hot path reads must come out the same: types, bounds, costs, kinds,
statuses, nnz and the ORDER of every row and column list (scaling, crash basis and ties depend on it)."""
import json
import os

import pytest

from glpk_js_b200 import glpk as F
import helpers as H

with open(os.path.join(H.GOLDEN, "ref_api_cases.json")) as f:
    CASES = {k: v for k, v in json.load(f).items() if not k.startswith("_")}


@pytest.mark.parametrize("name", sorted(CASES, key=lambda s: int(s.split("_")[1])))
def test_call_sequence_leaves_the_references_problem_object(name):
    case = CASES[name]
    P = F.glp_create_prob()
    for fn, args in case["calls"]:
        getattr(F, fn)(P, *args)
    st = case["state"]
    assert (P.m, P.n, P.nnz, P.dir, P.c0) == (st["m"], st["n"], st["nnz"], st["dir"], st["c0"])
    for i, (typ, lb, ub, stat, lst) in enumerate(st["rows"], 1):
        r = P.row[i]
        assert (r.type, r.lb, r.ub, r.stat, [list(e) for e in r.elems]) == (typ, lb, ub, stat, lst), ("row", i)
    for j, (typ, lb, ub, stat, coef, kind, lst) in enumerate(st["cols"], 1):
        c = P.col[j]
        assert (c.type, c.lb, c.ub, c.stat, c.coef, c.kind, [list(e) for e in c.elems]) == \
            (typ, lb, ub, stat, coef, kind, lst), ("col", j)


def test_the_sequences_reach_every_call():
    seen = {fn for c in CASES.values() for fn, _ in c["calls"]}
    assert {"glp_load_matrix", "glp_set_mat_row", "glp_set_mat_col", "glp_sort_matrix", "glp_del_rows", "glp_del_cols",
            "glp_set_col_kind", "glp_set_row_stat", "glp_set_col_stat", "glp_std_basis"} <= seen
