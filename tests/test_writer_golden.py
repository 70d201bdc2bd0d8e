"""glp_write_lp of the facade against the REFERENCE'S OWN writer (tests/golden/ref_writer_cases.json:
lib/glpcpx.js:755-999 run by minijs, oracle/jsref/make_writer_golden.py): every line, the return code and
the two messages, on the three fixtures (names kept, read through the native reader), on generated LPs /
MIPs with every row and column type, empty rows / columns and ranged rows, and on coefficients that
exercise JavaScript's number-to-string rules (17 digits, 1e21, 1e-7 ...)."""
import json
import os

import pytest

from glpk_js_b200 import glpk as F
import helpers as H
import test_presolve as T

with open(os.path.join(H.GOLDEN, "ref_writer_cases.json")) as f:
    CASES = {k: v for k, v in json.load(f).items() if not k.startswith("_")}


@pytest.mark.parametrize("name", sorted(CASES))
def test_writer_matches_reference(name):
    case = CASES[name]
    if "text" in case:
        P = F.glp_create_prob()
        assert F.glp_read_lp_from_string(P, None, case["text"]) == 0
    else:
        P = T.facade_problem(case["problem"])
    if "names" in case:
        nm = case["names"]
        P.name, P.obj = nm["prob"], nm["obj"]
        for i, x in enumerate(nm["rows"], 1):
            P.row[i].name = x
        for j, x in enumerate(nm["cols"], 1):
            P.col[j].name = x
    for writer in (F.glp_write_lp, F._write_lp_py):      # the native writer and its Python cross-check
        lines, msgs = [], []
        F.glp_set_print_func(msgs.append)
        try:
            ret = writer(P, None, lines.append)
        finally:
            F.glp_set_print_func(None)
        assert ret == case["ret"]
        assert lines == case["lines"]
        assert msgs == case["messages"]
