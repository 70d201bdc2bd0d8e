#!/usr/bin/env python
"""What the REFERENCE'S OWN reader (glp_read_lp, lib/glpcpx.js, executed by minijs) makes of a set
of CPLEX-LP texts -- valid ones (names, types, bounds, kinds, costs, matrix) and invalid ones (the
message it throws, with the line number).  Output: tests/golden/ref_reader_cases.json; the native
reader glpb_read_lp must reproduce every case (tests/test_hostprep.py).
Run in the build container only:   python oracle/jsref/make_reader_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import refjs  # noqa: E402
from minijs import JSThrow, NativeFunc  # noqa: E402

CASES = {
    "colon_after_blank": "Maximize\n obj : 3 x + 2 y\nSubject To\n c1 : x + y <= 4\n c2: x + 3 y <= 6\nEnd\n",
    "unnamed_rows": "Minimize\n x + y\nSubject To\n x + y >= 2\n\n x - y <= 1\n c3: x >= 0.5\n 2 x + y >= 1\nEnd\n",
    "keyword_indented": "Maximize\n obj: x + y\n Subject To\n c1: x + y <= 4\nEnd\n",
    "beyond_rhs": "Maximize\n obj: x + y\nSubject To\n c1: x + y <= 4 y\nEnd\n",
    "comment_after_rhs": "Maximize\n obj: x + y\nSubject To\n c1: x + y <= 4 \\ no comment allowed here\nEnd\n",
    "bad_exponent": "Maximize\n obj: 3e x + y\nSubject To\n c1: x + y <= 4\nEnd\n",
    "lonely_point": "Maximize\n obj: . x + y\nSubject To\n c1: x + y <= 4\nEnd\n",
    "bounds_general": "Minimize\n obj: x + y + z\nSubject To\n c1: x + y + z >= 1.5\nBounds\n x <= 4\n -1 <= y <= 2\n z free\n"
                      "General\n x\nBinary\n y\nEnd\n",
    "such_that_split": "Minimize\n x\nsuch that\n c1: x >= 1\nend\n",
    "st_dot": "max\n 2 a + 3 b\ns.t.\n a + b <= 10\n a - b >= -2\nbound\n a <= 6\nend\n",
    "missing_sense": "Minimize\n x\nSubject To\n c1: x + y\nEnd\n",
    "missing_rhs": "Minimize\n x\nSubject To\n c1: x + y >=\nEnd\n",
    "multiple_use": "Minimize\n x\nSubject To\n c1: x + y + x >= 1\nEnd\n",
    "row_twice": "Minimize\n x\nSubject To\n c1: x >= 1\n c1: y >= 1\nEnd\n",
    "no_end": "Minimize\n x\nSubject To\n c1: x >= 1\n",
    "extra_after_end": "Minimize\n x\nSubject To\n c1: x >= 1\nEnd\n x\n",
    "plus_inf_lower": "Minimize\n x\nSubject To\n c1: x >= 1\nBounds\n +inf <= x\nEnd\n",
    "fixed_and_neg": "Minimize\n - x + 2 y\nSubject To\n c1: - x - y >= -8\n c2: x = 3\nBounds\n y = 2.5\n x >= -1e1\nEnd\n",
    "zero_coefficients": "Minimize\n 0 x + y\nSubject To\n c1: 0 x + y + 0 z >= 1\n c2: x + z <= 2\nEnd\n",
    "names_with_symbols": "Minimize\n x_1 + y.2 + z(3)\nSubject To\n r!: x_1 + y.2 >= 1\n q#1: z(3) <= 7\nEnd\n",
}


def main():
    I = refjs.load()
    api, g = I.api, I.globals
    msgs = []
    api["glp_set_print_func"].call(g, [NativeFunc(lambda this, a: msgs.append(str(a[0])), "print")])

    def call(name, *a):
        return api[name].call(g, list(a))

    texts = dict(CASES)
    for fx in ("test", "gap", "todd"):
        texts["fixture_" + fx] = open(os.path.join(refjs.REF, "test", fx + ".lpt")).read()
    out = {}
    for name, text in texts.items():
        del msgs[:]
        lp = call("glp_create_prob")
        try:
            rc = refjs.read_lp_text(I, lp, text)
            err = None
        except JSThrow as e:
            rc, err = 1, str(e)
        rec = {"text": text, "rc": int(rc), "error": err, "warnings": [m for m in msgs if "warning" in m]}
        if rc == 0:
            m, n = call("glp_get_num_rows", lp), call("glp_get_num_cols", lp)
            rec.update(m=m, n=n, dir=call("glp_get_obj_dir", lp), obj_name=call("glp_get_obj_name", lp),
                       rows=[dict(name=call("glp_get_row_name", lp, i), type=call("glp_get_row_type", lp, i),
                                  lb=float(lp["row"][i]["lb"]), ub=float(lp["row"][i]["ub"])) for i in range(1, m + 1)],
                       cols=[dict(name=call("glp_get_col_name", lp, j), type=call("glp_get_col_type", lp, j),
                                  lb=float(lp["col"][j]["lb"]), ub=float(lp["col"][j]["ub"]), kind=call("glp_get_col_kind", lp, j),
                                  coef=float(call("glp_get_obj_coef", lp, j))) for j in range(1, n + 1)])
            mat = []
            for j in range(1, n + 1):
                col, aij = [], lp["col"][j]["ptr"]
                while aij is not None:
                    col.append([int(aij["row"]["i"]), float(aij["val"])])
                    aij = aij["c_next"]
                mat.append(col)
            rec["columns"] = mat          # in list order = ascending rows after glp_sort_matrix
        out[name] = rec
        print(name, rc, err, rec.get("m"), rec.get("n"), flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "ref_reader_cases.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
