import json,subprocess,sys,os
for eng in ("0","1"):
    e=dict(os.environ); e["GLPB_ENGINE"]=eng
    out=subprocess.run([sys.executable,"tests/run_pivots.py"],capture_output=True,text=True,env=e)
    if out.returncode!=0: print(out.stderr[-2000:]); continue
    res=json.loads(out.stdout.strip().splitlines()[-1])
    for r in res:
        if r["name"] in ("test","gap","todd") or r["name"].startswith("transport") or not r["same_sequence"]:
            print(eng, r["name"], r["meth"], "same" if r["same_sequence"] else "DIFF", "it", r["iterations"], r["ref_iterations"], "first", r["first_difference"], "ties", r["ties"], r["around"])
