#!/usr/bin/env python
"""Source-line view of an ncu capture for kernels whose code is spread over several
headers (ncu's CUDA-C source page exports only the kernel's own file): joins the SASS
page of the report (`ncu -i X.ncu-rep --page source --csv --print-source sass`) with the
line table of the cubin (`nvdisasm -g -c`), instruction by instruction, and prints the
source lines with the most executed warp instructions / stall samples.

    python tools/ncu_lines.py REPORT.ncu-rep LIB.so MANGLED_KERNEL [top]
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
lines = []
for cub in sorted(os.listdir(tmp)):
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    if ".text." + kern not in out:
        continue
    sec = out.split(".text." + kern + " ---", 1)[1]
    sec = sec.split("//--------------------- .", 1)[0]
    cur = ("?", 0)
    for ln in sec.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines.append((int(m.group(1), 16), cur, m.group(2).strip()))
    break
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
assert len(body) == len(lines), (len(body), len(lines))
acc = collections.defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for r, (_, cur, _) in zip(body, lines):
    a = acc[cur]
    a[0] += int(r[ci] or 0); a[1] += int(r[cs] or 0)
    tot_i += int(r[ci] or 0); tot_s += int(r[cs] or 0)
src = {}
def text(f, l):
    if f not in src:
        for root in ("glpk.js_b200/csrc", "."):
            p = os.path.join(root, f)
            if os.path.exists(p):
                src[f] = open(p).read().splitlines(); break
        else:
            src[f] = []
    return src[f][l - 1].strip()[:110] if 0 < l <= len(src[f]) else ""
print("total warp instructions %d, stall samples %d" % (tot_i, tot_s))
for (f, l), (ni, ns) in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100.0 * ni / max(1, tot_i), 100.0 * ns / max(1, tot_s), f, l, text(f, l)))
