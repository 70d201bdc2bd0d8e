"""Stand-alone streaming measurements of the per-kernel path (glpb_bench_kernel): the
pricing scans, the pivot-row SpMV and the rank-1 basis update over resident synthetic
data, from the sizes of the benchmark LPs (L2-resident, launch-bound) up to sizes where
the kernels are HBM-bound.  Prints one line per (kernel, size): us/launch, algorithmic
GB/s and the fraction of the measured copy peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glpk_js_b200 as G
nat = G.native
peak = 6453.1
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
        pk = json.load(f)
    for key in ("hbm_gbs", "hbm_gbps"):
        if key in pk:
            peak = float(pk[key])
except Exception:
    pass
cases = [("chuzc_primal", 0, n) for n in (4096, 32768, 1 << 20, 1 << 24, 1 << 26)]
cases += [("chuzr_dual", m, 0) for m in (16384, 1 << 20, 1 << 24)]          # basis header scattered at random
cases += [("chuzr_dual_seq", m, 0) for m in (1 << 20, 1 << 24)]             # header ascending (as after glp_factorize)
cases += [("trow", 16384, 32768), ("trow", 1 << 20, 1 << 21), ("trow", 1 << 22, 1 << 23)]    # rho gathered from 8-32 MB (L2)
cases += [("trow", 16384, 1 << 22), ("trow", 65536, 1 << 23)]                                 # the C3 row count, more columns
cases += [("update_rank1", k, 0) for k in (609, 2048, 6353)]
print("peak (measured copy) %.1f GB/s" % peak)
for name, m, n in cases:
    try:
        us, nb = nat.bench_kernel(name, m, n, reps=20 if max(m, n) >= (1 << 22) else 200)
        print("%-14s m=%-9d n=%-9d %10.2f us/launch %9.1f GB/s  %5.1f%% of peak  (%.3g algorithmic bytes)"
              % (name, m, n, us, nb / us / 1e3, 100.0 * nb / us / 1e3 / peak, nb))
    except Exception as e:
        print("%-14s m=%d n=%d failed: %s" % (name, m, n, e))
