# last GPU call of round 2: launch list of a complete C2 solve on the final build, then two probes of longer
# refactorisation periods on C3 (information for DESIGN "Next", not adopted)
C2="python bench.py --workload c2 --steps 1 --warmup 0 --no-profile --no-bnb --no-cpu-baseline"
$C2 > gpurun_out/r03_bench_c2_plain_before_ncu.json 2> /dev/null || exit 1
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r03_launches_c2.csv $C2 > gpurun_out/r03_launches_c2.out 2>&1
B="python bench.py --workload c3 --steps 1 --warmup 0 --no-c2 --no-bnb --no-cpu-baseline --no-profile"
for v in 8 16; do
  echo "== C3 MUL=$v"; GLPB_REFAC_MUL=$v timeout 60 $B 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['iterations_per_step'], d['refactorizations'], d['parity'])"
done
