B="python bench.py --workload c3 --steps 1 --warmup 0 --no-c2 --no-bnb --no-cpu-baseline --no-profile"
for v in 1 2 4 1; do
  echo "== C3 MUL=$v"; GLPB_REFAC_MUL=$v $B 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['iterations_per_step'], d['refactorizations'], d['parity'])"
done
B2="python bench.py --workload c2 --steps 3 --warmup 1 --no-bnb --no-cpu-baseline --no-profile"
for v in 1 2 4; do
  echo "== C2 MUL=$v"; GLPB_REFAC_MUL=$v $B2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('iterations_per_step'), d.get('refactorizations'), d.get('parity'))"
done
