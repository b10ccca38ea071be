set -x
for d in _old .; do
  (cd $d && ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_active --clock-control none -k regex:phmm_flat_f32_kernel -c 8 --csv --log-file /tmp/ncu_$$.csv python bench.py --regions 300 --steps 1 --warmup 3 --no-cpu-baseline $( [ "$d" = "." ] && echo --no-prefix-sharing ) > /dev/null 2>&1; python - <<PY
import csv
rows=[r for r in csv.reader(open("/tmp/ncu_$$.csv")) if len(r)>10 and r[0].isdigit()]
from collections import defaultdict
d=defaultdict(dict)
for r in rows: d[r[0]][r[-3]]=float(r[-1].replace(",",""))
for k,v in list(d.items())[:8]: print("$d", k, rows[[x[0] for x in rows].index(k)][4][:40], v)
PY
)
done
