#!/bin/bash
# pool_head under ncu (10 launches of the bench forward, graph replay off for the profiled process)
cd "$(dirname "$0")/.."
CNS_NO_GRAPHS=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pool_head -s 5 -c 10 --csv --log-file gpurun_out/ph.csv python scripts/enc_edge_time.py > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/ph.csv') if l.startswith('"'))]
v=[float(r[-1]) for r in rows[1:]]
print("pool_head %.1f us (n=%d)" % (sum(v)/len(v)/1e3, len(v)))
PY
