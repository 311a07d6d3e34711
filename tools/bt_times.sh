#!/bin/bash
# per-kernel device times of one Gon-gitsune run (ncu launch list): python tools/profile_run.py under ncu
ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/profile_run.py --workload ${1:-gon} --runs 2 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]; h=rows[hi]; ix={k:i for i,k in enumerate(h)}
for r in rows[hi+1:]:
    if len(r)>=len(h) and r[ix['Metric Name']]=='gpu__time_duration.sum': print(r[ix['Kernel Name']][:50].ljust(52), r[ix['Metric Value']], r[ix['Metric Unit']])
"
