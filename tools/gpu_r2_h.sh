#!/bin/bash
mkdir -p gpurun_out
python tools/prof_compact5.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_compact5.csv python tools/prof_compact5.py > gpurun_out/ncu_c5.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_compact5.csv')) if len(r)>10]
h=rows[0]; ix={k:i for i,k in enumerate(h)}
for r in rows[1:]:
    if r[ix['Metric Name']]=='gpu__time_duration.sum' or 'dram' in r[ix['Metric Name']]:
        print(r[ix['ID']], r[ix['Kernel Name']][:28], r[ix['Grid Size']], r[ix['Block Size']], r[ix['Metric Name']][:22], r[ix['Metric Value']], r[ix['Metric Unit']])
PY
