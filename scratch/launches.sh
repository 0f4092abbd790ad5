#!/bin/bash
# ncu launch list (time + DRAM bytes) of one bench workload: launches.sh cfg3 [bench args]
cd "$(dirname "$0")/.."
w=$1; shift
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"ordinal|classify|seg_|strata|fill_slots|pk_|compact|unpack" -c 60 --csv --log-file gpurun_out/r2_launches_$w.csv \
  python bench.py --workload $w "$@" --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/launch_$w.log 2>&1
python - $w <<'PY'
import csv, sys, collections
w=sys.argv[1]
rows=[r for r in csv.reader(open(f'gpurun_out/r2_launches_{w}.csv')) if len(r)>14 and r[0].isdigit()]
per=collections.OrderedDict()
for r in rows:
    per.setdefault(r[0], [r[4][:70], {}])[1][r[12]]=r[14]
for k,(name,m) in per.items():
    print(k, name, m.get('gpu__time_duration.sum'), m.get('dram__bytes_read.sum'), m.get('dram__bytes_write.sum'))
PY
