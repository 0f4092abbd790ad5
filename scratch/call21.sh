#!/bin/bash
cd /root/repo
for o in "" "--opt strata_dbg=1" "--opt strata_dbg=2" "--opt strata_dbg=3"; do
  echo "== $o" >> gpurun_out/c21_cfg5.log
  timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-e2e --no-cpu $o 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d.get('parity'))" >> gpurun_out/c21_cfg5.log 2>&1
done
cat gpurun_out/c21_cfg5.log
