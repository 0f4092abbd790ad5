#!/bin/bash
cd "$(dirname "$0")/.."
for o in "" "--opt strata_noreg=1" "--opt strata_loadfirst=1" "--opt strata_noreg=1 --opt strata_loadfirst=1"; do
  timeout 100 python bench.py --workload cfg5 --steps 5 --warmup 2 --no-e2e --no-cpu $o 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$o', d['roofline']['kernel_ms'])"
done
