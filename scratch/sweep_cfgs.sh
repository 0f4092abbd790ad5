#!/bin/bash
# run cfg2 bench over sweep kernel configurations (WK_TUNE_BLOCK = S*10000 + NT, WK_SWEEP_R)
for cfg in "$@"; do
  tb=${cfg%%:*}; r=${cfg##*:}
  out=$(WK_TUNE_BLOCK=$tb WK_SWEEP_R=$r python bench.py --no-cpu --no-e2e --steps 5 2>&1 | tail -1)
  echo "cfg $cfg -> $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), "ms")' 2>&1 | tail -1)"
done
