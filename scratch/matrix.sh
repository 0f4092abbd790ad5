#!/bin/bash
# seg kernel vs run-per-lane kernel on the plans both can take
pr='import json,sys; d=json.loads(sys.stdin.read()); print(d["roofline"]["kernel"], round(d["roofline"]["kernel_ms"],3))'
for noseg in "" 1; do
  for args in "--mode above" "--mode uniq" "--ranks phylum,genus,species" "--ranks phylum,genus,species --mode above" "--workload cfg4" "--workload cfg4 --mode default" "--ranks species --mode above" "--ranks phylum --mode above"; do
    echo -n "noseg=$noseg $args: "
    env ${noseg:+WK_NO_SEG=1} python bench.py $args --no-cpu --no-e2e --steps 5 2>&1 | tail -1 | python -c "$pr"
  done
done
