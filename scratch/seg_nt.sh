#!/bin/bash
# cfg2 over CTA shapes of classify_seg_kernel (options seg_nt / seg_wt)
cd "$(dirname "$0")/.."
for o in "" "--opt seg_wt=256" "--opt seg_wt=256 --opt seg_nt=704" "--opt seg_wt=256 --opt seg_nt=640" "--opt seg_wt=256 --opt seg_nt=512" "--opt seg_nt=512" "--opt seg_wt=256 --opt seg_nt=768"; do
  echo "== $o"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-extra $o 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d.get('parity_on_sample'))"
done
