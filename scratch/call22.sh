#!/bin/bash
cd /root/repo
python - <<'PY' > gpurun_out/c22_limit.log 2>&1
import ctypes
rt = ctypes.CDLL('libcudart.so.12')
v = ctypes.c_size_t()
print('get', rt.cudaDeviceGetLimit(ctypes.byref(v), 5), v.value)   # cudaLimitMaxL2FetchGranularity = 0x05
PY
for o in "" "--opt l2_fetch=32" "--opt l2_fetch=64"; do
  for w in cfg5 cfg3 cfg2; do
  echo "== $w $o" >> gpurun_out/c22.log
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu $o 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])" >> gpurun_out/c22.log 2>&1
  done
done
cat gpurun_out/c22_limit.log gpurun_out/c22.log
