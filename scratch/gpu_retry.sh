#!/bin/bash
# retry a gpurun call while the pod answers "busy" (exit code 3); usage: gpu_retry.sh TIMEOUT 'command'
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2" > /tmp/gpu_retry.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then cat /tmp/gpu_retry.log | tail -40; exit $rc; fi
  sleep 90
done
echo "gave up"; exit 3
