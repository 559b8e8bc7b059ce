#!/bin/bash
# Development helper: retry a gpurun call while the pod answers "transient" (nothing charged).
# usage: tools/gpurun_retry.sh TIMEOUT 'command'
for i in $(seq 1 20); do
  out=$(gpurun --timeout "$1" -- "$2" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "gave up: pod busy"; exit 3
