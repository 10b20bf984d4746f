#!/bin/bash
# gp.sh <timeout-seconds> <command...>: gpurun with retries while the pod answers busy (exit code 3 / transient)
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  if [ $rc -eq 3 ]; then sleep 45; continue; fi
  echo "$out" | tail -${TAILN:-25}; exit $rc
done
echo "gave up"; exit 3
