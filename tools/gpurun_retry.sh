#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <command...>   retries while the pod answers "transient / busy" (nothing is charged for those)
T=$1; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|retry in a few minutes"; then sleep 90; continue; fi
  echo "$out" | tail -60; exit $rc
done
echo "gave up after 20 transient answers"; exit 3
