#!/bin/bash
# tools/gpurun_retry.sh <tries> <gpurun args...>: re-issue a gpurun call while the pod answers "transient" (nothing charged)
tries=$1; shift
for i in $(seq 1 "$tries"); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out"
  if ! echo "$out" | grep -q "status=transient"; then exit 0; fi
  echo "[retry $i/$tries] pod busy, sleeping 90 s"
  sleep 90
done
exit 3
