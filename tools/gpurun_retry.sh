#!/usr/bin/env bash
# gpurun with retries while the pod answers "busy / draining" (nothing is charged for those): usage gpurun_retry.sh <gpurun args...>
for i in $(seq 1 20); do
  OUT=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$OUT"
  if echo "$OUT" | grep -q "status=transient"; then
    sleep 150
    continue
  fi
  break
done
