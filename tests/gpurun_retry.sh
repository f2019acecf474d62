#!/bin/bash
# usage: tests/gpurun_retry.sh <timeout-seconds> '<command>'   -- retries while the pod answers "busy" (exit 3)
for try in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
