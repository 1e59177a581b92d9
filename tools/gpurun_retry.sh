#!/bin/bash
# gpurun with retries while the pod answers "busy / no slot" (exit 3, nothing charged).
# usage: tools/gpurun_retry.sh <log> <gpurun args...>
log="$1"; shift
for attempt in $(seq 1 40); do
  gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc after $attempt attempt(s)" >> "$log"; exit $rc; fi
  sleep 90
done
echo "gpurun: gave up after 40 busy answers" >> "$log"; exit 3
