#!/bin/bash
# usage: [GPUS=N] scripts/gpurun_retry.sh <timeout-seconds> <command...>   -- retries while the pod answers "busy" (exit 3)
T=$1; shift
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $T ${GPUS:+--gpus $GPUS} -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] pod busy (attempt $attempt), sleeping 150 s"
  sleep 150
done
exit 3
