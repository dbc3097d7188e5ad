#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>' [gpus]   -- retries while the pod answers "transient" / busy
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 30); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$CMD" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD" 2>&1); fi
  RC=$?
  if echo "$OUT" | grep -q "status=transient\|status=busy" || [ $RC -eq 3 ]; then echo "[retry $i] transient/busy (rc=$RC)"; sleep 90; continue; fi
  echo "$OUT"; exit $RC
done
echo "gave up"; exit 3
