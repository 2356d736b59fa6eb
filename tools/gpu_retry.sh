#!/bin/bash
# usage: tools/gpu_retry.sh OUTFILE TIMEOUT 'command'   -- retries gpurun while the pod answers "transient/busy" (exit 3)
OUT=$1; TO=$2; CMD=$3
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$TO" -- "$CMD" > "$OUT" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$OUT"; then exit $rc; fi
  sleep 90
done
exit 3
