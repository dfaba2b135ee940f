#!/bin/bash
# local helper (build container): keep asking for a GPU box until the call is accepted (exit code 3 = no slot right now)
#   scripts/gpurun_retry.sh <log> <gpurun args...>
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then break; fi
  sleep 90
done
echo "gpurun rc=$rc" >> "$log"
echo finished >> "$log"
