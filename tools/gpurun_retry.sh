#!/bin/bash
# gpurun_retry.sh TIMEOUT 'command' -- retries while the pod answers busy (rc 3 / "transient"), every 2 minutes, for up to an hour
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "$2" 2>&1); echo "$out" | tail -120
  if echo "$out" | grep -q "status=transient\|status=busy\|no box\|rc=3"; then sleep 120; continue; fi
  break
done
