#!/bin/bash
# usage: scripts/gpurun_retry.sh OUTFILE TIMEOUT [--gpus N] -- CMD   (retries while the pod answers busy / transient)
OUT=$1; shift; TMO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO "$@" > $OUT 2>&1
  if grep -q "status=transient\|rc=3\|no box\|busy" $OUT && ! grep -q "status=ok" $OUT; then sleep 90; continue; fi
  break
done
