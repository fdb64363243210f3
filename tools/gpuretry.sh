#!/bin/bash
# usage: gpuretry.sh OUTFILE [gpurun args...]  -- retries while the pod answers busy/transient (nothing is charged for those)
out=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > $out 2>&1
  if ! grep -q "status=transient\|retry in a few minutes\|rc=3" $out; then break; fi
  sleep 150
done
echo finished >> $out
