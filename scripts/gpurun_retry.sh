#!/bin/bash
# usage: scripts/gpurun_retry.sh <out-file> <gpurun args...>   -- retries while the pod answers "busy" (nothing charged)
out=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  rc=$?
  if grep -q "status=transient\|nothing was charged" "$out" && ! grep -q "exit " "$out"; then sleep 120; continue; fi
  break
done
echo "gpurun rc=$rc attempts=$i" >> "$out"
