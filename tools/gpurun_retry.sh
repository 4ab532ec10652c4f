#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   — retries while the pod answers busy / transient (nothing charged)
log=$1; shift
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then echo "exit=$rc" >> "$log"; exit $rc; fi
  sleep 60
done
echo "exit=3 (gave up)" >> "$log"
